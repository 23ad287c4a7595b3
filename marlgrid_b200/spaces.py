"""Minimal gym-compatible space holders (gym itself is an optional dependency).

The reference builds `gym.spaces.Tuple` of per-agent `Discrete(7)` / `Box(0, 255, (V*ts, V*ts, 3), uint8)`
(marlgrid/base.py:376-386, marlgrid/agents.py:58-83).  When gym/gymnasium is importable its classes are
used so downstream code sees the real thing; otherwise these stand-ins expose the same attributes.
"""
import numpy as np

try:  # pragma: no cover - depends on the environment
    from gymnasium.spaces import Box, Discrete, Tuple, Dict  # type: ignore
except Exception:  # noqa: BLE001
    try:
        from gym.spaces import Box, Discrete, Tuple, Dict  # type: ignore
    except Exception:  # noqa: BLE001

        class Space:
            def __init__(self, shape=None, dtype=None):
                self.shape = None if shape is None else tuple(shape)
                self.dtype = None if dtype is None else np.dtype(dtype)
                self._rng = np.random.RandomState()

            def seed(self, seed=None):
                self._rng = np.random.RandomState(seed)
                return [seed]

            def sample(self):
                raise NotImplementedError

        class Box(Space):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                super().__init__(shape if shape is not None else np.shape(low), dtype)
                self.low, self.high = low, high

            def sample(self):
                if np.issubdtype(self.dtype, np.integer):
                    return self._rng.randint(int(np.min(self.low)), int(np.max(self.high)) + 1, size=self.shape).astype(self.dtype)
                lo = np.broadcast_to(np.asarray(self.low, dtype=np.float64), self.shape)
                hi = np.broadcast_to(np.asarray(self.high, dtype=np.float64), self.shape)
                u = self._rng.uniform(size=self.shape)
                bounded = np.isfinite(lo) & np.isfinite(hi)
                return np.where(bounded, lo + u * (hi - lo), self._rng.normal(size=self.shape)).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

            def __repr__(self):
                return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"

        class Discrete(Space):
            def __init__(self, n):
                super().__init__((), np.int64)
                self.n = int(n)

            def sample(self):
                return int(self._rng.randint(self.n))

            def contains(self, x):
                return 0 <= int(x) < self.n

            def __repr__(self):
                return f"Discrete({self.n})"

        class Tuple(Space):
            def __init__(self, spaces):
                super().__init__(None, None)
                self.spaces = tuple(spaces)

            def __len__(self):
                return len(self.spaces)

            def __getitem__(self, i):
                return self.spaces[i]

            def __iter__(self):
                return iter(self.spaces)

            def sample(self):
                return tuple(sp.sample() for sp in self.spaces)

            def contains(self, x):
                return len(x) == len(self.spaces) and all(sp.contains(v) for sp, v in zip(self.spaces, x))

            def __repr__(self):
                return "Tuple(" + ", ".join(repr(sp) for sp in self.spaces) + ")"

        class Dict(Space):
            def __init__(self, spaces=None, **kw):
                super().__init__(None, None)
                self.spaces = dict(spaces or {}, **kw)

            def __getitem__(self, k):
                return self.spaces[k]

            def sample(self):
                return {k: sp.sample() for k, sp in self.spaces.items()}

            def contains(self, x):
                return set(x) == set(self.spaces) and all(sp.contains(x[k]) for k, sp in self.spaces.items())

            def __repr__(self):
                return "Dict(" + ", ".join(f"{k}: {sp!r}" for k, sp in self.spaces.items()) + ")"
