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

        class Box(Space):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                super().__init__(shape if shape is not None else np.shape(low), dtype)
                self.low, self.high = low, high

            def __repr__(self):
                return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"

        class Discrete(Space):
            def __init__(self, n):
                super().__init__((), np.int64)
                self.n = int(n)

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

        class Dict(Space):
            def __init__(self, spaces=None, **kw):
                super().__init__(None, None)
                self.spaces = dict(spaces or {}, **kw)

            def __getitem__(self, k):
                return self.spaces[k]
