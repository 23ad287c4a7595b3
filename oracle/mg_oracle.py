"""ctypes binding of oracle/libmg_oracle.so (the CPU restatement).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (marlgrid_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmg_oracle.so")

_lib = None


def build(force=False):
    src = os.path.join(HERE, "mg_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "marlgrid_b200.h")
    if (
        force
        or not os.path.exists(LIB_PATH)
        or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    ):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mgo_get_threads.restype = ctypes.c_int
        _lib.mgo_hw_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class OracleBatch:
    """B independent envs stepped by the C restatement, on numpy buffers with the device layout."""

    def __init__(self, cfg, n_envs, seed=1337, env_offset=0, threads=1):
        self.cfg = cfg
        self.B = int(n_envs)
        self.seed = int(seed)
        self.env_offset = int(env_offset)
        self.A = cfg.n_agents
        self.V = cfg.view_size
        self.ts = cfg.view_tile_size
        self.grid = np.zeros((self.B, 3, cfg.plane_stride), np.uint8)
        self.agents = np.zeros((self.B, self.A, 16), np.uint8)
        self.envrec = np.zeros((self.B, 4), np.int32)
        self.prestige = np.zeros((self.B, self.A), np.float64)  # GridAgentInterface.prestige (agents.py:141-153)
        self.threads = threads
        L = lib()
        L.mgo_init(ctypes.byref(cfg), _p(self.grid), _p(self.agents), _p(self.envrec), ctypes.c_int64(self.B))

    def _thr(self):
        lib().mgo_set_threads(int(self.threads))
        lib().mgo_set_prestige(_p(self.prestige))  # (a library-wide pointer: set before every call of this batch)

    def reset(self, mask=None):
        self._thr()
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().mgo_reset(
            ctypes.byref(self.cfg), _p(self.grid), _p(self.agents), _p(self.envrec), ctypes.c_int64(self.B),
            ctypes.c_uint64(self.seed), ctypes.c_int64(self.env_offset), _p(m),
        )

    def step(self, actions, autoreset=False, with_obs=False):
        self._thr()
        act = np.ascontiguousarray(actions, np.int32).reshape(self.B, self.A)
        rew = np.zeros((self.B, self.A), np.float64)
        done = np.zeros((self.B,), np.uint8)
        obs = np.zeros((self.B, self.A, self.V, self.V, 3), np.uint8) if with_obs else None
        lib().mgo_step(
            ctypes.byref(self.cfg), _p(self.grid), _p(self.agents), _p(self.envrec), ctypes.c_int64(self.B),
            ctypes.c_uint64(self.seed), ctypes.c_int64(self.env_offset), _p(act), _p(rew), _p(done),
            ctypes.c_int(int(autoreset)), _p(obs),
        )
        if with_obs:
            return obs, rew, done
        return rew, done

    def rollout(self, actions, autoreset=True, n_steps=None):
        """actions int32 [P, B, A]; n_steps (default P) fused steps per env inside C, cycling through the
        P action batches (threads own env ranges)."""
        self._thr()
        act = np.ascontiguousarray(actions, np.int32)
        P = act.shape[0]
        T = P if n_steps is None else int(n_steps)
        assert act.shape[1:] == (self.B, self.A)
        if not hasattr(self, "_ro"):
            self._ro = (np.zeros((self.B, self.A, self.V, self.V, 3), np.uint8), np.zeros((self.B, self.A), np.float64), np.zeros((self.B,), np.uint8))
        obs, rew, done = self._ro
        lib().mgo_rollout(
            ctypes.byref(self.cfg), _p(self.grid), _p(self.agents), _p(self.envrec), ctypes.c_int64(self.B), ctypes.c_uint64(self.seed),
            ctypes.c_int64(self.env_offset), _p(act), ctypes.c_int64(T), ctypes.c_int64(P), _p(rew), _p(done), ctypes.c_int(int(autoreset)), _p(obs),
        )
        return obs, rew, done

    def obs_encode(self):
        self._thr()
        obs = np.zeros((self.B, self.A, self.V, self.V, 3), np.uint8)
        lib().mgo_obs_encode(ctypes.byref(self.cfg), _p(self.grid), _p(self.agents), _p(self.envrec), ctypes.c_int64(self.B), _p(obs))
        return obs

    def obs_rgb(self, atlas):
        self._thr()
        n = self.V * self.ts
        obs = np.zeros((self.B, self.A, n, n, 3), np.uint8)
        atlas = np.ascontiguousarray(atlas, np.uint8)
        lib().mgo_obs_rgb(ctypes.byref(self.cfg), _p(self.grid), _p(self.agents), _p(self.envrec), ctypes.c_int64(self.B), _p(atlas), _p(obs))
        return obs

    def vis(self):
        out = np.zeros((self.B, self.A, self.V, self.V), np.uint8)
        lib().mgo_vis(ctypes.byref(self.cfg), _p(self.grid), _p(self.agents), _p(self.envrec), ctypes.c_int64(self.B), _p(out))
        return out

    # ---- decoded views of the SoA records -------------------------------------------------
    def planes(self):
        W, H = self.cfg.width, self.cfg.height
        return self.grid[:, :, : W * H].reshape(self.B, 3, W, H)

    @property
    def agent_x(self):
        return self.agents[:, :, 0]

    @property
    def agent_y(self):
        return self.agents[:, :, 1]

    @property
    def agent_dir(self):
        return self.agents[:, :, 2]

    @property
    def agent_flags(self):
        return self.agents[:, :, 3]

    @property
    def agent_carry(self):
        return self.agents[:, :, 4:7]

    @property
    def agent_stamp(self):
        return self.agents[:, :, 8:12].copy().view(np.int32)[..., 0]

    @property
    def step_count(self):
        return self.envrec[:, 0]

    @property
    def err(self):
        return (self.envrec[:, 3].view(np.uint32) >> 16).astype(np.int32)

    def agent_rank(self):
        """Queue position of each placed agent on its cell (0 = head), -1 if not placed."""
        rank = np.full((self.B, self.A), -1, np.int32)
        st = self.agent_stamp
        for e in range(self.B):
            for a in range(self.A):
                if not (self.agent_flags[e, a] & 1):
                    continue
                r = 0
                for q in range(self.A):
                    if q != a and (self.agent_flags[e, q] & 1) and self.agent_x[e, q] == self.agent_x[e, a] and self.agent_y[e, q] == self.agent_y[e, a] and st[e, q] < st[e, a]:
                        r += 1
                rank[e, a] = r
        return rank


def los_batch(transparent, ax, ay):
    t = np.ascontiguousarray(transparent, np.uint8)
    n, V, _ = t.shape
    out = np.zeros_like(t)
    lib().mgo_los_batch(_p(t), _p(out), ctypes.c_int64(n), ctypes.c_int(V), ctypes.c_int(ax), ctypes.c_int(ay))
    return out


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    o = np.zeros(4, np.uint32)
    lib().mgo_philox(_p(c), _p(k), _p(o))
    return tuple(int(x) for x in o)


def order(seed, g, t, A):
    o = np.zeros(A, np.int32)
    lib().mgo_order(ctypes.c_uint64(seed), ctypes.c_uint64(g), ctypes.c_uint32(t), ctypes.c_int(A), _p(o))
    return [int(x) for x in o]
