"""On-device policy hand-off: the caller loop of the reference's README.md:43-57

    act = agents.action_step(obs); obs, rew, done, _ = env.step(act)

closed on the GPU.  `LinearPolicy` is the policy the rollout kernel can evaluate itself: one int8 linear layer per agent over
the encoded observation, argmax (lowest index on ties), epsilon-greedy exploration from a Philox stream -- exact integer
arithmetic, so a rollout is reproducible bit for bit (tests replay it on the CPU).  `env.rollout_policy(policy, first_actions,
n_steps)` plays n_steps steps in ONE kernel launch for the registered shapes (include/marlgrid_b200.h: mg_rollout_policy) and
returns every step's observations, rewards, done flags and the actions played -- the learner's batch.
"""
import ctypes

import numpy as np


class MgLinearPolicy(ctypes.Structure):
    """include/marlgrid_b200.h: MgLinearPolicy."""

    _fields_ = [
        ("weights", ctypes.c_void_p),
        ("bias", ctypes.c_void_p),
        ("n_actions", ctypes.c_int32),
        ("epsilon", ctypes.c_uint32),
        ("seed", ctypes.c_uint64),
    ]


def n_obs_words(view_size):
    """Words of 4 observation bytes per agent view (the last one zero-padded)."""
    return (view_size * view_size * 3 + 3) // 4


def pack_weights(weights, view_size):
    """int8 [A][n_actions][V*V*3] -> the kernels' layout int8 [A][NW][8][4]: byte b of word i of action k's row =
    w[a][k][4 i + b]; zero beyond the observation and for k >= n_actions."""
    w = np.asarray(weights)
    if w.dtype != np.int8:
        if np.any(w != np.clip(np.rint(w), -128, 127)):
            raise ValueError("policy weights must be integers in [-128, 127]")
        w = w.astype(np.int8)
    A, K, n = w.shape
    if n != view_size * view_size * 3 or not 1 <= K <= 7:
        raise ValueError(f"weights must be [A][n_actions <= 7][{view_size * view_size * 3}], got {w.shape}")
    nw = n_obs_words(view_size)
    full = np.zeros((A, 8, nw * 4), np.int8)
    full[:, :K, :n] = w
    return np.ascontiguousarray(full.reshape(A, 8, nw, 4).transpose(0, 2, 1, 3))


def pack_bias(bias, n_agents, n_actions):
    b = np.zeros((n_agents, 8), np.int32)
    if bias is not None:
        b[:, :n_actions] = np.asarray(bias, np.int64).reshape(n_agents, n_actions)
    return b


def epsilon_to_u32(epsilon):
    """Exploration probability -> threshold on a uniform u32 (the kernels explore when draw < threshold)."""
    if not 0.0 <= epsilon <= 1.0:
        raise ValueError("epsilon must be in [0, 1]")
    return min(int(round(epsilon * 4294967296.0)), 0xFFFFFFFF)


class LinearPolicy:
    """weights int8 [A][n_actions][V*V*3] (observation bytes in the order of the encoded view, agents.py:268-296 /
    base.py:196-214), bias int32 [A][n_actions]; action = argmax_k(bias[a][k] + weights[a][k] . obs), lowest k on ties; with
    probability `epsilon` a uniform action in [0, n_actions) instead."""

    def __init__(self, weights, bias=None, epsilon=0.0, seed=0, view_size=None):
        w = np.asarray(weights)
        if w.ndim != 3:
            raise ValueError("weights must be [n_agents][n_actions][V*V*3]")
        self.n_agents, self.n_actions = int(w.shape[0]), int(w.shape[1])
        self.view_size = int(view_size) if view_size is not None else int(round((w.shape[2] // 3) ** 0.5))
        if w.dtype != np.int8:
            if np.any(w != np.clip(np.rint(w), -128, 127)):
                raise ValueError("policy weights must be integers in [-128, 127]")
            w = w.astype(np.int8)
        if w.shape[2] != self.view_size * self.view_size * 3 or not 1 <= self.n_actions <= 7:
            raise ValueError(f"weights must be [A][n_actions <= 7][V*V*3], got {w.shape}")
        self.weights = np.ascontiguousarray(w)
        self.bias = pack_bias(bias, self.n_agents, self.n_actions)[:, : self.n_actions].copy()
        self.epsilon = float(epsilon)
        self.epsilon_u32 = epsilon_to_u32(epsilon)
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._dev = {}

    @classmethod
    def random(cls, n_agents, view_size, n_actions=7, epsilon=0.0, seed=0, rng_seed=0):
        rng = np.random.RandomState(rng_seed)
        w = rng.randint(-128, 128, size=(n_agents, n_actions, view_size * view_size * 3)).astype(np.int8)
        b = rng.randint(-1000, 1000, size=(n_agents, n_actions)).astype(np.int32)
        return cls(w, b, epsilon=epsilon, seed=seed, view_size=view_size)

    def invalidate(self):
        """Call after changing `weights` / `bias` in place: the packed device copies are rebuilt on the next use."""
        self._dev = {}

    def device_struct(self, device):
        """(MgLinearPolicy, keep-alive tensors) with the packed weights on `device` (uploaded on first use per device)."""
        import torch

        key = str(device)
        if key not in self._dev:
            w = torch.from_numpy(pack_weights(self.weights, self.view_size)).to(device)
            b = torch.from_numpy(pack_bias(self.bias, self.n_agents, self.n_actions)).to(device)
            self._dev[key] = (w, b)
        w, b = self._dev[key]
        return MgLinearPolicy(w.data_ptr(), b.data_ptr(), self.n_actions, self.epsilon_u32, self.seed), (w, b)
