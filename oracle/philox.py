"""Philox4x32-10 RNG contract -- host statement (TEST INFRASTRUCTURE ONLY).

The reference draws all env randomness from `self.np_random` at exactly two call sites:
`np_random.shuffle(iter_order)` once per step (marlgrid/base.py:516) and
`np_random.randint(top, bottom)` once per placement try (marlgrid/base.py:699).  The
device replaces gym's MT19937 stream by a counter-based Philox4x32-10 stream, and parity
is defined by feeding the SAME draws to the reference through the duck-typed object below
(SURVEY.md 0.6 / B.3).  This file is the host statement of that contract; the C oracle
(oracle/mg_oracle.c) and the CUDA kernels (marlgrid_b200/csrc/mg_philox.cuh) restate it.

Contract (g = global env index (u64), s = seed (u64), all words u32):
  key      = (lo32(s), hi32(s))
  shuffle  at lifetime step t (number of step() calls made on env g before this one):
             r = philox(ctr=(lo32(g), hi32(g), t, 0), key);  idx = mulhi32(r[0], A!)
             perm = arange(A); for i = A-1 .. 1: j = idx % (i+1); idx //= (i+1); swap(perm[i], perm[j])
  reset placement try k of episode e (e = number of resets of env g before this one):
             r = philox(ctr=(lo32(g), hi32(g), e, 0x80000000 | (k >> 1)), key)
             (x, y) = (mulhi32(r[2*(k&1)], W), mulhi32(r[2*(k&1)+1], H))
  in-step placement try k during lifetime step t (spawn-delay / respawn, base.py:505,643):
             same with ctr word 2 = t and word 3 = 0x40000000 | (k >> 1)
"""
import math

import numpy as np

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

TAG_SHUFFLE = 0x00000000
TAG_RESET = 0x80000000
TAG_INSTEP = 0x40000000


def philox4x32_10(ctr, key):
    """ctr: 4 u32, key: 2 u32 -> 4 u32 (Salmon et al., SC'11; 10 rounds)."""
    c0, c1, c2, c3 = (int(c) & MASK for c in ctr)
    k0, k1 = (int(k) & MASK for k in key)
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return (c0, c1, c2, c3)


def mulhi32(a, b):
    return (int(a) * int(b)) >> 32


def shuffle_perm(seed, g, t, n):
    """The permutation the contract assigns to lifetime step t of env g with n agents."""
    r = philox4x32_10((g & MASK, (g >> 32) & MASK, t & MASK, TAG_SHUFFLE), (seed & MASK, (seed >> 32) & MASK))
    idx = mulhi32(r[0], math.factorial(n))
    perm = list(range(n))
    for i in range(n - 1, 0, -1):
        j = idx % (i + 1)
        idx //= i + 1
        perm[i], perm[j] = perm[j], perm[i]
    return perm


def placement_try(seed, g, c2, tag, k, w, h):
    r = philox4x32_10((g & MASK, (g >> 32) & MASK, c2 & MASK, tag | (k >> 1)), (seed & MASK, (seed >> 32) & MASK))
    o = 2 * (k & 1)
    return mulhi32(r[o], w), mulhi32(r[o + 1], h)


class PhiloxNpRandom:
    """Duck-typed `np_random` for the reference (`.shuffle`, `.randint`), SURVEY.md B.3.

    The harness calls begin_reset() before env.reset() and begin_step() before env.step()
    so the object knows which counter family the next draws belong to.
    """

    def __init__(self, seed, env_index):
        self.seed = int(seed)
        self.g = int(env_index)
        self.episode = 0  # resets performed so far
        self.t = 0  # step() calls performed so far
        self.mode = None
        self.k = 0
        self.c2 = 0
        self.tag = TAG_RESET
        self.log = []

    def begin_reset(self):
        self.mode = "reset"
        self.c2 = self.episode
        self.tag = TAG_RESET
        self.k = 0
        self.episode += 1

    def begin_step(self):
        self.mode = "step"
        self.c2 = self.t
        self.tag = TAG_INSTEP
        self.k = 0
        self._step_t = self.t
        self.t += 1

    # marlgrid/base.py:516
    def shuffle(self, arr):
        assert self.mode == "step"
        perm = shuffle_perm(self.seed, self.g, self._step_t, len(arr))
        vals = [arr[p] for p in perm]
        for i, v in enumerate(vals):
            arr[i] = v
        self.log.append(("shuffle", tuple(perm)))

    # marlgrid/base.py:699
    def randint(self, low, high=None, size=None):
        lo = np.asarray(low)
        hi = np.asarray(high)
        if lo.shape == ():  # scalar draw (doorkey.py `_rand_int`): one try slot, its first word
            x, _ = placement_try(self.seed, self.g, self.c2, self.tag, self.k, int(hi) - int(lo), 1)
            self.k += 1
            return int(lo) + x
        assert lo.shape == (2,)  # place_obj(top, size): pos = randint(top, bottom), base.py:690-699
        x, y = placement_try(self.seed, self.g, self.c2, self.tag, self.k, int(hi[0]) - int(lo[0]), int(hi[1]) - int(lo[1]))
        self.k += 1
        return np.array([int(lo[0]) + x, int(lo[1]) + y])
