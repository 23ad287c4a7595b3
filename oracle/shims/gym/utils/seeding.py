"""gym.utils.seeding.np_random stand-in (marlgrid/base.py:373).

Parity never depends on this stream: every oracle run replaces `env.np_random` with the
Philox contract object (oracle/philox.py) after construction.
"""
import numpy as np


def np_random(seed=None):
    if seed is None:
        seed = 0
    rng = np.random.RandomState(int(seed) % (2 ** 32))
    return rng, seed
