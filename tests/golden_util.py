"""Loading / replaying the golden fixtures recorded from the reference (oracle/gen_golden.py)."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def trajectory_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "traj_*.npz")))


def render_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "render_*.npz")))


def load_traj(path):
    from marlgrid_b200.config import make_config

    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    cfg = make_config(**meta["config"])
    return cfg, meta, z


def rank_from(x, y, placed, stamp):
    """Queue position of each placed agent on its cell (0 = head), -1 if not placed."""
    A = len(x)
    rank = np.full(A, -1, np.int32)
    for a in range(A):
        if not placed[a]:
            continue
        rank[a] = sum(1 for q in range(A) if q != a and placed[q] and x[q] == x[a] and y[q] == y[a] and stamp[q] < stamp[a])
    return rank


def load_los():
    z = np.load(os.path.join(GOLDEN, "los.npz"))
    cases = []
    for key in z.files:
        if not key.startswith("t_"):
            continue
        _, v, ax, ay = key.split("_")
        V, ax, ay = int(v[1:]), int(ax), int(ay)
        n = int(z[f"n_V{V}_{ax}_{ay}"][0])
        t = np.unpackbits(z[key])[: n * V * V].reshape(n, V, V)
        m = np.unpackbits(z[f"m_V{V}_{ax}_{ay}"])[: n * V * V].reshape(n, V, V)
        cases.append((V, ax, ay, t.astype(np.uint8), m.astype(np.uint8)))
    return cases
