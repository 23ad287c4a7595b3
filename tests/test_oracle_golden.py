"""CPU suite: the C restatement (oracle/) against the golden vectors recorded from the reference.

These run where /root/reference does not exist.  They pin the oracle; the GPU suite then compares
the CUDA path with the oracle and with the same fixtures.
"""
import os

import numpy as np
import pytest

from golden_util import GOLDEN, load_los, load_traj, rank_from, trajectory_files


def test_philox_known_answers(oracle):
    """Random123 known-answer vectors for philox4x32-10 (python statement and C restatement)."""
    from oracle import philox as px

    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kats:
        assert px.philox4x32_10(ctr, key) == want
        assert oracle.philox(ctr, key) == want
    for A in range(1, 9):
        for t in range(50):
            perm = px.shuffle_perm(99, 7 + t, t, A)
            assert sorted(perm) == list(range(A))
            assert perm == oracle.order(99, 7 + t, t, A)


def test_los_golden(oracle):
    """occlude_mask restatement == numba reference results (incl. the A.4 asymmetry hand cases)."""
    n = 0
    for V, ax, ay, t, m in load_los():
        got = oracle.los_batch(t, ax, ay)
        assert np.array_equal(got, m), f"V={V} pos=({ax},{ay})"
        n += len(t)
    assert n >= 6000


def test_los_asymmetry_hand_case(oracle):
    """Column-1 wall hides column 0 except rows 5,6; the mirrored column-5 wall leaves column 6 visible."""
    V = 7
    t = np.ones((2, V, V), np.uint8)
    t[0, 1, : V - 1] = 0
    t[1, V - 2, : V - 1] = 0
    m = oracle.los_batch(t, 3, 6)
    assert m[0, 0].tolist() == [0, 0, 0, 0, 0, 1, 1]
    assert m[1, 6].tolist() == [1, 1, 1, 1, 1, 1, 1]


@pytest.mark.parametrize("ts", [8, 5, 11])
def test_atlas_matches_reference_tiles(ts):
    """Host atlas builder == tiles rendered by the reference's MultiGrid.render_tile."""
    from marlgrid_b200.atlas import build_atlas

    z = np.load(os.path.join(GOLDEN, f"atlas_ts{ts}.npz"))
    got = build_atlas([int(c) for c in z["colors"]], ts)
    assert got.shape == z["atlas"].shape
    assert np.array_equal(got, z["atlas"])
    if ts <= 10:  # equivariance the RGB kernel's dir-remap mode relies on
        A = len(z["colors"])
        per = 1 + 4 * A
        for kind in range(4):
            assert all(np.array_equal(got[kind * per, k], got[kind * per, 0]) for k in range(4))
            for q in range(A):
                for d in range(4):
                    for k in range(4):
                        assert np.array_equal(got[kind * per + 1 + 4 * q + d, k], got[kind * per + 1 + 4 * q + (d + k) % 4, 0])


@pytest.mark.parametrize("path", trajectory_files(), ids=lambda p: os.path.basename(p)[5:-4])
def test_oracle_replays_reference_trajectory(oracle, path):
    """Replay the recorded event stream through the C oracle; everything the reference showed must match."""
    from marlgrid_b200.atlas import build_atlas

    cfg, meta, z = load_traj(path)
    ob = oracle.OracleBatch(cfg, 1, seed=meta["seed"], env_offset=meta["env_index"])
    atlas = build_atlas([int(c) for c in cfg.agent_color[: cfg.n_agents]], cfg.view_tile_size) if "rgb" in z.files else None
    n_rgb = len(z["rgb"]) if atlas is not None else 0
    kinds = z["kind"]
    for i, kind in enumerate(kinds):
        if kind == 0:
            ob.reset()
        elif kind == 2:
            ob.planes()[0][...] = z["grid"][i]
        elif kind == 3:
            ob.step(z["actions"][i][None], autoreset=False)
            assert int(ob.err[0]) & int(z["err"][i])
            ob.envrec[0, 3] &= 0xFFFF
            ob.agents[0, :, 2] = z["dir"][i]
            continue
        else:
            rew, done = ob.step(z["actions"][i][None], autoreset=False)
            assert np.array_equal(rew[0].view(np.uint64), z["rew"][i].view(np.uint64)), f"event {i}: reward bits"
            assert bool(done[0]) == bool(z["done"][i]), f"event {i}: done"
        assert np.array_equal(ob.planes()[0], z["grid"][i]), f"event {i}: planes"
        fl = z["flags"][i]
        placed = (fl & 1).astype(bool)
        assert np.array_equal(ob.agent_flags[0] & 7, fl), f"event {i}: flags"
        assert np.array_equal(ob.agent_x[0][placed], z["x"][i][placed]) and np.array_equal(ob.agent_y[0][placed], z["y"][i][placed])
        assert np.array_equal(ob.agent_dir[0], z["dir"][i])
        assert np.array_equal(ob.agent_carry[0], z["carry"][i])
        assert np.array_equal(rank_from(ob.agent_x[0], ob.agent_y[0], placed, ob.agent_stamp[0]), z["rank"][i]), f"event {i}: queue order"
        assert int(ob.step_count[0]) == int(z["step_count"][i])
        assert np.array_equal(ob.obs_encode()[0], z["enc"][i]), f"event {i}: encoded obs"
        if i < n_rgb:
            assert np.array_equal(ob.obs_rgb(atlas)[0], z["rgb"][i]), f"event {i}: rgb obs"
    assert int(ob.err[0]) == 0


def test_oracle_threads_and_offsets_agree(oracle):
    """Multi-threaded batch == single-threaded; shard [k, k+n) with env_offset == slice of the full batch."""
    from marlgrid_b200.config import make_config

    cfg = make_config(11, 11, ["red", "blue", "purple"], n_clutter=12)
    full = oracle.OracleBatch(cfg, 64, seed=5, threads=1)
    mt = oracle.OracleBatch(cfg, 64, seed=5, threads=4)
    shard = oracle.OracleBatch(cfg, 16, seed=5, env_offset=32)
    rng = np.random.RandomState(0)
    for b in (full, mt, shard):
        b.reset()
    for t in range(120):
        act = rng.randint(0, 7, size=(64, 3)).astype(np.int32)
        o1, r1, d1 = full.step(act, autoreset=True, with_obs=True)
        o2, r2, d2 = mt.step(act, autoreset=True, with_obs=True)
        o3, r3, d3 = shard.step(act[32:48], autoreset=True, with_obs=True)
        assert np.array_equal(o1, o2) and np.array_equal(r1, r2) and np.array_equal(d1, d2)
        assert np.array_equal(o1[32:48], o3) and np.array_equal(r1[32:48], r3) and np.array_equal(d1[32:48], d3)
    assert np.array_equal(full.grid, mt.grid) and np.array_equal(full.agents, mt.agents) and np.array_equal(full.envrec, mt.envrec)
    assert full.envrec[:, 1].min() >= 2  # episodes really rolled over


def test_render_composition_against_reference_frames(oracle):
    """The host-side composition of marlgrid_b200/render.py (tile lookup, highlight, agent-view columns) against the frames
    the reference rendered; its two GPU inputs -- line-of-sight masks and the agents' RGB views -- come from the oracle here."""
    import glob
    import json

    import torch

    from marlgrid_b200 import render as R
    from marlgrid_b200.atlas import build_atlas
    from marlgrid_b200.config import make_config

    class HostEnv:  # what render() reads from an env
        def __init__(self, cfg, ob):
            self.cfg, self.ob, self.obs_mode, self.atlas = cfg, ob, "encoded", None

        planes = property(lambda self: torch.from_numpy(self.ob.planes().copy()))
        agent_rec = property(lambda self: torch.from_numpy(self.ob.agents.copy()))
        prestige = property(lambda self: torch.from_numpy(self.ob.prestige.copy()))

    def views(env, index=0):
        c = env.cfg
        return env.ob.obs_rgb(build_atlas([int(x) for x in c.agent_color[: c.n_agents]], c.view_tile_size, c.n_static_kinds))[index]

    saved = R.visibility_masks, R.agent_views_rgb
    R.visibility_masks, R.agent_views_rgb = (lambda env, index=0: env.ob.vis()[index] != 0), views
    try:
        files = sorted(glob.glob(os.path.join(GOLDEN, "render_*.npz")))
        assert len(files) >= 6
        for path in files:
            z = np.load(path)
            meta = json.loads(bytes(z["meta"]).decode())
            cfg = make_config(**meta["config"])
            ob = oracle.OracleBatch(cfg, 1, seed=meta["seed"], env_offset=meta["env_index"])
            env = HostEnv(cfg, ob)
            for i, kind in enumerate(z["kind"]):
                if kind == 0:
                    ob.reset()
                else:
                    ob.step(z["actions"][i][None].astype(np.int32), autoreset=False)
                if i % 3 == 0 or i == len(z["kind"]) - 1:
                    assert np.array_equal(R.render(env, 0), z["img"][i]), f"{os.path.basename(path)} frame {i}"
    finally:
        R.visibility_masks, R.agent_views_rgb = saved
