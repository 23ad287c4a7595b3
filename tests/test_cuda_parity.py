"""GPU suite: the CUDA path (through the C ABI) against the oracle and the reference's golden vectors.

Bit-exact bar: observations, float64 rewards (bit pattern), done and the full SoA state.
"""
import ctypes
import os

import numpy as np
import pytest

from golden_util import render_files, load_los, load_traj, rank_from, trajectory_files

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _env(cfg, B, **kw):
    from marlgrid_b200.env import BatchedMultiGridEnv

    return BatchedMultiGridEnv(cfg, num_envs=B, device="cuda:0", **kw)


def _state_equal(env, ob, what=""):
    assert np.array_equal(env.grid.cpu().numpy(), ob.grid), f"{what}: grid planes"
    ag = env.agent_rec.cpu().numpy()[:, :, :12].copy()
    ag[:, :, 3] &= 0x7F  # bit 7 of the flags byte is derived state of the device (queue head), not part of the contract
    assert np.array_equal(ag, ob.agents[:, :, :12]), f"{what}: agent records differ at {np.argwhere(ag != ob.agents[:, :, :12])[:4]}"
    assert np.array_equal(env.envrec.cpu().numpy(), ob.envrec), f"{what}: env records"


def test_library_is_cuda_native(cuda_lib):
    assert cuda_lib.mg_version() == 1
    assert b"sm_100a" in cuda_lib.mg_build_info()
    assert torch.cuda.is_available()


def test_los_kernel_matches_golden_and_oracle(cuda_lib, oracle):
    from marlgrid_b200 import _lib

    rng = np.random.RandomState(3)
    cases = load_los()
    for V in (3, 4, 5, 6, 7, 8):
        for ay in (V - 1, V - 2):
            t = (rng.rand(4000, V, V) > rng.rand(4000, 1, 1) * 0.6).astype(np.uint8)
            cases.append((V, V // 2, ay, t, oracle.los_batch(t, V // 2, ay)))
    for V, ax, ay, t, want in cases:
        dt = torch.from_numpy(np.ascontiguousarray(t)).cuda()
        dm = torch.zeros_like(dt)
        _lib.check(cuda_lib.mg_los_batch(dt.data_ptr(), dm.data_ptr(), len(t), V, ax, ay, None), "mg_los_batch")
        torch.cuda.synchronize()
        assert np.array_equal(dm.cpu().numpy(), want), f"V={V} pos=({ax},{ay})"


@pytest.fixture(params=["fused", "general_fused", "two_kernels"])
def step_impl(request, cuda_lib):
    """env.step has three implementations (specialised fused launch for the registered shapes / general fused launch /
    per-env step kernel + observe kernel): test all of them."""
    cuda_lib.mg_debug_force_two_kernels(1 if request.param == "two_kernels" else 0)
    cuda_lib.mg_debug_force_general_fused(1 if request.param == "general_fused" else 0)
    yield request.param
    cuda_lib.mg_debug_force_two_kernels(0)
    cuda_lib.mg_debug_force_general_fused(0)


@pytest.mark.parametrize("path", trajectory_files(), ids=lambda p: os.path.basename(p)[5:-4])
def test_cuda_replays_reference_trajectory(path, step_impl):
    """The event streams recorded from the unmodified reference, replayed through the kernels."""
    cfg, meta, z = load_traj(path)
    has_rgb = "rgb" in z.files
    env = _env(cfg, 1, seed=meta["seed"], env_offset=meta["env_index"], obs_mode="encoded", autoreset=False)
    env_rgb = _env(cfg, 1, seed=meta["seed"], env_offset=meta["env_index"], obs_mode="rgb", autoreset=False) if has_rgb else None
    n_rgb = len(z["rgb"]) if has_rgb else 0
    W, H = cfg.width, cfg.height
    for i, kind in enumerate(z["kind"]):
        act = torch.from_numpy(z["actions"][i][None].astype(np.int32)).cuda()
        if kind == 0:
            obs = env.reset()
            if i < n_rgb:
                rgb = env_rgb.reset()
        elif kind == 2:
            env.planes[0].copy_(torch.from_numpy(z["grid"][i]).cuda())
            env.sync_derived()
            obs = env.observe()
        elif kind == 3:
            env.step(act)
            assert int(env.err[0].item()) & int(z["err"][i])
            env.envrec[:, 3] &= 0xFFFF
            env.agent_rec[0, :, 2] = torch.from_numpy(z["dir"][i].astype(np.uint8)).cuda()
            env.sync_derived()
            continue
        else:
            obs, rew, done, _ = env.step(act)
            assert np.array_equal(rew[0].cpu().numpy().view(np.uint64), z["rew"][i].view(np.uint64)), f"event {i}: reward bits"
            assert bool(done[0].item()) == bool(z["done"][i]), f"event {i}: done"
            if i < n_rgb:
                rgb, rew2, done2, _ = env_rgb.step(act)
                assert torch.equal(rew2, rew) and torch.equal(done2, done)
        assert np.array_equal(obs[0].cpu().numpy(), z["enc"][i]), f"event {i}: encoded obs"
        if i < n_rgb:
            assert np.array_equal(rgb[0].cpu().numpy(), z["rgb"][i]), f"event {i}: rgb obs"
        assert np.array_equal(env.planes[0].cpu().numpy(), z["grid"][i]), f"event {i}: planes"
        ag = env.agent_rec[0].cpu().numpy()
        fl = z["flags"][i]
        placed = (fl & 1).astype(bool)
        assert np.array_equal(ag[:, 3] & 7, fl), f"event {i}: flags"
        assert np.array_equal(ag[placed, 0], z["x"][i][placed]) and np.array_equal(ag[placed, 1], z["y"][i][placed]), f"event {i}: pos"
        assert np.array_equal(ag[:, 2], z["dir"][i]) and np.array_equal(ag[:, 4:7], z["carry"][i])
        stamp = ag[:, 8:12].copy().view(np.int32)[:, 0]
        assert np.array_equal(rank_from(ag[:, 0], ag[:, 1], placed, stamp), z["rank"][i]), f"event {i}: queue order"
        assert int(env.step_count[0].item()) == int(z["step_count"][i])
    assert int(env.err[0].item()) == 0


CONFIGS = {
    "cfg2_3AgentCluttered11x11": ("MarlGrid-3AgentCluttered11x11-v0", 4096, 230),
    "3AgentCluttered15x15_small": ("MarlGrid-3AgentCluttered15x15-v0", 1000, 230),
    "4AgentEmpty9x9": ("MarlGrid-4AgentEmpty9x9-v0", 777, 230),
    "2AgentEmpty9x9_b1": ("MarlGrid-2AgentEmpty9x9-v0", 1, 330),
    "1AgentCluttered_V5": ("MarlGrid-1AgentCluttered15x15-v0", 333, 230),
    "Goalcycle": ("Goalcycle-demo-solo-v0", 300, 230),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_batched_lockstep_vs_oracle(oracle, name, step_impl):
    """Seeded random rollouts with auto-reset: every output and the whole state, every step."""
    from marlgrid_b200 import envs

    env_id, B, T = CONFIGS[name]
    env = envs.make(env_id, num_envs=B, obs_mode="encoded", seed=4242, env_offset=10_000_000_000)
    ob = oracle.OracleBatch(env.cfg, B, seed=4242, env_offset=10_000_000_000, threads=8)
    obs = env.reset()
    ob.reset()
    assert np.array_equal(obs.cpu().numpy(), ob.obs_encode())
    _state_equal(env, ob, "reset")
    rng = np.random.RandomState(1)
    for t in range(T):
        act = rng.randint(0, 7, size=(B, env.num_agents)).astype(np.int32)
        act[rng.rand(B, env.num_agents) < 0.4] = 2
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        o2, r2, d2 = ob.step(act, autoreset=True, with_obs=True)
        assert np.array_equal(obs.cpu().numpy(), o2), f"step {t}: obs"
        assert np.array_equal(rew.cpu().numpy().view(np.uint64), r2.view(np.uint64)), f"step {t}: reward bits"
        assert np.array_equal(done.cpu().numpy(), d2.astype(bool)), f"step {t}: done"
        if t % 16 == 0 or t == T - 1:
            _state_equal(env, ob, f"step {t}")
    assert int(env.episode.min().item()) >= 3
    assert int(env.err.max().item()) == 0


def test_reset_fall_through_routes_vs_oracle(oracle):
    """A world so cluttered that runs of 32 failed placement tries are common: the warp / table resets of the fused kernel hand
    such envs to the sequential code (nothing committed before), which must leave exactly the reference's world (base.py:690-708)."""
    from marlgrid_b200 import envs

    B, T = 512, 70
    kw = dict(num_envs=B, obs_mode="encoded", seed=99, clutter_density=None, n_clutter=66, max_steps=6)
    env = envs.make("MarlGrid-3AgentCluttered11x11-v0", **kw)
    ob = oracle.OracleBatch(env.cfg, B, seed=99, env_offset=0, threads=8)
    obs = env.reset()
    ob.reset()
    assert np.array_equal(obs.cpu().numpy(), ob.obs_encode())
    rng = np.random.RandomState(5)
    for t in range(T):
        act = rng.randint(0, 7, size=(B, env.num_agents)).astype(np.int32)
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        o2, r2, d2 = ob.step(act, autoreset=True, with_obs=True)
        assert np.array_equal(obs.cpu().numpy(), o2), f"step {t}: obs"
        assert np.array_equal(done.cpu().numpy(), d2.astype(bool)), f"step {t}: done"
        _state_equal(env, ob, f"step {t}")


def test_batched_object_interactions_vs_oracle(oracle, step_impl):
    """Keys, balls and doors scattered over full tiles of a batch: the envs whose planes change mid-step (effective
    pickup / drop / toggle, base.py:590-613) take the sequential replay inside the fused kernels, next to envs that do not."""
    from marlgrid_b200.config import make_config

    B, A, T = 200, 3, 220
    cfg = make_config(8, 8, ["red", "blue", "purple"], max_steps=60)
    env = _env(cfg, B, seed=77, obs_mode="encoded")
    ob = oracle.OracleBatch(cfg, B, seed=77, threads=8)
    rng = np.random.RandomState(9)
    objects = [(9, 3, 0), (9, 0, 0), (10, 2, 0), (11, 3, 2), (11, 0, 3), (11, 6, 1), (10, 5, 0)]  # Key blue/red, Ball green, Doors closed/locked/open, Ball purple

    def scatter():  # the same objects on the same free cells of both worlds (about every second env)
        planes = env.planes.cpu().numpy().copy()
        pos = env.agent_pos.cpu().numpy()
        for b in range(B):
            if rng.rand() < 0.5:
                continue
            free = [(x, y) for x in range(1, 7) for y in range(1, 7) if planes[b, 0, x, y] == 0 and not ((pos[b, :, 0] == x) & (pos[b, :, 1] == y)).any()]
            rng.shuffle(free)
            for (t, c, st), (x, y) in zip(objects, free):
                planes[b, :, x, y] = (t, c, st)
        env.planes.copy_(torch.from_numpy(planes).cuda())
        env.sync_derived()
        ob.planes()[...] = planes

    obs = env.reset()
    ob.reset()
    scatter()
    carried = 0
    for t in range(T):
        act = rng.randint(0, 7, size=(B, A)).astype(np.int32)
        act[rng.rand(B, A) < 0.3] = 2
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        o2, r2, d2 = ob.step(act, autoreset=True, with_obs=True)
        assert np.array_equal(obs.cpu().numpy(), o2), f"step {t}: obs"
        assert np.array_equal(rew.cpu().numpy().view(np.uint64), r2.view(np.uint64)), f"step {t}: reward bits"
        assert np.array_equal(done.cpu().numpy(), d2.astype(bool)), f"step {t}: done"
        if t % 10 == 0 or t == T - 1:
            _state_equal(env, ob, f"step {t}")
        carried = max(carried, int((env.agent_carrying[:, :, 0] != 0).sum().item()))
        if d2.any():  # fresh worlds have no objects: scatter again (all envs of the batch end together at max_steps)
            scatter()
    assert carried > 20  # pickups did happen; error bits (e.g. a door closed on an agent, base.py:558) are part of the compared env records


def test_rgb_batched_vs_oracle(oracle, step_impl):
    """RGB tile path (config 4 family) against the oracle fed the same atlas."""
    from marlgrid_b200 import envs
    from marlgrid_b200.atlas import build_atlas

    B = 257
    env = envs.make("MarlGrid-4AgentEmpty9x9-v0", num_envs=B, obs_mode="rgb", seed=7)
    ob = oracle.OracleBatch(env.cfg, B, seed=7, threads=8)
    atlas = build_atlas([int(c) for c in env.cfg.agent_color[:4]], 8)
    obs = env.reset()
    ob.reset()
    assert np.array_equal(obs.cpu().numpy(), ob.obs_rgb(atlas))
    rng = np.random.RandomState(2)
    for t in range(120):
        act = rng.randint(0, 7, size=(B, 4)).astype(np.int32)
        act[rng.rand(B, 4) < 0.5] = 2
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        r2, d2 = ob.step(act, autoreset=True)
        assert np.array_equal(rew.cpu().numpy().view(np.uint64), r2.view(np.uint64))
        assert np.array_equal(done.cpu().numpy(), d2.astype(bool))
        if t % 5 == 0:
            assert np.array_equal(obs.cpu().numpy(), ob.obs_rgb(atlas)), f"step {t}: rgb obs"


@pytest.mark.parametrize("ts,V,off", [(5, 7, 1), (11, 5, 0), (8, 3, 0), (4, 8, 1)])
def test_rgb_generic_tile_sizes(oracle, ts, V, off):
    """Tile sizes with grid lines (ts >= 11: orientation-indexed atlas) and the byte-wise store path."""
    from marlgrid_b200.agents import GridAgentInterface
    from marlgrid_b200.atlas import build_atlas
    from marlgrid_b200.envs import ClutteredMultiGrid

    B = 65
    ags = [GridAgentInterface(color=c, view_size=V, view_tile_size=ts, view_offset=off) for c in ("red", "blue", "pink")]
    env = ClutteredMultiGrid(agents=ags, grid_size=9, n_clutter=6, num_envs=B, obs_mode="rgb", seed=3)
    ob = oracle.OracleBatch(env.cfg, B, seed=3, threads=4)
    atlas = build_atlas([int(c) for c in env.cfg.agent_color[:3]], ts)
    obs = env.reset()
    ob.reset()
    assert np.array_equal(obs.cpu().numpy(), ob.obs_rgb(atlas))
    rng = np.random.RandomState(5)
    for t in range(40):
        act = rng.randint(0, 3, size=(B, 3)).astype(np.int32) + (rng.rand(B, 3) < 0.5)
        act = np.minimum(act, 2).astype(np.int32)
        obs, _, _, _ = env.step(torch.from_numpy(act).cuda())
        ob.step(act, autoreset=True)
        assert np.array_equal(obs.cpu().numpy(), ob.obs_rgb(atlas)), f"step {t}"


def test_split_kernels_equal_fused(oracle):
    """mg_step + mg_obs_encode == mg_step_fused; mg_reset(mask) resets only the masked envs."""
    from marlgrid_b200 import envs

    B = 500
    a = envs.make("MarlGrid-3AgentCluttered11x11-v0", num_envs=B, obs_mode="encoded", seed=11)
    b = envs.make("MarlGrid-3AgentCluttered11x11-v0", num_envs=B, obs_mode="encoded", seed=11)
    a.reset()
    b.reset()
    g = torch.Generator(device="cpu").manual_seed(0)
    for t in range(130):
        act = torch.randint(0, 7, (B, 3), generator=g, dtype=torch.int32).cuda()
        o1, r1, d1, _ = a.step(act)
        r2, d2 = b.step_only(act)
        o2 = b.observe()
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
    mask = (torch.arange(B) % 3 == 0).to(torch.uint8).cuda()
    ep = a.episode.clone()
    grid = a.grid.clone()
    a.reset(mask=mask)
    assert torch.equal(a.episode, ep + mask.int())
    assert torch.equal(a.grid[mask == 0], grid[mask == 0])
    assert bool((a.step_count[mask == 1] == 0).all())


def test_full_size_properties():
    """BASELINE config 3 at its full batch (65 536 envs): size-independent properties."""
    from marlgrid_b200 import envs

    B = 65536
    env = envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=B, obs_mode="encoded", seed=1337)
    half = envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=B // 2, obs_mode="encoded", seed=1337, env_offset=B // 2)
    obs = env.reset()
    obs_h = half.reset()
    assert torch.equal(obs[B // 2:], obs_h)  # sharding invariance: RNG keyed by the global env index
    t = env.grid_type
    assert bool((t[:, 0, :] == 8).all() and (t[:, -1, :] == 8).all() and (t[:, :, 0] == 8).all() and (t[:, :, -1] == 8).all())
    assert bool((t[:, 13, 13] == 4).all())
    assert bool(((t == 8).sum(dim=(1, 2)) == 56 + 25).all())  # border + n_clutter = int(.15*13*13)
    assert bool(env.agent_active.all() and env.agent_placed.all())
    own = obs[:, :, 3, 6, 0]
    assert bool(((own == 13) | (own == 4)).all())  # at its own cell an agent sees an agent (itself / the one below it) or the goal it spawned on
    for step in range(105):
        act = env.random_actions(step)
        ep_before = env.episode.clone()
        obs, rew, done, _ = env.step(act)
        obs_h, rew_h, done_h, _ = half.step(act[B // 2:])
        if step in (0, 50, 99, 100, 104):
            assert torch.equal(obs[B // 2:], obs_h) and torch.equal(rew[B // 2:], rew_h) and torch.equal(done[B // 2:], done_h)
            assert bool(((rew == 0) | ((rew > 0.09) & (rew <= 0.991))).all())
            inactive = ~env.agent_active
            assert bool((obs[inactive] == 0).all())  # inactive agents observe nothing (base.py:420-425)
        if step == 99:
            assert bool(done[ep_before == 1].all())  # max_steps = 100: a first episode still running ends here
            assert float(done.float().mean().item()) > 0.9
    assert int(env.episode.min().item()) == 2 and int(env.err.max().item()) == 0
    assert bool((env.step_count[env.episode == 2] <= 105).all())


def test_error_bits_mirror_reference_exceptions():
    from marlgrid_b200 import envs

    env = envs.make("MarlGrid-2AgentEmpty9x9-v0", num_envs=40, obs_mode="encoded")
    env.reset()
    act = torch.zeros((40, 2), dtype=torch.int32, device="cuda")
    act[7, 1] = 9
    env.step(act)
    assert int(env.err[7].item()) == 1 and int(env.err.sum().item()) == 1
    with pytest.raises(ValueError):
        env.check_errors()
    assert int(env.err.sum().item()) == 0


@pytest.mark.parametrize("B", [3000, 5000])
def test_engine_host_buffer_api(cuda_lib, oracle, B):
    """mg_engine_*: host buffers in, host buffers out (the e2e path of bench.py); 5000 envs = four overlapped slices."""
    from marlgrid_b200 import _lib
    from marlgrid_b200.config import make_config

    cfg = make_config(15, 15, ["red", "blue", "purple"], n_clutter=25)
    h = ctypes.c_void_p()
    _lib.check(cuda_lib.mg_engine_create(ctypes.byref(h), ctypes.byref(cfg), B, 0, 1337, 0, 0, None, 0), "mg_engine_create")
    ob = oracle.OracleBatch(cfg, B, seed=1337, threads=8)
    obs = np.zeros((B, 3, 7, 7, 3), np.uint8)
    rew = np.zeros((B, 3), np.float64)
    done = np.zeros((B,), np.uint8)
    _lib.check(cuda_lib.mg_engine_reset(h, obs.ctypes.data), "mg_engine_reset")
    ob.reset()
    assert np.array_equal(obs, ob.obs_encode())
    rng = np.random.RandomState(0)
    for t in range(110):
        act = rng.randint(0, 7, size=(B, 3)).astype(np.int32)
        _lib.check(cuda_lib.mg_engine_step(h, act.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data, 1), "mg_engine_step")
        o2, r2, d2 = ob.step(act, autoreset=True, with_obs=True)
        assert np.array_equal(obs, o2) and np.array_equal(rew.view(np.uint64), r2.view(np.uint64)) and np.array_equal(done, d2)
    cuda_lib.mg_engine_destroy(h)


def test_round_robin_rollout_equals_separate_families(cuda_lib):
    """mg_rollout_fused_rr (bench.py's timed loop): R families stepped round robin from C == each family stepped alone."""
    from marlgrid_b200 import _lib, envs
    from marlgrid_b200.config import MgState

    R, B, T = 3, 2048, 36
    fams = [envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=B, obs_mode="encoded", seed=11, env_offset=r * B) for r in range(R)]
    refs = [envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=B, obs_mode="encoded", seed=11, env_offset=r * B) for r in range(R)]
    for e in fams + refs:
        e.reset()
    actions = torch.stack([fams[0].random_actions(t) for t in range(T)])
    states = (MgState * R)(*[f._state for f in fams])
    PP = ctypes.c_void_p * R
    _lib.check(cuda_lib.mg_rollout_fused_rr(ctypes.byref(fams[0].cfg), states, R, actions.data_ptr(), T,
                                            PP(*[f.rewards.data_ptr() for f in fams]), PP(*[f.done.data_ptr() for f in fams]),
                                            PP(*[f.obs.data_ptr() for f in fams]), 1, None), "mg_rollout_fused_rr")
    for t in range(T):
        refs[t % R].step(actions[t])
    torch.cuda.synchronize()
    for f, g in zip(fams, refs):
        assert torch.equal(f.obs, g.obs) and torch.equal(f.rewards, g.rewards) and torch.equal(f.done, g.done)
        assert torch.equal(f.grid, g.grid) and torch.equal(f.agent_rec, g.agent_rec) and torch.equal(f.envrec, g.envrec)


def test_checkpoint_resume_is_exact():
    from marlgrid_b200 import envs

    a = envs.make("MarlGrid-3AgentCluttered11x11-v0", num_envs=300, obs_mode="encoded", seed=5)
    a.reset()
    for t in range(37):
        a.step(a.random_actions(t))
    sd = a.state_dict()
    b = envs.make("MarlGrid-3AgentCluttered11x11-v0", num_envs=300, obs_mode="encoded", seed=999)
    b.load_state_dict(sd)
    for t in range(37, 160):
        act = a.random_actions(t)
        o1, r1, d1, _ = a.step(act)
        o2, r2, d2, _ = b.step(act)
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)


def test_unbatched_gym_surface():
    """The reference's single-env surface: lists of per-agent obs, float64 reward array, python bool done."""
    from marlgrid_b200 import envs

    env = envs.make("MarlGrid-2AgentEmpty9x9-v0").unbatched()
    obs_list = env.reset()
    assert len(obs_list) == 2 and tuple(obs_list[0].shape) == (56, 56, 3) and obs_list[0].dtype == torch.uint8
    obs_list, rew, done, info = env.step([2, 0])
    assert tuple(rew.shape) == (2,) and rew.dtype == torch.float64 and isinstance(done, bool) and info == {}
    with pytest.raises(ValueError):
        env.step([2, 17])
    with pytest.raises(AssertionError):
        env.step([2])


def test_rich_observation_style_matches_reference():
    """observation_style='rich' (marlgrid/base.py:461-471): the dict of batched tensors against the dicts the reference returned
    (tests/golden/rich_*.npz, recorded by oracle/gen_golden.py gen_rich), event by event."""
    import json

    from marlgrid_b200.agents import GridAgentInterface
    from marlgrid_b200.envs import EmptyMultiGrid

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "rich_Empty7x7x3.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    c = meta["config"]
    from marlgrid_b200.objects import IDX_TO_COLOR

    ags = [GridAgentInterface(color=IDX_TO_COLOR[ci], view_size=c["view_size"], view_tile_size=c["view_tile_size"], observation_style="rich",
                              observe_rewards=True, observe_position=True, observe_orientation=True) for ci in c["agent_colors"]]
    env = EmptyMultiGrid(agents=ags, grid_size=c["width"], max_steps=c["max_steps"], num_envs=1, obs_mode="rgb", seed=meta["seed"],
                         env_offset=meta["env_index"], autoreset=False)
    for i, kind in enumerate(z["kind"]):
        if kind == 0:
            obs = env.reset()
        else:
            obs, rew, done, _ = env.step(torch.from_numpy(z["actions"][i][None].astype(np.int32)).cuda())
        assert set(obs) == {"pov", "reward", "position", "orientation"}
        assert np.array_equal(obs["pov"][0].cpu().numpy(), z["pov"][i]), f"event {i}: pov"
        assert np.array_equal(obs["position"][0].cpu().numpy().view(np.uint64), z["position"][i].view(np.uint64)), f"event {i}: position bits"
        assert np.array_equal(obs["orientation"][0].cpu().numpy(), z["orientation"][i]), f"event {i}: orientation"
        assert np.array_equal(obs["reward"][0].cpu().numpy().astype(np.float64), z["reward"][i]), f"event {i}: reward"
    one = EmptyMultiGrid(agents=[a.clone() for a in ags], grid_size=7, num_envs=1, obs_mode="rgb", seed=3).unbatched()
    lst = one.reset()
    assert isinstance(lst, list) and len(lst) == 3 and set(lst[0]) == {"pov", "reward", "position", "orientation"}
    assert tuple(lst[0]["pov"].shape) == (56, 56, 3) and tuple(lst[1]["position"].shape) == (2,)


@pytest.mark.parametrize("path", render_files(), ids=lambda p: os.path.basename(p)[7:-4])
def test_render_matches_reference_frames(path):
    """env.render(): the reference's whole-grid view env.render(mode='rgb_array') (base.py:714-795), frame by frame."""
    cfg, meta, z = load_traj(path)
    env = _env(cfg, 1, seed=meta["seed"], env_offset=meta["env_index"], obs_mode="encoded", autoreset=False)
    for i, kind in enumerate(z["kind"]):
        if kind == 0:
            env.reset()
        else:
            env.step(torch.from_numpy(z["actions"][i][None].astype(np.int32)).cuda())
        img = env.render(0)
        assert img.dtype == np.uint8 and img.shape == z["img"][i].shape and np.array_equal(img, z["img"][i]), f"frame {i}"
    plain = env.render(0, highlight=False, show_agent_views=False)
    assert plain.shape == (cfg.height * 32, cfg.width * 32, 3)


def test_grid_recorder(tmp_path):
    """GridRecorder (marlgrid/utils/video.py:55-154): frames of one env of a batch, exported at the next reset."""
    from PIL import Image

    from marlgrid_b200 import envs
    from marlgrid_b200.utils.video import GridRecorder

    env = GridRecorder(envs.make("MarlGrid-3AgentCluttered11x11-v0", num_envs=8, obs_mode="encoded", autoreset=False),
                       save_root=str(tmp_path), max_steps=20, index=3)
    env.recording = True
    env.reset()
    first = env.env.render(index=3)
    for t in range(6):
        env.step(env.random_actions(t))
    assert env.ptr == 6 and np.array_equal(env.frames[0], first)
    env.reset()
    frames_dir = os.path.join(str(tmp_path), "frames_8")
    assert sorted(os.listdir(frames_dir)) == sorted(f"frame_{k}.png" for k in range(7))
    assert np.array_equal(np.asarray(Image.open(os.path.join(frames_dir, "frame_0.png"))), first)
    assert any(f.startswith("video_8.") for f in os.listdir(str(tmp_path)))


def test_previous_observation_survives_the_next_step():
    """obs_buffers=2 (default): `save_step(obs, act, next_obs, ...)` of the reference's loop (README.md:43-57) sees two different
    observations without copying; obs_buffers=1 returns the same storage every step."""
    from marlgrid_b200 import envs

    env = envs.make("MarlGrid-3AgentCluttered11x11-v0", num_envs=64, obs_mode="encoded", seed=2)
    obs = env.reset()
    for t in range(12):
        keep = obs.clone()
        next_obs, rew, done, _ = env.step(env.random_actions(t))
        assert next_obs.data_ptr() != obs.data_ptr() and torch.equal(obs, keep)
        obs = next_obs
    one = envs.make("MarlGrid-3AgentCluttered11x11-v0", num_envs=64, obs_mode="encoded", seed=2, obs_buffers=1)
    o0 = one.reset()
    o1, _, _, _ = one.step(one.random_actions(0))
    assert o1.data_ptr() == o0.data_ptr()


@pytest.mark.parametrize("env_id,B", [("MarlGrid-3AgentCluttered15x15-v0", 4096), ("MarlGrid-2AgentEmpty9x9-v0", 1008), ("MarlGrid-4AgentEmpty9x9-v0", 65536 + 32),
                                       ("MarlGrid-3AgentCluttered11x11-v0", 1001)])  # (1001: per-step slices off the 16-byte grid -> launch per step through a scratch buffer)
def test_persistent_rollout_equals_step_by_step(env_id, B):
    """mg_rollout_persistent: T steps in one launch (state resident in shared memory) == T calls of env.step, every step's
    outputs compared; the last case does not fit the resident CTAs and takes the step-by-step fallback; resets included."""
    from marlgrid_b200 import envs

    T = 108
    a = envs.make(env_id, num_envs=B, obs_mode="encoded", seed=21)
    b = envs.make(env_id, num_envs=B, obs_mode="encoded", seed=21)
    a.reset()
    b.reset()
    actions = torch.stack([a.random_actions(t) for t in range(T)])
    obs, rew, done = a.rollout_all(actions)
    for t in range(T):
        o, r, d, _ = b.step(actions[t])
        assert torch.equal(obs[t], o) and torch.equal(rew[t], r) and torch.equal(done[t], d), f"step {t}"
    assert torch.equal(a.grid, b.grid) and torch.equal(a.envrec, b.envrec) and torch.equal(a.cellbits, b.cellbits)
    assert torch.equal(a.agent_rec[:, :, :12], b.agent_rec[:, :, :12]) and int(a.episode.min().item()) >= 2


@pytest.mark.parametrize("env_id,B,eps,impl", [
    ("MarlGrid-3AgentCluttered15x15-v0", 4096, 0.0, "fused"),
    ("MarlGrid-3AgentCluttered15x15-v0", 1000, 0.25, "fused"),       # ragged last tile, exploration draws
    ("MarlGrid-3AgentCluttered11x11-v0", 777, 0.1, "fused"),
    ("MarlGrid-2AgentEmpty9x9-v0", 333, 0.5, "fused"),
    ("MarlGrid-3AgentCluttered15x15-v0", 300, 0.25, "general_fused"),  # step launch + policy launch per step
    ("MarlGrid-3AgentCluttered15x15-v0", 300, 0.25, "two_kernels"),
    ("MarlGrid-4AgentEmpty9x9-v0", 65536 + 32, 0.05, "fused"),        # more tiles than resident CTAs: the per-step fallback
])
def test_policy_rollout_closed_loop_vs_oracle(cuda_lib, oracle, env_id, B, eps, impl):
    """mg_rollout_policy: step t + 1 plays what the on-device policy chose from step t's observations -- compared step by step
    (observations, reward bits, done, the actions played, final state) with the oracle closing the same loop on the CPU."""
    from marlgrid_b200 import envs
    from marlgrid_b200.policy import LinearPolicy
    from oracle import policy_oracle

    cuda_lib.mg_debug_force_two_kernels(1 if impl == "two_kernels" else 0)
    cuda_lib.mg_debug_force_general_fused(1 if impl == "general_fused" else 0)
    try:
        T = 40 if B > 5000 else 115
        env = envs.make(env_id, num_envs=B, obs_mode="encoded", seed=77, env_offset=5)
        env.reset()
        A, V = env.cfg.n_agents, env.cfg.view_size
        pol = LinearPolicy.random(A, V, n_actions=7, epsilon=eps, seed=0xABCDEF0123, rng_seed=B)
        ob = oracle.OracleBatch(env.cfg, B, seed=77, env_offset=5, threads=8)
        ob.reset()
        first = np.random.RandomState(B).randint(0, 7, size=(B, A)).astype(np.int32)
        obs, rew, done, act = env.rollout_policy(pol, first, T)
        torch.cuda.synchronize()
        o2, r2, d2, a2 = policy_oracle.closed_loop(ob, pol, first, T)
        act, obs, rew, done = act.cpu().numpy(), obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        for t in range(T):
            assert np.array_equal(act[t], a2[t]), f"step {t}: actions differ at {np.argwhere(act[t] != a2[t])[:4]}"
            assert np.array_equal(obs[t], o2[t]), f"step {t}: obs"
            assert np.array_equal(rew[t].view(np.uint64), r2[t].view(np.uint64)), f"step {t}: rewards"
            assert np.array_equal(done[t], d2[t].astype(bool)), f"step {t}: done"
        _state_equal(env, ob, "after the closed loop")
        assert len(np.unique(act[1:])) >= 3  # (the policy is not degenerate)
        if T > 100:
            assert int(env.episode.min().item()) >= 2  # crossed the auto-reset
    finally:
        cuda_lib.mg_debug_force_two_kernels(0)
        cuda_lib.mg_debug_force_general_fused(0)


def test_policy_act_host_loop_vs_oracle(cuda_lib, oracle):
    """mg_policy_act between two env.step calls (the reference's README loop with the policy as a kernel) == the CPU statement."""
    from marlgrid_b200 import envs
    from marlgrid_b200.policy import LinearPolicy
    from oracle import policy_oracle

    B = 515
    env = envs.make("MarlGrid-3AgentCluttered11x11-v0", num_envs=B, obs_mode="encoded", seed=3)
    obs = env.reset()
    pol = LinearPolicy.random(3, 7, n_actions=3, epsilon=0.3, seed=11, rng_seed=1)  # restrict_actions-style: 3 actions
    g = np.arange(B)
    for t in range(30):
        act = env.policy_act(pol, obs)
        want = policy_oracle.linear_policy_actions(obs.cpu().numpy(), pol.weights, pol.bias, 3, pol.epsilon_u32, pol.seed, g, env.envrec[:, 2].cpu().numpy())
        assert np.array_equal(act.cpu().numpy(), want), f"step {t}"
        assert int(act.max()) <= 2
        obs, _, _, _ = env.step(act)


def test_linear_q_trainer_closes_the_loop_on_the_device():
    """LinearQTrainer: rollouts by the quantised policy inside the kernel, TD(0) updates in torch.  The rollout it learns from is
    the one the policy really played (actions recomputed from the observations), and within a few dozen iterations the reward
    rate of the goal-seeking task rises above the untrained policy's."""
    from marlgrid_b200 import envs
    from marlgrid_b200.learners import LinearQLearner, LinearQTrainer, quantized_policy
    from oracle import policy_oracle

    B = 4096
    env = envs.make("MarlGrid-2AgentEmpty9x9-v0", num_envs=B, obs_mode="encoded", seed=5)
    env.reset()
    learners = [LinearQLearner(view_size=7, device=env.device, seed=k, color=c) for k, c in enumerate(("red", "blue"))]
    tr = LinearQTrainer(env, learners, horizon=32, epsilon=0.15, seed=3)
    pol0 = quantized_policy(learners, 0.15, seed=3)
    first = tr.iterate()
    obs, rew, done, act = (x.cpu().numpy() for x in tr.out)
    life = env.envrec[:, 2].cpu().numpy().astype(np.int64)  # lifetime steps after the rollout: step t's observation was made at life - (T - 1 - t)
    for t in (0, 7, 30):
        want = policy_oracle.linear_policy_actions(obs[t], pol0.weights, pol0.bias, 7, pol0.epsilon_u32, pol0.seed, np.arange(B), life - (31 - t))
        assert np.array_equal(act[t + 1], want), f"actions of step {t + 1} are not the policy's choice from step {t}'s observations"
    rates = [first["reward_per_env_step"]] + [tr.iterate()["reward_per_env_step"] for _ in range(59)]
    assert all(np.isfinite(r) for r in rates)
    best = max(np.mean(rates[i: i + 10]) for i in range(10, 51))  # (TD(0) on a linear model is not monotone: best later window)
    assert best > 1.5 * np.mean(rates[:5]), f"reward rate did not improve: {np.mean(rates[:5]):.5f} -> best window {best:.5f}"


@pytest.mark.parametrize("hidden", [["Wall"], ["Goal"], ["Agent"], ["Goal", "Agent"], ["Wall", "Agent", "Key", "Door"]])
def test_hide_item_types_lockstep_vs_oracle(oracle, hidden, step_impl):
    """hide_item_types (agents.py:30, base.py:441-449) in a crowded world -- stacked agents on goals and on each other, keys and
    doors in view -- in lock step with the oracle: the mask variant of the encoded observe (general fused kernel, bit-plane
    worlds) and the cell-by-cell variant (observe kernel) must both follow the reference's replace-once rule."""
    from marlgrid_b200.config import make_config
    from marlgrid_b200.objects import hide_mask

    B, T = 600, 130
    cfg = make_config(9, 9, ["red", "blue", "purple", "orange"], n_clutter=6, max_steps=40, hide_types=hide_mask(hidden))
    env = _env(cfg, B, seed=31, env_offset=77)
    ob = oracle.OracleBatch(cfg, B, seed=31, env_offset=77, threads=8)
    env.reset()
    ob.reset()
    for w in (env.planes, ob.planes()):  # a key and a closed door in every world (objects beyond the generators')
        w[:, 0, 3, 4], w[:, 1, 3, 4], w[:, 2, 3, 4] = 9, 3, 0
        w[:, 0, 5, 2], w[:, 1, 5, 2], w[:, 2, 5, 2] = 11, 2, 1
    env.sync_derived()
    assert np.array_equal(env.observe().cpu().numpy(), ob.obs_encode())
    rng = np.random.RandomState(5)
    for t in range(T):
        act = rng.randint(0, 3, size=(B, 4)).astype(np.int32)  # left / right / forward: the planes stay as edited
        act[rng.rand(B, 4) < 0.5] = 2
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        o2, r2, d2 = ob.step(act, autoreset=True, with_obs=True)  # (the edited objects live until the first reset, step 40)
        assert np.array_equal(obs.cpu().numpy(), o2), f"step {t}: obs differ at {np.argwhere(obs.cpu().numpy() != o2)[:3]}"
        assert np.array_equal(rew.cpu().numpy().view(np.uint64), r2.view(np.uint64)) and np.array_equal(done.cpu().numpy(), d2.astype(bool)), f"step {t}"
    assert len(np.unique(o2[..., 0])) >= 3


def test_respawn_lockstep_vs_oracle(oracle, step_impl):
    """respawn=True (base.py:626-644) with many goal hits: an agent that finishes is placed anew inside the step.  The specialised
    kernel replays such envs with the sequential code; every output and the whole state against the oracle, every step."""
    from marlgrid_b200.config import make_config

    B, T = 1500, 230
    cfg = make_config(9, 9, ["red", "blue", "purple", "orange"], respawn=True, max_steps=100)
    env = _env(cfg, B, seed=9, env_offset=3)
    ob = oracle.OracleBatch(cfg, B, seed=9, env_offset=3, threads=8)
    env.reset()
    ob.reset()
    rng = np.random.RandomState(2)
    hits = 0
    for t in range(T):
        act = rng.randint(0, 7, size=(B, 4)).astype(np.int32)
        act[rng.rand(B, 4) < 0.6] = 2
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        o2, r2, d2 = ob.step(act, autoreset=True, with_obs=True)
        assert np.array_equal(obs.cpu().numpy(), o2), f"step {t}: obs"
        assert np.array_equal(rew.cpu().numpy().view(np.uint64), r2.view(np.uint64)), f"step {t}: reward bits"
        assert np.array_equal(done.cpu().numpy(), d2.astype(bool)), f"step {t}: done"
        hits += int((r2 > 0).sum())
        if t % 16 == 0 or t == T - 1:
            _state_equal(env, ob, f"step {t}")
    assert hits > 500 and int(env.episode.min().item()) >= 3 and int(env.err.max().item()) == 0


def test_reference_readme_loop_with_torch_learners():
    """The training loop of the reference's README.md:29-57, verbatim, with torch learners (LinearQLearner) as the env's agents:
    IndependentLearners.action_step / save_step / episode() on device tensors, one TD update per learner at the end of an episode."""
    from marlgrid_b200 import IndependentLearners, envs
    from marlgrid_b200.learners import LinearQLearner

    agents = IndependentLearners(*[LinearQLearner(view_size=7, device="cuda", seed=k, color=c, epsilon=0.3) for k, c in enumerate(("red", "blue"))])
    env = envs.ClutteredMultiGrid(agents, grid_size=15, n_clutter=10, obs_mode="encoded", max_steps=40)
    w_before = [l.W.detach().clone() for l in agents]
    for i_episode in range(3):
        obs_array = env.reset()
        with agents.episode():
            episode_over = False
            steps = 0
            while not episode_over:
                action_array = agents.action_step(obs_array)
                next_obs_array, reward_array, done, _ = env.step(action_array)
                agents.save_step(obs_array, action_array, next_obs_array.clone(), reward_array.clone(), done.clone())
                obs_array = next_obs_array.clone()
                episode_over = done
                steps += 1
            assert steps == 40 or bool(done)
    assert all(l.updates == 3 and l.buffer == [] for l in agents)
    assert all(not torch.equal(w, l.W.detach()) for w, l in zip(w_before, agents))


def test_human_player_config_lockstep_vs_oracle(oracle, step_impl):
    """The reference's only example config (examples/human_player.py:33-55: ClutteredGoalCycleEnv 13x13, respawn=True, no reward
    decay, 3 bonus tiles, penalty -1.5, view_offset 1), encoded observations, three agents: with no Goal / Lava in the world a
    respawn never happens, and the specialised kernel takes the config; in lock step with the oracle."""
    from marlgrid_b200 import envs
    from marlgrid_b200.agents import GridAgentInterface

    B, T = 1200, 270
    agents = [GridAgentInterface(view_size=7, view_offset=1, view_tile_size=11, see_through_walls=False, color=c) for c in ("red", "blue", "green")]
    env = envs.ClutteredGoalCycleEnv(agents=agents, grid_size=13, max_steps=250, clutter_density=0.15, respawn=True, ghost_mode=True, reward_decay=False,
                                     n_bonus_tiles=3, initial_reward=True, penalty=-1.5, num_envs=B, obs_mode="encoded", seed=17, env_offset=1000)
    ob = oracle.OracleBatch(env.cfg, B, seed=17, env_offset=1000, threads=8)
    env.reset()
    ob.reset()
    rng = np.random.RandomState(3)
    bonus = 0
    for t in range(T):
        act = rng.randint(0, 7, size=(B, 3)).astype(np.int32)
        act[rng.rand(B, 3) < 0.5] = 2
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        o2, r2, d2 = ob.step(act, autoreset=True, with_obs=True)
        assert np.array_equal(obs.cpu().numpy(), o2), f"step {t}: obs"
        assert np.array_equal(rew.cpu().numpy().view(np.uint64), r2.view(np.uint64)), f"step {t}: reward bits"
        assert np.array_equal(done.cpu().numpy(), d2.astype(bool)), f"step {t}: done"
        bonus += int((r2 != 0).sum())
        if t % 32 == 0 or t == T - 1:
            _state_equal(env, ob, f"step {t}")
    assert bonus > 1000 and int(env.episode.min().item()) >= 2 and int(env.err.max().item()) == 0


def test_policy_rollout_other_agent_counts_and_view_sizes(oracle):
    """tools/policy_shapes_check.py: the one-launch closed-loop route for A = 1 / view size 5, A = 4 / view size 7 and a batch that
    is a multiple of 16 but not of 32 envs, against the CPU statement (the MMA fragment layout is generic in A and V)."""
    import runpy

    runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "policy_shapes_check.py"), run_name="__main__")
