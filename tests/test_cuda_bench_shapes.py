"""GPU suite, part 2: the CUDA path against the oracle AT THE SHAPES bench.py TIMES (BASELINE.json configs[2..4]).

The lock-step cases of test_cuda_parity.py stop at 5 000 envs = one tile per CTA.  Here every CTA of the persistent
kernel runs several rounds (input-stage swap / prefetch for encoded observations, the single-stage re-issue path for RGB),
the all-reset step is crossed, and the engine's slices are compared against the oracle on the exact batch of the bench.
Bit-exact bar: observations, float64 reward bit patterns, done, full SoA state (marlgrid/base.py:402-416,501-653).
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

HOST_THREADS = 16


def _state_equal(env, ob, what=""):
    assert np.array_equal(env.grid.cpu().numpy(), ob.grid), f"{what}: grid planes"
    ag = env.agent_rec.cpu().numpy()[:, :, :12].copy()
    ag[:, :, 3] &= 0x7F
    assert np.array_equal(ag, ob.agents[:, :, :12]), f"{what}: agent records"
    assert np.array_equal(env.envrec.cpu().numpy(), ob.envrec), f"{what}: env records"


def _lockstep(env, ob, T, state_every=16, act_seed=1, forward_bias=0.0):
    B, A = env.num_envs, env.num_agents
    obs = env.reset()
    ob.reset()
    assert np.array_equal(obs.cpu().numpy(), ob.obs_encode()), "reset: obs"
    _state_equal(env, ob, "reset")
    rng = np.random.RandomState(act_seed)
    for t in range(T):
        act = rng.randint(0, 7, size=(B, A)).astype(np.int32)
        if forward_bias:
            act[rng.rand(B, A) < forward_bias] = 2
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        o2, r2, d2 = ob.step(act, autoreset=True, with_obs=True)
        assert np.array_equal(obs.cpu().numpy(), o2), f"step {t}: obs"
        assert np.array_equal(rew.cpu().numpy().view(np.uint64), r2.view(np.uint64)), f"step {t}: reward bits"
        assert np.array_equal(done.cpu().numpy(), d2.astype(bool)), f"step {t}: done"
        if t % state_every == 0 or t == T - 1:
            _state_equal(env, ob, f"step {t}")
    assert int(env.err.max().item()) == 0


def test_cfg3_full_batch_lockstep_vs_oracle(oracle):
    """BASELINE configs[2]: 3AgentCluttered15x15 at 65 536 envs (2 048 tiles = 2 rounds of the persistent CTAs), 105 steps:
    crosses the step on which all 65 536 episodes end and are regenerated inside the kernel (base.py:402-416)."""
    from marlgrid_b200 import envs

    B = 65536
    env = envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=B, obs_mode="encoded", seed=1337)
    ob = oracle.OracleBatch(env.cfg, B, seed=1337, threads=HOST_THREADS)
    _lockstep(env, ob, 105)
    assert int(env.episode.min().item()) == 2


def test_cfg3_desynchronised_episodes_vs_oracle(oracle):
    """The same batch with a forward-biased policy and a short horizon: episodes end at irregular times, so most steps find a
    few finished envs in many tiles (warp-cooperative reset route) next to tiles without any."""
    from marlgrid_b200 import envs

    B = 32768
    env = envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=B, obs_mode="encoded", seed=7, max_steps=37)
    ob = oracle.OracleBatch(env.cfg, B, seed=7, threads=HOST_THREADS)
    _lockstep(env, ob, 120, forward_bias=0.5)
    assert int(env.episode.min().item()) >= 3


def test_cfg5_per_gpu_share_lockstep_vs_oracle(oracle):
    """BASELINE configs[4]'s per-GPU share: 131 072 envs (4 rounds per CTA), a short horizon so that the all-reset step falls
    inside the window; global env indices of the last of 8 shards."""
    from marlgrid_b200 import envs

    B = 131072
    env = envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=B, obs_mode="encoded", seed=1337, env_offset=7 * B, max_steps=9)
    ob = oracle.OracleBatch(env.cfg, B, seed=1337, env_offset=7 * B, threads=HOST_THREADS)
    _lockstep(env, ob, 12, state_every=4)


def test_cfg4_rgb_multi_round_vs_oracle(oracle):
    """BASELINE configs[3] family: 4AgentEmpty9x9 RGB at 16 384 envs = 512 tiles on ~444 resident CTAs: the single-stage
    kernel re-issues its input loads after the previous tile has drained (the NST == 1 path)."""
    from marlgrid_b200 import envs
    from marlgrid_b200.atlas import build_atlas

    B = 16384
    env = envs.make("MarlGrid-4AgentEmpty9x9-v0", num_envs=B, obs_mode="rgb", seed=7, max_steps=6)
    ob = oracle.OracleBatch(env.cfg, B, seed=7, threads=HOST_THREADS)
    atlas = build_atlas([int(c) for c in env.cfg.agent_color[:4]], 8)
    obs = env.reset()
    ob.reset()
    assert np.array_equal(obs.cpu().numpy(), ob.obs_rgb(atlas)), "reset: rgb obs"
    rng = np.random.RandomState(2)
    for t in range(8):
        act = rng.randint(0, 7, size=(B, 4)).astype(np.int32)
        act[rng.rand(B, 4) < 0.5] = 2
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        r2, d2 = ob.step(act, autoreset=True)
        assert np.array_equal(rew.cpu().numpy().view(np.uint64), r2.view(np.uint64)), f"step {t}: reward bits"
        assert np.array_equal(done.cpu().numpy(), d2.astype(bool)), f"step {t}: done"
        if t in (0, 5, 7):  # 5 = the all-reset step (max_steps 6)
            assert np.array_equal(obs.cpu().numpy(), ob.obs_rgb(atlas)), f"step {t}: rgb obs"
    _state_equal(env, ob, "end")


def test_cfg4_rgb_full_batch_strided_vs_oracle(oracle):
    """BASELINE configs[3] at its full batch (262 144 envs, 9.9 GB of observations per step, ~19 rounds per CTA): the whole
    state against the oracle, the RGB observations of every 61st env (+ the first and last tile)."""
    from marlgrid_b200 import envs
    from marlgrid_b200.atlas import build_atlas

    B = 262144
    env = envs.make("MarlGrid-4AgentEmpty9x9-v0", num_envs=B, obs_mode="rgb", seed=11, obs_buffers=1)
    ob = oracle.OracleBatch(env.cfg, B, seed=11, threads=HOST_THREADS)
    atlas = build_atlas([int(c) for c in env.cfg.agent_color[:4]], 8)
    idx = np.unique(np.concatenate([np.arange(0, B, 61), np.arange(32), np.arange(B - 32, B)]))
    sub = oracle.OracleBatch(env.cfg, len(idx), seed=11, threads=HOST_THREADS)
    idx_t = torch.from_numpy(idx).cuda()

    def check(obs, what):
        sub.grid[...] = ob.grid[idx]; sub.agents[...] = ob.agents[idx]; sub.envrec[...] = ob.envrec[idx]
        assert np.array_equal(obs[idx_t].cpu().numpy(), sub.obs_rgb(atlas)), f"{what}: rgb obs"

    obs = env.reset()
    ob.reset()
    check(obs, "reset")
    rng = np.random.RandomState(3)
    for t in range(3):
        act = rng.randint(0, 7, size=(B, 4)).astype(np.int32)
        act[rng.rand(B, 4) < 0.5] = 2
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        r2, d2 = ob.step(act, autoreset=True)
        assert np.array_equal(rew.cpu().numpy().view(np.uint64), r2.view(np.uint64)), f"step {t}: reward bits"
        assert np.array_equal(done.cpu().numpy(), d2.astype(bool)), f"step {t}: done"
        check(obs, f"step {t}")
    _state_equal(env, ob, "end")


def test_engine_e2e_full_batch_vs_oracle(cuda_lib, oracle):
    """bench.py's e2e arm on its exact batch: mg_engine_step (host buffers, 4 overlapped env slices) at 65 536 envs."""
    from marlgrid_b200 import _lib
    from marlgrid_b200.config import make_config

    B = 65536
    cfg = make_config(15, 15, ["red", "blue", "purple"], n_clutter=25)
    h = ctypes.c_void_p()
    _lib.check(cuda_lib.mg_engine_create(ctypes.byref(h), ctypes.byref(cfg), B, 3 * B, 1337, 0, 0, None, 0), "mg_engine_create")
    ob = oracle.OracleBatch(cfg, B, seed=1337, env_offset=3 * B, threads=HOST_THREADS)
    obs = np.zeros((B, 3, 7, 7, 3), np.uint8)
    rew = np.zeros((B, 3), np.float64)
    done = np.zeros((B,), np.uint8)
    _lib.check(cuda_lib.mg_engine_reset(h, obs.ctypes.data), "mg_engine_reset")
    ob.reset()
    assert np.array_equal(obs, ob.obs_encode())
    rng = np.random.RandomState(0)
    for t in range(12):
        act = rng.randint(0, 7, size=(B, 3)).astype(np.int32)
        _lib.check(cuda_lib.mg_engine_step(h, act.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data, 1), "mg_engine_step")
        o2, r2, d2 = ob.step(act, autoreset=True, with_obs=True)
        assert np.array_equal(obs, o2) and np.array_equal(rew.view(np.uint64), r2.view(np.uint64)) and np.array_equal(done, d2), f"step {t}"
    cuda_lib.mg_engine_destroy(h)


def test_pregenerated_worlds_equal_in_kernel_generation(cuda_lib):
    """The background world generator (MgState.pregen): a batch whose finished envs copy pre-generated worlds and a batch
    that generates every world inside the step kernel produce identical outputs and state, step by step, in the desynchronised
    regime (a few resets per tile and step) and across all-reset steps; most resets of the first batch are copies."""
    import ctypes as C

    from marlgrid_b200 import envs

    B, T = 8192, 260
    a = envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=B, obs_mode="encoded", seed=5, max_steps=40)
    b = envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=B, obs_mode="encoded", seed=5, max_steps=40, pregen=False)
    a.reset()
    b.reset()
    spread = torch.randint(0, 40, (B // 2,), device="cuda", dtype=torch.int32)
    a.envrec[: B // 2, 0] = spread  # half of the batch desynchronised, the other half in lock step
    b.envrec[: B // 2, 0] = spread
    stats = (C.c_uint64 * 2)()
    cuda_lib.mg_pregen_stats(stats, 1)
    for t in range(T):
        act = a.random_actions(t)
        o1, r1, d1, _ = a.step(act)
        o2, r2, d2, _ = b.step(act)
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2), f"step {t}"
        if t % 20 == 0 or t == T - 1:
            assert torch.equal(a.grid, b.grid) and torch.equal(a.agent_rec[:, :, :12], b.agent_rec[:, :, :12]) and torch.equal(a.envrec, b.envrec)
            assert torch.equal(a.cellbits, b.cellbits)
    cuda_lib.mg_pregen_stats(stats, 0)
    hits, misses = int(stats[0]), int(stats[1])
    assert hits + misses >= B * (T // 40 - 1)
    assert hits > 0.5 * (hits + misses), (hits, misses)
    a.seed(6)  # a new seed invalidates every slot: the same world as a batch that never had any
    b.seed(6)
    a._sync_state_struct(); b._sync_state_struct()
    a._fast = b._fast = None
    for t in range(45):
        act = a.random_actions(t)
        o1, r1, d1, _ = a.step(act)
        o2, r2, d2, _ = b.step(act)
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2), f"after reseeding, step {t}"
    assert torch.equal(a.grid, b.grid) and torch.equal(a.envrec, b.envrec)
