"""Long lock-step runs of the CUDA path against the oracle (tools/gpu_check.sh soak): thousands of steps with auto-reset,
forward-biased random actions (goal hits, stacking, irregular episode ends), every output compared every step."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marlgrid_b200 import envs  # noqa: E402
from marlgrid_b200.atlas import build_atlas  # noqa: E402
from oracle import mg_oracle  # noqa: E402

mg_oracle.build()


def soak(env_id, B, T, mode, seed, p_forward):
    env = envs.make(env_id, num_envs=B, obs_mode=mode, seed=seed, env_offset=123456789)
    ob = mg_oracle.OracleBatch(env.cfg, B, seed=seed, env_offset=123456789, threads=16)
    A = env.num_agents
    atlas = build_atlas([int(c) for c in env.cfg.agent_color[:A]], env.cfg.view_tile_size) if mode == "rgb" else None
    env.reset()
    ob.reset()
    rng = np.random.RandomState(seed)
    t0 = time.time()
    hits = 0
    for t in range(T):
        act = rng.randint(0, 7, size=(B, A)).astype(np.int32)
        act[rng.rand(B, A) < p_forward] = 2
        obs, rew, done, _ = env.step(torch.from_numpy(act).cuda())
        if mode == "rgb":
            r2, d2 = ob.step(act, autoreset=True)
            if t % 10 == 0:
                assert np.array_equal(obs.cpu().numpy(), ob.obs_rgb(atlas)), f"{env_id} step {t}: rgb obs"
        else:
            o2, r2, d2 = ob.step(act, autoreset=True, with_obs=True)
            assert np.array_equal(obs.cpu().numpy(), o2), f"{env_id} step {t}: obs"
        assert np.array_equal(rew.cpu().numpy().view(np.uint64), r2.view(np.uint64)), f"{env_id} step {t}: reward bits"
        assert np.array_equal(done.cpu().numpy(), d2.astype(bool)), f"{env_id} step {t}: done"
        hits += int((r2 > 0).sum())
        if t % 100 == 0:
            assert np.array_equal(env.grid.cpu().numpy(), ob.grid) and np.array_equal(env.envrec.cpu().numpy(), ob.envrec), f"{env_id} step {t}: state"
    assert int(env.err.max().item()) == 0
    print(f"{env_id:36s} {mode:8s} B={B:6d} T={T:5d}: ok  ({B * T} env-steps, {hits} goal rewards, episodes {int(env.episode.min())}..{int(env.episode.max())}, {time.time() - t0:.0f} s)")


def soak_policy(env_id, B, chunks, T, seed, eps):
    """Closed loop (mg_rollout_policy, T steps per launch) against the CPU statement closing the same loop on the oracle."""
    from marlgrid_b200.policy import LinearPolicy
    from oracle import policy_oracle

    env = envs.make(env_id, num_envs=B, obs_mode="encoded", seed=seed, env_offset=987654321)
    ob = mg_oracle.OracleBatch(env.cfg, B, seed=seed, env_offset=987654321, threads=16)
    env.reset()
    ob.reset()
    A, V = env.num_agents, env.cfg.view_size
    t0 = time.time()
    hits = 0
    for c in range(chunks):
        pol = LinearPolicy.random(A, V, n_actions=7, epsilon=eps, seed=seed * 1000 + c, rng_seed=seed * 1000 + c)  # a new policy every launch
        pol.weights[:, 2] += 40  # bias towards `forward`: goal hits, stacking, irregular episode ends
        pol.weights = np.clip(pol.weights, -128, 127).astype(np.int8)
        first = np.random.RandomState(c).randint(0, 7, size=(B, A)).astype(np.int32)
        obs, rew, done, act = (x.cpu().numpy() for x in env.rollout_policy(pol, first, T))
        o2, r2, d2, a2 = policy_oracle.closed_loop(ob, pol, first, T)
        assert np.array_equal(act, a2), f"{env_id} launch {c}: actions"
        assert np.array_equal(obs, o2), f"{env_id} launch {c}: obs"
        assert np.array_equal(rew.view(np.uint64), r2.view(np.uint64)) and np.array_equal(done, d2.astype(bool)), f"{env_id} launch {c}: rewards / done"
        assert np.array_equal(env.grid.cpu().numpy(), ob.grid) and np.array_equal(env.envrec.cpu().numpy(), ob.envrec), f"{env_id} launch {c}: state"
        hits += int((r2 > 0).sum())
    print(f"{env_id:36s} policy   B={B:6d} T={chunks * T:5d}: ok  ({B * chunks * T} env-steps closed loop, {hits} goal rewards, episodes {int(env.episode.min())}..{int(env.episode.max())}, {time.time() - t0:.0f} s)")


if __name__ == "__main__":
    soak_policy("MarlGrid-3AgentCluttered15x15-v0", 4096, 15, 100, 21, 0.2)
    soak_policy("MarlGrid-2AgentEmpty9x9-v0", 2048, 15, 100, 22, 0.1)
    soak("MarlGrid-3AgentCluttered15x15-v0", 4096, 3000, "encoded", 11, 0.5)
    soak("MarlGrid-3AgentCluttered11x11-v0", 2048, 3000, "encoded", 12, 0.6)
    soak("MarlGrid-4AgentEmpty9x9-v0", 512, 1500, "rgb", 13, 0.6)
    soak("MarlGrid-2AgentEmpty9x9-v0", 1024, 3000, "encoded", 14, 0.7)
    soak("Goalcycle-demo-solo-v0", 1024, 2000, "encoded", 15, 0.6)
