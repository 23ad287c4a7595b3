"""One-launch closed-loop rollout (mg_rollout_policy) on the agent counts / view sizes the GPU suite does not reach on that route:
A = 1 with view size 5, A = 4 with view size 7, against the CPU statement of the same loop."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from marlgrid_b200 import envs  # noqa: E402
from marlgrid_b200.policy import LinearPolicy  # noqa: E402
from oracle import mg_oracle, policy_oracle  # noqa: E402

for env_id, B, T in (("MarlGrid-1AgentCluttered15x15-v0", 2048, 30), ("MarlGrid-4AgentEmpty9x9-v0", 4096, 30), ("MarlGrid-2AgentEmpty9x9-v0", 992, 30)):
    env = envs.make(env_id, num_envs=B, obs_mode="encoded", seed=5, env_offset=11)
    env.reset()
    A, V = env.num_agents, env.cfg.view_size
    pol = LinearPolicy.random(A, V, n_actions=7, epsilon=0.2, seed=77, rng_seed=A)
    ob = mg_oracle.OracleBatch(env.cfg, B, seed=5, env_offset=11, threads=8)
    ob.reset()
    first = np.random.RandomState(A).randint(0, 7, size=(B, A)).astype(np.int32)
    l0 = env._lib.mg_launch_count()
    obs, rew, done, act = (x.cpu().numpy() for x in env.rollout_policy(pol, first, T))
    launches = env._lib.mg_launch_count() - l0
    o2, r2, d2, a2 = policy_oracle.closed_loop(ob, pol, first, T)
    ok = np.array_equal(act, a2) and np.array_equal(obs, o2) and np.array_equal(rew.view(np.uint64), r2.view(np.uint64)) and np.array_equal(done, d2.astype(bool))
    print(f"{env_id} A={A} V={V} B={B}: {'ok' if ok else 'MISMATCH'} ({launches} launches for {T} steps)", flush=True)
    assert ok
