"""Actor on the device, learner in torch (marlgrid_b200.learners.LinearQTrainer): every iteration plays `horizon` steps of all envs
in ONE kernel launch with the int8-quantised epsilon-greedy policy (mg_rollout_policy), then each agent's linear Q-function takes one
TD(0) step on the returned batch.  Under torchrun (one process per GPU) every rank steps its own shard of the envs and the learners'
gradients are averaged with NCCL -- the system's only collective.

    python tools/train_linear_q.py [iterations] [num_envs] [env_id]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from marlgrid_b200 import envs  # noqa: E402
from marlgrid_b200.learners import LinearQLearner, LinearQTrainer  # noqa: E402
from marlgrid_b200.sharding import shard_range  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 150
total = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
env_id = sys.argv[3] if len(sys.argv) > 3 else "MarlGrid-2AgentEmpty9x9-v0"
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
off, cnt = shard_range(total, rank, world)
env = envs.make(env_id, num_envs=cnt, obs_mode="encoded", seed=1337, env_offset=off, device=f"cuda:{local}")
env.reset()
colors = ["red", "blue", "purple", "orange"]
learners = [LinearQLearner(view_size=env.cfg.view_size, gamma=0.95, lr=1e-3, target_period=10, device=env.device, seed=k, color=colors[k]) for k in range(env.num_agents)]
trainer = LinearQTrainer(env, learners, horizon=32, epsilon=0.15)
t0 = time.perf_counter()
window = []
for it in range(iters):
    stats = trainer.iterate()
    window.append(stats)
    if (it + 1) % 10 == 0:
        r = sum(s["reward_per_env_step"] for s in window) / len(window)
        ep = sum(s["episodes"] for s in window)
        loss = sum(sum(s["loss"]) for s in window) / len(window)
        if world > 1:
            v = torch.tensor([r, float(ep), loss], device=env.device, dtype=torch.float64)
            torch.distributed.all_reduce(v)
            r, ep, loss = float(v[0]) / world, float(v[1]), float(v[2]) / world
        if rank == 0:
            steps = (it + 1) * trainer.horizon * total
            print(f"iter {it + 1:4d}  env-steps {steps:.3e}  reward/env-step {r:.5f}  episodes ended {int(ep):7d}  td-loss {loss:.5f}  "
                  f"{steps / (time.perf_counter() - t0):.3e} env-steps/s incl. learner", flush=True)
        window = []
torch.cuda.synchronize()
if world > 1:
    torch.distributed.destroy_process_group()
