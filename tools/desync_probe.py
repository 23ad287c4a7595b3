"""How the step time evolves while the episodes of a batch drift apart (tools/gpu_check.sh desync): a forward-biased random
policy ends episodes at irregular times, so after a few thousand steps every step finds a few finished envs in many tiles."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marlgrid_b200 import envs  # noqa: E402

B = 65536
env = envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=B, obs_mode="encoded", seed=1337)
env.reset()
acts = torch.stack([env.random_actions(t) for t in range(125)])
acts[torch.rand(acts.shape, device=acts.device) < 0.5] = 2  # forward-biased: agents do reach the goal
for block in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(8):
        env.rollout(acts)
    e1.record()
    torch.cuda.synchronize()
    sc = env.step_count
    print(f"steps {block * 1000:5d}..{block * 1000 + 999:5d}: {e0.elapsed_time(e1):6.2f} us per step; step_count spread: min {int(sc.min())} max {int(sc.max())}, "
          f"envs at step_count < 10: {int((sc < 10).sum())}")
