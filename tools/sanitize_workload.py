"""Small workload for compute-sanitizer (tools/gpu_check.sh sanitize): every path of the specialised fused kernel --
encoded and RGB observation, full and ragged tiles, more tiles than CTAs' first round, the all-reset step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marlgrid_b200 import envs  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 103
enc = envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=200, obs_mode="encoded", seed=5)
rgb = envs.make("MarlGrid-4AgentEmpty9x9-v0", num_envs=72, obs_mode="rgb", seed=6)
enc.reset()
rgb.reset()
for t in range(steps):
    enc.step(enc.random_actions(t))
    rgb.step(rgb.random_actions(t))
torch.cuda.synchronize()
assert int(enc.episode.min().item()) >= (2 if steps >= 100 else 1) and int(enc.err.max().item()) == 0 and int(rgb.err.max().item()) == 0
print("sanitize workload ok:", steps, "steps; checksums", int(enc.obs.sum().item()), int(rgb.obs.long().sum().item()))
