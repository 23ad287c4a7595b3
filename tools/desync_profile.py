"""One family of 65 536 envs whose step counters are spread uniformly over the episode length (the steady state of a
long-running batch: ~655 envs finish per step, in ~27 % of the tiles); N steps.  Run under ncu to capture one such launch:
  ncu --set full --import-source on -k regex:fused2_kernel -s 120 -c 1 -o gpurun_out/prof_desync python tools/desync_profile.py 130"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marlgrid_b200 import envs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 130
B = 65536
env = envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=B, obs_mode="encoded", seed=1337)
env.reset()
env.envrec[:, 0] = torch.randint(0, 100, (B,), device="cuda", dtype=torch.int32)
acts = torch.stack([env.random_actions(t) for t in range(16)])
import ctypes

stats = (ctypes.c_uint64 * 2)()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for t in range(n):
    if t == n // 2:
        env._lib.mg_pregen_stats(stats, 1)
        e0.record()
    env.step(acts[t % 16])
e1.record()
torch.cuda.synchronize()
env._lib.mg_pregen_stats(stats, 0)
print(f"{1e3 * e0.elapsed_time(e1) / (n - n // 2):.2f} us per step; envs finishing per step ~{int((env.step_count == 0).sum())}; "
      f"worlds copied from the generator {int(stats[0])}, generated in the step kernel {int(stats[1])}")
