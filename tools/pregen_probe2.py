"""Generator passes timed on the side stream while step kernels run back to back (desynchronised family)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marlgrid_b200 import _lib, envs  # noqa: E402

L = _lib.load()
B = 65536
env = envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=B, obs_mode="encoded", seed=1337)
env.reset()
env.envrec[:, 0] = torch.randint(0, 100, (B,), device="cuda", dtype=torch.int32)
acts = torch.stack([env.random_actions(t) for t in range(96)])
cfg, st = ctypes.byref(env.cfg), ctypes.byref(env._state)
L.mg_pregen_set_auto(0)
side = torch.cuda.Stream(priority=0)
side_p = ctypes.c_void_p(side.cuda_stream)
main = torch.cuda.current_stream()
for every in (1, 4, 8):
    env.rollout(acts)  # settle
    torch.cuda.synchronize()
    evs = []
    stats = (ctypes.c_uint64 * 2)()
    L.mg_pregen_stats(stats, 1)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for t in range(480):
        env.step(acts[t % 96])
        if t % every == 0:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(side)
            L.mg_pregen_run(cfg, st, side_p)
            b.record(side)
            evs.append((a, b))
    s1.record()
    torch.cuda.synchronize()
    L.mg_pregen_stats(stats, 0)
    d = sorted(1e3 * a.elapsed_time(b) for a, b in evs)
    lag = 1e3 * s0.elapsed_time(evs[-1][1])
    print(f"every {every}: step {1e3 * s0.elapsed_time(s1) / 480:.2f} us; generator pass min {d[0]:.1f} median {d[len(d) // 2]:.1f} max {d[-1]:.1f} us; "
          f"last pass ended {lag - 1e3 * s0.elapsed_time(s1):+.0f} us relative to the last step; copied {int(stats[0])} generated in step kernel {int(stats[1])}")
