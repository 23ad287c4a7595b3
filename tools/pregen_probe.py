"""How much the background world generator (mg_pregen.cu) costs the step kernel: back-to-back steps of one family in lock step
with the automatic generator launches off / on, and the duration of one generator pass alone (nothing to do / everything)."""
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
acts = torch.stack([env.random_actions(t) for t in range(96)])


def timed(fn, n=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


cfg, st = ctypes.byref(env.cfg), ctypes.byref(env._state)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for auto in (0, 1, 0, 1):
    L.mg_pregen_set_auto(auto)
    env.reset()
    env.rollout(acts[:8])
    L.mg_pregen_drain()
    us = timed(lambda: [env.rollout(acts) for _ in range(5)], 5 * 96)
    print(f"auto generator {'on ' if auto else 'off'}: {us:.2f} us per step (480 steps back to back, 5 all-reset steps)")
L.mg_pregen_set_auto(0)
env.reset()
torch.cuda.synchronize()
print(f"generator pass on the main stream, every slot stale: {timed(lambda: L.mg_pregen_run(cfg, st, stream)):.1f} us")
print(f"generator pass on the main stream, nothing to do:    {timed(lambda: L.mg_pregen_run(cfg, st, stream)):.1f} us")
print(f"10 passes, nothing to do:                            {timed(lambda: [L.mg_pregen_run(cfg, st, stream) for _ in range(10)], 10):.1f} us each")
# steps with an explicit pass on the side stream after every n-th step
side = ctypes.c_void_p(-1)
for every in (1, 2, 4, 8, 16):
    env.reset()
    L.mg_pregen_run(cfg, st, stream)
    torch.cuda.synchronize()

    def run():
        for t in range(90):
            env.step(acts[t])
            if t % every == 0:
                L.mg_pregen_run(cfg, st, side)

    us = timed(run, 90)
    L.mg_pregen_drain()
    print(f"90 steps (no reset), side-stream pass after every {every:2d}. step: {us:.2f} us per step")
