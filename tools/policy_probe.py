"""Cost of the policy hand-off inside the persistent rollout kernel: mg_rollout_policy (closed loop, int8 linear policy evaluated
on the observation tile in shared memory) next to mg_rollout_persistent (open loop, fixed action tape) and to the host loop
env.step + a torch policy (what the reference's README loop costs on the same GPU)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from marlgrid_b200 import envs
from marlgrid_b200.policy import LinearPolicy

B, T = 65536, 100
env = envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=B, obs_mode="encoded", seed=1337)
env.reset()
A, V = env.cfg.n_agents, env.cfg.view_size
pol = LinearPolicy.random(A, V, n_actions=7, epsilon=0.1, seed=5)
tape = torch.stack([env.random_actions(t) for t in range(T)])
out = (torch.empty((T, B, A, V, V, 3), dtype=torch.uint8, device="cuda"), torch.empty((T, B, A), dtype=torch.float64, device="cuda"),
       torch.empty((T, B), dtype=torch.bool, device="cuda"), torch.empty((T, B, A), dtype=torch.int32, device="cuda"))
w = torch.from_numpy(pol.weights).cuda().float()  # [A][K][n]
b = torch.from_numpy(pol.bias).cuda().float()


def host_loop():
    act = tape[0]
    for _ in range(T):
        obs, _, _, _ = env.step(act)
        logits = torch.einsum("bai,aki->bak", obs.reshape(B, A, -1).float(), w) + b
        act = logits.argmax(-1).to(torch.int32)


import ctypes
import time


def timed(name, fn, reps=5, gap=False):
    """gap: the launches are separated by a host synchronisation + 2 ms (a learner update between two rollouts), timed one by one."""
    fn(); torch.cuda.synchronize()
    st = (ctypes.c_uint64 * 2)()
    env._lib.mg_pregen_stats(st, 1)
    if gap:
        ms = 0.0
        for _ in range(reps):
            time.sleep(0.002)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    us = ms * 1e3 / (reps * T)
    env._lib.mg_pregen_stats(st, 0)
    hit = f"  pre-generated worlds used {st[0]} of {st[0] + st[1]} resets" if st[0] + st[1] else ""
    print(f"{name}: {us:.2f} us per step, {B / us * 1e6:.3e} env-steps/s{hit}", flush=True)


timed("open loop, mg_rollout_persistent (1 launch / 100 steps)", lambda: env.rollout_all(tape, out=out[:3]))
timed("closed loop, mg_rollout_policy (1 launch / 100 steps, policy in the kernel)", lambda: env.rollout_policy(pol, tape[0], T, out=out))
env._lib.mg_debug_force_general_fused(1)
timed("closed loop, step launch + policy launch per step (fallback route)", lambda: env.rollout_policy(pol, tape[0], T, out=out))
env._lib.mg_debug_force_general_fused(0)
timed("closed loop on the host: env.step + torch einsum/argmax per step", host_loop, reps=2)

# the steady state of a long-running batch: step counters spread uniformly over the episode length, ~1 % of the envs time out and
# are regenerated in every step.  One launch per step waits for its slowest CTA every step; K steps per launch do not: a CTA that
# meets a reset falls behind and catches up, only the sum over the launch counts.
env.reset()
env.envrec[:, 0] = torch.randint(0, 100, (B,), device="cuda", dtype=torch.int32)
env.rollout(tape)  # spreads the episode numbers, lets the background generator catch up
timed("desynchronised episodes, one launch per step (mg_rollout_fused)", lambda: env.rollout(tape))
timed("desynchronised episodes, open loop, 1 launch / 100 steps", lambda: env.rollout_all(tape, out=out[:3]))
timed("desynchronised episodes, closed loop, 1 launch / 100 steps", lambda: env.rollout_policy(pol, tape[0], T, out=out))
timed("desynchronised episodes, open loop, 1 launch / 100 steps, 2 ms between launches", lambda: env.rollout_all(tape, out=out[:3]), gap=True)
timed("desynchronised episodes, closed loop, 1 launch / 100 steps, 2 ms between launches", lambda: env.rollout_policy(pol, tape[0], T, out=out), gap=True)
