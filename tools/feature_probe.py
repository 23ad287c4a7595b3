"""Throughput of env kwargs that leave the specialised fused kernel (general fused kernel / step kernel + observe kernel):
cfg3 (MarlGrid-3AgentCluttered15x15, 65 536 envs, encoded observations) with one kwarg changed, 200 warm steps enqueued from C."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from marlgrid_b200 import envs  # noqa: E402
from marlgrid_b200.agents import GridAgentInterface  # noqa: E402

B = 65536
L = None
for name, akw, ekw in (("(specialised kernel)", {}, {}), ("(general fused kernel, forced)", {}, {}), ("(step + observe kernels, forced)", {}, {}), ("hide_item_types=['Goal']", {"hide_item_types": ["Goal"]}, {}), ("hide_item_types=['Wall']", {"hide_item_types": ["Wall"]}, {}),
                       ("hide_item_types=['Agent']", {"hide_item_types": ["Agent"]}, {}), ("ghost_mode=False", {}, {"ghost_mode": False}),
                       ("respawn=True", {}, {"respawn": True}), ("see_through_walls=True", {"see_through_walls": True}, {}),
                       ("spawn_delay=5", {"spawn_delay": 5}, {}), ("view_offset=1", {"view_offset": 1}, {}),
                       ("human_player.py config (goal cycle 13x13, respawn=True, 3 agents)", {}, {})):
    if name.startswith("human_player"):
        env = envs.ClutteredGoalCycleEnv(agents=[GridAgentInterface(color=c, view_size=7, view_offset=1, view_tile_size=11) for c in ("red", "blue", "purple")],
                                         grid_size=13, max_steps=250, clutter_density=0.15, respawn=True, reward_decay=False, n_bonus_tiles=3,
                                         initial_reward=True, penalty=-1.5, num_envs=B, obs_mode="encoded", seed=1337)
    else:
      env = envs.ClutteredMultiGrid(agents=[GridAgentInterface(color=c, view_size=7, view_tile_size=8, **akw) for c in ("red", "blue", "purple")],
                                  grid_size=15, clutter_density=0.15, num_envs=B, obs_mode="encoded", seed=1337, **ekw)
    L = env._lib
    L.mg_debug_force_general_fused(1 if "general fused" in name else 0)
    L.mg_debug_force_two_kernels(1 if "step + observe" in name else 0)
    env.reset()
    actions = torch.stack([env.random_actions(t) for t in range(100)])
    env.rollout(actions[:20])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.rollout(actions[:60])  # steps 20..79 of the first episode: no env finishes by time-out in this window
    e1.record()
    torch.cuda.synchronize()
    us_steady = 1e3 * e0.elapsed_time(e1) / 60
    env.rollout(actions[:20])
    torch.cuda.synchronize()
    l0 = L.mg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.rollout(actions)
    env.rollout(actions)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 200
    print(f"{name:34s} {us:8.2f} us per step  {B / us * 1e6:.3e} env-steps/s  {(L.mg_launch_count() - l0) / 200:.2f} launches per step   ({us_steady:.2f} us per step between the all-reset steps)", flush=True)
    del env
