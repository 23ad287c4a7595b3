"""marlgrid_b200 -- B200-native batched MarlGrid (drop-in for the hot path of kandouss/marlgrid).

    import marlgrid_b200 as marlgrid
    env = marlgrid.envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=65536, obs_mode="encoded")
    obs = env.reset(); obs, rew, done, _ = env.step(actions)      # torch tensors on the GPU

Importing the package does not need a GPU; constructing an env does (there is no CPU fallback).
"""
from . import agents, config, objects  # noqa: F401
from .agents import GridAgentInterface, IndependentLearners, LearningAgent  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):  # lazy: `envs`/`env` import torch
    if name in ("envs", "env"):
        import importlib

        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
