"""marlgrid.envs -> marlgrid_b200.envs (reference: marlgrid/envs/__init__.py: scenario classes, registry ids, env_from_config)."""
from marlgrid_b200.envs import *  # noqa: F401,F403
from marlgrid_b200.envs import (ClutteredGoalCycleEnv, ClutteredMultiGrid, EmptyMultiGrid, MultiGridEnv, env_from_config, make,  # noqa: F401
                                register_marl_env, registered_envs, registry)
