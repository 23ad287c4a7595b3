"""marlgrid.base -> the batched env runtime (reference: marlgrid/base.py MultiGridEnv)."""
from marlgrid_b200.env import BatchedMultiGridEnv, UnbatchedView  # noqa: F401
from marlgrid_b200.envs import MultiGridEnv  # noqa: F401
