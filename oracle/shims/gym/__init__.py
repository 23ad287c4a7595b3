"""Minimal stand-in for the third-party `gym` package (gym<=0.21 API surface).

TEST INFRASTRUCTURE ONLY.  The reference (kandouss/marlgrid) imports `gym`, which is
not installed in this image and cannot be installed (no network).  This shim supplies
exactly the names the reference touches:
  gym.Env                      (marlgrid/base.py:334)
  gym.core.Wrapper             (marlgrid/utils/video.py:55)
  gym.spaces.{Box,Discrete,Tuple,Dict}  (marlgrid/agents.py:58-83, base.py:378,384)
  gym.utils.seeding.np_random  (marlgrid/base.py:373)
  gym.envs.registration.register / gym.make   (marlgrid/envs/__init__.py:10,55)
It lets the UNMODIFIED reference source be imported from /root/reference.
"""
import importlib

from . import spaces, core, utils, envs  # noqa: F401
from .core import Env, Wrapper  # noqa: F401
from .envs.registration import register, make, registry  # noqa: F401
