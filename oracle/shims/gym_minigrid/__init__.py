"""Stand-in for the third-party `gym_minigrid` package (only `rendering` is used,
marlgrid/base.py:6,14 and marlgrid/objects.py:3-8).  TEST INFRASTRUCTURE ONLY."""
from . import rendering  # noqa: F401
