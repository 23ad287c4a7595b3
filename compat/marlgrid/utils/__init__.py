"""marlgrid.utils -> marlgrid_b200.utils."""
from marlgrid_b200.utils import *  # noqa: F401,F403
