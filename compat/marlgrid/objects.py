"""marlgrid.objects -> marlgrid_b200.objects (reference: marlgrid/objects.py; encodings only: the object model lives in the kernels)."""
from marlgrid_b200.objects import *  # noqa: F401,F403
