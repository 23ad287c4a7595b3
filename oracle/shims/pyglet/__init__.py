"""Empty stand-in for pyglet (marlgrid/rendering.py:1-2 imports it unconditionally;
GUI windows are out of scope).  TEST INFRASTRUCTURE ONLY."""
from . import gl  # noqa: F401
