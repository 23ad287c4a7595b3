"""marlgrid.utils.video -> marlgrid_b200.utils.video (reference: marlgrid/utils/video.py)."""
from marlgrid_b200.utils.video import GridRecorder, export_video, render_frames  # noqa: F401
