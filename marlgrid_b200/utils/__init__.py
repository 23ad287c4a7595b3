"""marlgrid.utils mirror (reference: marlgrid/utils/): GridRecorder."""
from .video import GridRecorder, export_video, render_frames  # noqa: F401
