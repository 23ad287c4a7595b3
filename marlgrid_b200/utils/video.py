"""Filming episodes of a batched env.

Public surface of the reference's `marlgrid/utils/video.py` (`export_video` :7-35, `render_frames` :38-53, `GridRecorder`
:55-154: same names, constructor keywords, attributes `recording` / `frames` / `ptr` / `reset_count`, same file names on disk),
built around the batched device env instead of one Python env: a recorder films ONE env of the batch (`index`) or a mosaic of
several (`index=[...]`), pulls each frame from `env.render(index=...)` (marlgrid_b200/render.py) only while a reel is open, and
keeps the frames of the running episode in a growing list -- nothing is allocated for episodes that are not filmed.

Forced differences: no gym.core.Wrapper to inherit from (unknown attributes are forwarded to the wrapped env); `max_steps=None`
means "the env's own max_steps" (the reference names an undefined `default_max_steps` there, video.py:91); without moviepy the
movie is written as an animated GIF through PIL.
"""
import os

import numpy as np


# ---------------------------------------------------------------------------------------------------------------
# writers
# ---------------------------------------------------------------------------------------------------------------
def _as_uint8_stack(frames):
    x = np.stack(list(frames)) if not isinstance(frames, np.ndarray) else frames
    if x.dtype.kind == "f" and x.size and float(x.max()) < 1.0:  # unit-range floats (video.py:13-14)
        x = np.clip(x * 255.0, 0, 255).astype(np.uint8)
    return x.astype(np.uint8, copy=False)


def _enlarge(x, factor):
    factor = int(factor) if factor else 1
    return x if factor == 1 else x.repeat(factor, axis=1).repeat(factor, axis=2)


def _target(path):
    path = os.path.abspath(os.path.expanduser(path))
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    return path


def export_video(X, outfile, fps=30, rescale_factor=2):
    """Frames [T, H, W, 3] -> a movie at `outfile` (an animated .gif beside it when moviepy is missing).  Returns the path written."""
    clip = _enlarge(_as_uint8_stack(X), rescale_factor)
    outfile = _target(outfile)
    try:
        from moviepy.editor import ImageSequenceClip  # type: ignore
    except Exception:  # noqa: BLE001 -- not installed here
        from PIL import Image

        gif = os.path.splitext(outfile)[0] + ".gif"
        pages = [Image.fromarray(np.ascontiguousarray(f), "RGB") for f in clip]
        pages[0].save(gif, save_all=True, append_images=pages[1:], duration=max(1, round(1000 / fps)), loop=0)
        return gif
    ImageSequenceClip(list(clip), fps=fps).write_videofile(outfile, fps=fps)
    return outfile


def render_frames(X, path, ext="png"):
    """One image per frame, `path`/frame_<k>.<ext>; an extension on `path` itself is ignored.  Returns the directory."""
    from PIL import Image

    stem, suffix = os.path.splitext(path)
    folder = os.path.abspath(os.path.expanduser(stem if suffix and os.sep not in suffix else path))
    os.makedirs(folder, exist_ok=True)
    for k, frame in enumerate(_as_uint8_stack(X)):
        Image.fromarray(np.ascontiguousarray(frame), "RGB").save(os.path.join(folder, f"frame_{k}.{ext}"))
    return folder


# ---------------------------------------------------------------------------------------------------------------
# recorder
# ---------------------------------------------------------------------------------------------------------------
class _Reel:
    """The frames of the episode being filmed (bounded by `limit`)."""

    def __init__(self, limit):
        self.limit = int(limit)
        self.shots = []

    def full(self):
        return len(self.shots) >= self.limit

    def add(self, frame):
        if not self.full():
            self.shots.append(np.array(frame, copy=True))


class GridRecorder:
    """Wraps an env; episodes are filmed while `recording` is set, or whenever `auto_save_interval` resets have gone by since
    the last saved one.  The frame BEFORE every step is kept; at the next reset the final frame is added and the episode is
    written as `frames_<n>/frame_<k>.png` and / or `video_<n>.mp4` under `save_root`, n = the reset counter (which advances by
    the number of parallel envs per reset, like the reference's)."""

    default_max_len = 1000
    default_video_kwargs = {"fps": 20, "rescale_factor": 1}

    def __init__(self, env, save_root, max_steps=1000, auto_save_images=True, auto_save_videos=True, auto_save_interval=None,
                 render_kwargs={}, video_kwargs={}, index=0):
        self.env = env
        self.index = index
        self.save_root = self.fix_path(save_root)
        self.auto_save_images, self.auto_save_videos, self.auto_save_interval = auto_save_images, auto_save_videos, auto_save_interval
        self.render_kwargs = dict(render_kwargs)
        self.video_kwargs = dict(self.default_video_kwargs, **video_kwargs)
        self.n_parallel = getattr(env, "num_envs", 1)
        horizon = max_steps if max_steps is not None else (getattr(env, "max_steps", 0) or self.default_max_len)
        self.max_steps = horizon + 1  # every pre-step frame of a full episode + the closing one
        self.recording = False
        self.reset_count = 0
        self.last_save = -10000
        self._reel = None

    def __getattr__(self, name):  # only reached for names the recorder does not define: forwarded, as a gym Wrapper would
        if name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    # ---- what the reference exposes -------------------------------------------------------------------------
    @staticmethod
    def fix_path(path):
        return os.path.abspath(os.path.expanduser(path))

    @property
    def frames(self):
        return None if self._reel is None else self._reel.shots

    @property
    def ptr(self):
        return 0 if self._reel is None else len(self._reel.shots)

    @property
    def should_record(self):
        due = self.auto_save_interval is not None and self.reset_count - self.last_save >= self.auto_save_interval
        return bool(self.recording or due)

    # ---- frames ---------------------------------------------------------------------------------------------
    def _shoot(self):
        batched = hasattr(self.env, "num_envs")
        if isinstance(self.index, (list, tuple)):  # a mosaic: the filmed envs side by side
            return np.concatenate([self.env.render(index=i, mode="rgb_array", **self.render_kwargs) for i in self.index], axis=1)
        if batched:
            return self.env.render(index=self.index, mode="rgb_array", **self.render_kwargs)
        return self.env.render(mode="rgb_array", **self.render_kwargs)

    def append_current_frame(self):
        if not self.should_record:
            return
        if self._reel is None:
            self._reel = _Reel(self.max_steps)
        if not self._reel.full():
            self._reel.add(self._shoot())

    # ---- export ---------------------------------------------------------------------------------------------
    def _where(self, save_root, name):
        return os.path.join(self.fix_path(self.save_root if save_root is None else save_root), name)

    def export_frames(self, episode_id=None, save_root=None):
        return render_frames(self.frames, self._where(save_root, f"frames_{self.reset_count}" if episode_id is None else episode_id))

    def export_video(self, episode_id=None, save_root=None):
        return export_video(self.frames, self._where(save_root, f"video_{self.reset_count}.mp4" if episode_id is None else episode_id),
                            **self.video_kwargs)

    def export_both(self, episode_id, save_root=None):
        self.export_frames(f"{episode_id}_frames", save_root=save_root)
        self.export_video(f"{episode_id}.mp4", save_root=save_root)

    # ---- gym surface ----------------------------------------------------------------------------------------
    def reset(self, **kwargs):
        if self.should_record and self.ptr > 0:  # close the reel of the episode that just ended
            self.append_current_frame()
            if self.auto_save_images:
                self.export_frames()
            if self.auto_save_videos:
                self.export_video()
            self.last_save = self.reset_count
        self._reel = None
        self.reset_count += self.n_parallel
        return self.env.reset(**kwargs)

    def step(self, action):
        self.append_current_frame()
        return self.env.step(action)
