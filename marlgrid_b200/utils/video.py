"""Episode recording for batched envs -- host-side mirror of marlgrid/utils/video.py (GridRecorder :55-154,
export_video :7-35, render_frames :38-53).  Off the hot path: frames come from env.render() (marlgrid_b200/render.py).

Differences from the reference, all forced by the environment: there is no gym.core.Wrapper to inherit from (attribute access
is forwarded by hand), `max_steps=None` falls back to `env.max_steps` or 1000 (the reference names an undefined
`default_max_steps` there, video.py:91), and export_video writes an animated GIF through PIL when moviepy is not installed.
"""
import os

import numpy as np


def export_video(X, outfile, fps=30, rescale_factor=2):
    """video.py:7-35: frames [T, H, W, 3] -> movie file (moviepy when available, else an animated .gif next to `outfile`)."""
    if isinstance(X, list):
        X = np.stack(X)
    if np.issubdtype(X.dtype, np.floating) and X.max() < 1:
        X = (X * 255).astype(np.uint8).clip(0, 255)
    if rescale_factor is not None and rescale_factor != 1:
        X = np.kron(X, np.ones((1, int(rescale_factor), int(rescale_factor), 1))).astype(np.uint8)
    outfile = os.path.abspath(os.path.expanduser(outfile))
    os.makedirs(os.path.dirname(outfile), exist_ok=True)
    try:
        import moviepy.editor as mpy  # type: ignore
    except Exception:  # noqa: BLE001
        from PIL import Image

        gif = os.path.splitext(outfile)[0] + ".gif"
        frames = [Image.fromarray(np.ascontiguousarray(f), "RGB") for f in X]
        frames[0].save(gif, save_all=True, append_images=frames[1:], duration=max(1, int(1000 / fps)), loop=0)
        return gif
    clip = mpy.VideoClip(lambda t: X[min(int(t * fps), len(X) - 1)], duration=len(X) / fps)
    clip.write_videofile(outfile, fps=fps)
    return outfile


def render_frames(X, path, ext="png"):
    """video.py:38-53: one image file per frame under `path` (a file extension in `path` is dropped)."""
    from PIL import Image

    if "." in os.path.basename(path):
        path = os.path.splitext(path)[0]
    os.makedirs(path, exist_ok=True)
    for k, frame in enumerate(X):
        Image.fromarray(np.ascontiguousarray(frame), "RGB").save(os.path.join(path, f"frame_{k}.{ext}"))
    return path


class GridRecorder:
    """video.py:55-154.  Wraps an env (batched: env `index` is the one filmed); while `recording` is set, or every
    `auto_save_interval` resets, the frame before every step is kept and exported at the next reset."""

    default_max_len = 1000
    default_video_kwargs = {"fps": 20, "rescale_factor": 1}

    def __init__(self, env, save_root, max_steps=1000, auto_save_images=True, auto_save_videos=True, auto_save_interval=None,
                 render_kwargs={}, video_kwargs={}, index=0):
        self.env = env
        self.index = index
        self.frames = None
        self.ptr = 0
        self.reset_count = 0
        self.last_save = -10000
        self.recording = False
        self.save_root = self.fix_path(save_root)
        self.auto_save_videos = auto_save_videos
        self.auto_save_images = auto_save_images
        self.auto_save_interval = auto_save_interval
        self.render_kwargs = dict(render_kwargs)
        self.video_kwargs = {**self.default_video_kwargs, **video_kwargs}
        self.n_parallel = getattr(env, "num_envs", 1)
        if max_steps is None:
            max_steps = getattr(env, "max_steps", 0) or self.default_max_len
        self.max_steps = max_steps + 1

    def __getattr__(self, name):  # what gym.core.Wrapper does for the reference
        return getattr(self.env, name)

    @staticmethod
    def fix_path(path):
        return os.path.abspath(os.path.expanduser(path))

    @property
    def should_record(self):
        if self.recording:
            return True
        if self.auto_save_interval is None:
            return False
        return (self.reset_count - self.last_save) >= self.auto_save_interval

    def export_frames(self, episode_id=None, save_root=None):
        save_root = self.save_root if save_root is None else save_root
        episode_id = f"frames_{self.reset_count}" if episode_id is None else episode_id
        return render_frames(self.frames[: self.ptr], os.path.join(self.fix_path(save_root), episode_id))

    def export_video(self, episode_id=None, save_root=None):
        save_root = self.save_root if save_root is None else save_root
        episode_id = f"video_{self.reset_count}.mp4" if episode_id is None else episode_id
        return export_video(self.frames[: self.ptr], os.path.join(self.fix_path(save_root), episode_id), **self.video_kwargs)

    def export_both(self, episode_id, save_root=None):
        self.export_frames(f"{episode_id}_frames", save_root=save_root)
        self.export_video(f"{episode_id}.mp4", save_root=save_root)

    def reset(self, **kwargs):
        if self.should_record and self.ptr > 0:
            self.append_current_frame()
            if self.auto_save_images:
                self.export_frames()
            if self.auto_save_videos:
                self.export_video()
            self.last_save = self.reset_count
        self.frames = None
        self.ptr = 0
        self.reset_count += self.n_parallel
        return self.env.reset(**kwargs)

    def append_current_frame(self):
        if self.should_record and self.ptr < self.max_steps:
            if hasattr(self.env, "num_envs"):
                new_frame = self.env.render(index=self.index, mode="rgb_array", **self.render_kwargs)
            else:
                new_frame = self.env.render(mode="rgb_array", **self.render_kwargs)
            if self.frames is None:
                self.frames = np.zeros((self.max_steps, *new_frame.shape), dtype=new_frame.dtype)
            self.frames[self.ptr] = new_frame
            self.ptr += 1

    def step(self, action):
        self.append_current_frame()
        return self.env.step(action)
