"""Host-side tile atlas for the RGB observation kernel.

The reference draws every visible cell through MultiGrid.render_tile (marlgrid/base.py:275-299),
whose inputs are a small closed set: {empty, Wall, Goal, BonusTile} x {no agent, agent q facing d}.
All those tiles are rendered ONCE here (init time, numpy) and uploaded; the kernel only gathers.

Layout: uint8 [n_tiles][4][ts][ts][3]; tile id = kind*(1+4A) + (0 | 1 + 4*q + dir_q), second axis
is the view orientation k, holding rotate_grid(tile, k) (base.py:67-80,324) so the kernel never
rotates pixels.  kind: 0 empty cell, 1 Wall('worst'), 2 Goal('green'), 3 BonusTile('yellow').

The pixel maths restates (independently, vectorised) what the reference obtains from
gym_minigrid.rendering -- fill_coords / point_in_triangle / rotate_fn / downsample -- at its call
sites GridAgent.render (marlgrid/objects.py:150-153), Wall/Goal/BonusTile.render (objects.py:209,
226,288), MultiGrid.render_object (base.py:252-258), empty_tile (base.py:245-250) and blend_tiles
(base.py:260-273).  gym_minigrid is an unpinned third-party package absent from the reference
tree; tests/golden/atlas_*.npz freezes the tiles the reference itself renders under the shim.
"""
import math

import numpy as np

from .objects import COLORS, IDX_TO_COLOR

SUBDIVS = 3  # base.py:277
STATIC_KIND_COLOURS = ("worst", "green", "yellow")  # Wall() objects.py:47 default, Goal cluttered.py:29, BonusTile goalcycle.py:37


def rotate_tile(tile, k):
    """rotate_grid on the (y, x) axes of an image tile (base.py:67-80)."""
    k %= 4
    if k == 0:
        return tile
    if k == 1:
        return np.moveaxis(tile[::-1, :], 0, 1)
    if k == 2:
        return tile[::-1, ::-1]
    return np.moveaxis(tile[:, ::-1], 0, 1)


def empty_tile(ts):
    alpha = max(0, min(20, ts - 10))
    img = np.full((ts, ts, 3), alpha, dtype=np.uint8)
    img[1:, :-1] = 0
    return img


def _triangle_mask(n, theta):
    """Supersampled membership of the agent triangle rotated by theta about the tile centre."""
    a = (0.12, 0.19)
    b = (0.87, 0.50)
    c = (0.12, 0.81)
    v0 = (c[0] - a[0], c[1] - a[1])
    v1 = (b[0] - a[0], b[1] - a[1])
    dot00 = v0[0] * v0[0] + v0[1] * v0[1]
    dot01 = v0[0] * v1[0] + v0[1] * v1[1]
    dot11 = v1[0] * v1[0] + v1[1] * v1[1]
    inv_denom = 1 / (dot00 * dot11 - dot01 * dot01)
    cs, sn = math.cos(-theta), math.sin(-theta)
    mask = np.zeros((n, n), dtype=bool)
    for y in range(n):
        yf = (y + 0.5) / n
        for x in range(n):
            xf = (x + 0.5) / n
            xr, yr = xf - 0.5, yf - 0.5
            x2 = 0.5 + xr * cs - yr * sn
            y2 = 0.5 + yr * cs + xr * sn
            v2 = (x2 - a[0], y2 - a[1])
            dot02 = v0[0] * v2[0] + v0[1] * v2[1]
            dot12 = v1[0] * v2[0] + v1[1] * v2[1]
            u = (dot11 * dot02 - dot01 * dot12) * inv_denom
            v = (dot00 * dot12 - dot01 * dot02) * inv_denom
            mask[y, x] = (u >= 0) and (v >= 0) and (u + v) < 1
    return mask


def _downsample(img, f):
    h, w = img.shape[0] // f, img.shape[1] // f
    return img.reshape(h, f, w, f, 3).mean(axis=3).mean(axis=1)


def agent_base_tile(colour_idx, direction, ts):
    big = np.zeros((ts * SUBDIVS, ts * SUBDIVS, 3), dtype=np.uint8)
    big[_triangle_mask(ts * SUBDIVS, 0.5 * np.pi * direction)] = COLORS[IDX_TO_COLOR[colour_idx]]
    return _downsample(big, SUBDIVS).astype(np.uint8)


def solid_tile(colour_name, ts):
    big = np.zeros((ts * SUBDIVS, ts * SUBDIVS, 3), dtype=np.uint8)
    big[:, :] = COLORS[colour_name]
    return _downsample(big, SUBDIVS).astype(np.uint8)


def blend_tiles(img1, img2):
    alpha = img2.sum(2, keepdims=True)
    max_alpha = alpha.max()
    if max_alpha == 0:
        return img1
    return ((img1 * (max_alpha - alpha) + img2 * alpha) / max_alpha).astype(img1.dtype)


def _with_border(img, ts):
    corners = img[([0, 0, -1, -1], [0, -1, 0, -1])]
    if (corners == 0).all(axis=-1).any():
        img = img + empty_tile(ts)  # uint8 wrap-around, like the reference (base.py:297-298)
    return img


def build_atlas(agent_colors, ts, n_static_kinds=3):
    """-> uint8 [n_tiles, 4, ts, ts, 3] for agents with the given colour indices."""
    A = len(agent_colors)
    per_kind = 1 + 4 * A
    tiles = np.zeros(((n_static_kinds + 1) * per_kind, ts, ts, 3), dtype=np.uint8)
    agent = [[agent_base_tile(c, d, ts) for d in range(4)] for c in agent_colors]
    tiles[0] = empty_tile(ts)  # obj is None: no border pass (base.py:279-280)
    for q in range(A):
        for d in range(4):
            tiles[1 + 4 * q + d] = _with_border(agent[q][d], ts)
    for k in range(1, n_static_kinds + 1):
        base = solid_tile(STATIC_KIND_COLOURS[k - 1], ts)
        tiles[k * per_kind] = _with_border(base, ts)
        for q in range(A):
            for d in range(4):
                tiles[k * per_kind + 1 + 4 * q + d] = _with_border(blend_tiles(base, agent[q][d]), ts)
    atlas = np.stack([np.stack([rotate_tile(t, k) for k in range(4)]) for t in tiles])
    return np.ascontiguousarray(atlas)
