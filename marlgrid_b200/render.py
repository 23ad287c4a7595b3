"""Whole-grid "human" view of one env of a batch: MultiGridEnv.render(mode='rgb_array') of the reference
(marlgrid/base.py:714-795) -- the full grid at `tile_size` pixels per cell with the cells some agent can see
highlighted (MultiGrid.render with highlight_mask, base.py:301-331), and the agents' own views in columns beside it.

Off the hot path: one env at a time, composed on the host with numpy from
  * the env's planes and agent records (copied from the device),
  * the tile atlas (marlgrid_b200/atlas.py; the same tiles the RGB observation kernel gathers),
  * line-of-sight masks computed by the CUDA entry point mg_los_batch (occlude_mask, agents.py:298-343),
  * the agents' RGB observations computed by mg_obs_rgb.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import atlas as _atlas
from .config import AF_ACTIVE, AF_PLACED, F_SEE_THROUGH, MgState
from .objects import COLOR_TO_IDX

_PRESTIGE = COLOR_TO_IDX["prestige"]

_ATLAS_CACHE = {}


def _atlas_for(colors, ts, n_static_kinds):
    key = (tuple(colors), int(ts), int(n_static_kinds))
    if key not in _ATLAS_CACHE:
        _ATLAS_CACHE[key] = _atlas.build_atlas(list(colors), int(ts), int(n_static_kinds))
    return _ATLAS_CACHE[key]


def _view_exts(x, y, d, V, vo):
    """agents.py:237-266 get_view_exts -> (topX, topY)."""
    h = V // 2
    if d == 0:
        return x - vo, y - h
    if d == 1:
        return x - h, y - vo
    if d == 2:
        return x - V + 1 + vo, y - h
    return x - h, y - V + 1 + vo


def _sub_state(env, index):
    """MgState of the single env `index` (pointers offset into the batch tensors)."""
    st = MgState()
    st.grid = env.grid[index].data_ptr()
    st.agents = env.agent_rec[index].data_ptr()
    st.envrec = env.envrec[index].data_ptr()
    st.pregen = None
    st.prestige = env.prestige[index].data_ptr() if getattr(env, "prestige", None) is not None else None
    st.cellbits = None  # the bit-planes are tile-transposed (32 envs interleaved): a single env takes the byte-plane kernels
    st.n_envs, st.env_offset, st.seed = 1, env.env_offset + index, env._seed
    return st


def agent_views_rgb(env, index=0):
    """uint8 [A, V*ts, V*ts, 3]: the RGB observations of env `index` (gen_agent_obs base.py:453-460), whatever the env's obs_mode."""
    cfg = env.cfg
    A, V, ts = cfg.n_agents, cfg.view_size, cfg.view_tile_size
    if env.obs_mode == "rgb" and env.atlas is not None:
        at = env.atlas
    else:
        at = torch.from_numpy(_atlas_for([int(c) for c in cfg.agent_color[:A]], ts, cfg.n_static_kinds)).to(env.device)
    out = torch.empty((1, A, V * ts, V * ts, 3), dtype=torch.uint8, device=env.device)
    st = _sub_state(env, index)
    with torch.cuda.device(env.device):
        _lib.check(env._lib.mg_obs_rgb(ctypes.byref(cfg), ctypes.byref(st), at.data_ptr(), out.data_ptr(), env._stream()), "mg_obs_rgb")
    return out[0].cpu().numpy()


def visibility_masks(env, index=0):
    """bool [A, V, V]: each agent's line-of-sight mask in VIEW coordinates (gen_obs_grid's vis_mask, base.py:418-439); all
    False for inactive agents.  The crop / rotation is done here, the line of sight by the CUDA entry point mg_los_batch."""
    cfg = env.cfg
    A, V, vo, W, H = cfg.n_agents, cfg.view_size, cfg.view_offset, cfg.width, cfg.height
    planes = env.planes[index].cpu().numpy()           # [3, W, H]
    ag = env.agent_rec[index].cpu().numpy()               # [A, 16]
    opaque = (planes[0] == 8) | ((planes[0] == 11) & (planes[2] != 1))  # Wall / Door that is not open (objects.py:281-282,330-331)
    transp = np.ones((A, V, V), dtype=np.uint8)
    active = (ag[:, 3] & AF_ACTIVE) != 0
    for a in range(A):
        if not active[a]:
            continue
        x, y, d = int(ag[a, 0]), int(ag[a, 1]), int(ag[a, 2]) & 3
        tx, ty = _view_exts(x, y, d, V, vo)
        sub = np.ones((V, V), dtype=np.uint8)          # outside the world: empty, transparent (base.py:132-141)
        x0, x1, y0, y1 = max(0, tx), min(tx + V, W), max(0, ty), min(ty + V, H)
        sub[x0 - tx:x1 - tx, y0 - ty:y1 - ty] = ~opaque[x0:x1, y0:y1]
        transp[a] = _atlas.rotate_tile(sub, d + 1)     # rotate_grid(sub, rot_k = dir + 1), base.py:429-431
    if cfg.flags & F_SEE_THROUGH:                      # agents.py:294-295
        return np.repeat(active[:, None, None], V, 1).repeat(V, 2)
    t = torch.from_numpy(np.ascontiguousarray(transp)).to(env.device)
    m = torch.zeros_like(t)
    with torch.cuda.device(env.device):
        _lib.check(env._lib.mg_los_batch(t.data_ptr(), m.data_ptr(), A, V, V // 2, V - 1 - vo, env._stream()), "mg_los_batch")
    return (m.cpu().numpy() != 0) & active[:, None, None]


def prestige_colour(prestige, scale, allow_negative):
    """agents.py:103-111: new_color = (prestige_scaled * blue + (1 - prestige_scaled) * red).astype(int)."""
    s = 1 / (1 + np.exp(-prestige / scale)) if allow_negative else np.tanh(prestige / scale)
    return (s * np.array([0, 0, 255]) + (1.0 - s) * np.array([255, 0, 0])).astype(np.int64)


def prestige_tile(at, kind, agent_slot, per_kind, colour):
    """The tile of a cell whose agent is coloured 'prestige': the atlas holds that agent in white, i.e. its triangle's alpha
    (+ the empty tile's border, base.py:245-250,296-298); render_post multiplies alpha with the colour (agents.py:113-115),
    render_tile blends the result over the cell's object (blend_tiles, base.py:260-273)."""
    empty = at[0].astype(np.int64)
    alpha = at[agent_slot].astype(np.int64)[..., 0] - empty[..., 0]
    agent = np.right_shift(alpha[..., None] * colour, 8)
    if kind == 0:
        return (agent + empty).astype(np.uint8)
    base = at[kind * per_kind].astype(np.int64)
    sa = agent.sum(2, keepdims=True)
    m = int(sa.max())
    return (base if m == 0 else (base * (m - sa) + agent * sa) // m).astype(np.uint8)


def render(env, index=0, highlight=True, tile_size=32, show_agent_views=True, max_agents_per_col=3, agent_col_width_frac=0.3,
           agent_col_padding_px=2, pad_grey=100):
    """-> uint8 [H*tile_size, W*tile_size (+ agent view columns), 3]; same keyword arguments as the reference's render."""
    cfg = env.cfg
    A, V, vo, W, H, ts = cfg.n_agents, cfg.view_size, cfg.view_offset, cfg.width, cfg.height, int(tile_size)
    planes = env.planes[index].cpu().numpy()
    ag = env.agent_rec[index].cpu().numpy()
    placed = (ag[:, 3] & AF_PLACED) != 0
    stamp = ag[:, 8:12].copy().view(np.int32)[:, 0]
    per_kind = 1 + 4 * A
    # tile of every cell: the static object's kind, plus the head of the cell's agent queue (render_tile base.py:275-299, top_agent=None)
    kinds = np.array([cfg.kind_of_type[t] for t in range(15)] + [0xFF], dtype=np.int64)[np.minimum(planes[0], 15)]
    if (kinds == 0xFF).any():
        raise NameError("object has no working render() in the reference (marlgrid/objects.py:274-277,309-321,370)")
    tiles = kinds * per_kind
    head = {}
    for q in np.argsort(stamp, kind="stable"):
        if placed[q]:
            head.setdefault((int(ag[q, 0]), int(ag[q, 1])), int(q))
    for (x, y), q in head.items():
        tiles[x, y] += 1 + 4 * q + (int(ag[q, 2]) & 3)
    at = _atlas_for([int(c) for c in cfg.agent_color[:A]], ts, cfg.n_static_kinds)[:, 0]   # grid orientation 0
    img = at[tiles.T]                                             # [H, W, ts, ts, 3]: image row = y, column = x (base.py:319-324)
    for (x, y), q in head.items():  # 'prestige'-coloured agents: GridAgentInterface.render_post (agents.py:92-119) on the white tile
        if int(cfg.agent_color[q]) == _PRESTIGE and (ag[q, 3] & AF_ACTIVE):
            img[y, x] = prestige_tile(at, int(kinds[x, y]), 1 + 4 * q + (int(ag[q, 2]) & 3), per_kind,
                                      prestige_colour(float(env.prestige[index, q].item()), float(cfg.prestige_scale[q]), bool((cfg.prestige_neg_mask >> q) & 1)))
    img = np.ascontiguousarray(img.transpose(0, 2, 1, 3, 4)).reshape(H * ts, W * ts, 3)
    if highlight:  # cells inside some active agent's line of sight (base.py:741-753)
        vis = visibility_masks(env, index)
        hm = np.zeros((W, H), dtype=bool)
        for a in range(A):
            if not (ag[a, 3] & AF_ACTIVE):
                continue
            x, y, d = int(ag[a, 0]), int(ag[a, 1]), int(ag[a, 2]) & 3
            xl, yl = _view_exts(x, y, d, V, vo)
            xh, yh = xl + V, yl + V
            dxl, dyl, dxh, dyh = max(0, -xl), max(0, -yl), max(0, xh - W), max(0, yh - H)
            world_aligned = _atlas.rotate_tile(vis[a], -(d + 1))  # rotate_grid(vis_mask, orientation = (0 - rot_k) % 4)
            hm[xl + dxl:xh - dxh, yl + dyl:yh - dyh] |= world_aligned[dxl:V - dxh, dyl:V - dyh]
        glow = np.kron(hm.T, np.full((ts, ts), 255, dtype=np.uint16))[..., None]
        img = np.right_shift(img.astype(np.uint16) * 8 + glow * 2, 3).clip(0, 255).astype(np.uint8)  # base.py:326-329
    if show_agent_views:  # the agents' own observations, scaled by whole factors, stacked in columns (base.py:764-786)
        col_w = int(img.shape[0] * agent_col_width_frac - 2 * agent_col_padding_px)
        slot_h = (img.shape[1] - 2 * agent_col_padding_px) // max_agents_per_col
        views = []
        for v in agent_views_rgb(env, index):
            f = int(min(col_w / v.shape[0], slot_h / v.shape[1]))
            views.append(np.kron(v, np.ones((f, f, 1))))
        cols = []
        for c0 in range(0, A, max_agents_per_col):
            col = np.full((img.shape[0], col_w + 2 * agent_col_padding_px, 3), pad_grey, dtype=np.uint8)
            for k, v in enumerate(views[c0:c0 + max_agents_per_col]):
                oy = (slot_h - v.shape[1]) // 2 + agent_col_padding_px + k * slot_h
                ox = (col_w - v.shape[0]) // 2 + agent_col_padding_px
                col[oy:oy + v.shape[0], ox:ox + v.shape[1], :] = v
            cols.append(col)
        img = np.concatenate((img, *cols), axis=1)
    return img
