"""MgConfig: the POD description of one env family handed to the C ABI (include/marlgrid_b200.h).

Collects the reference's constructor kwargs -- MultiGridEnv (marlgrid/base.py:335-347), the
scenario classes (marlgrid/envs/empty.py, cluttered.py:9-20, goalcycle.py:9-28) and the
GridAgentInterface geometry (marlgrid/agents.py:19-35) -- into one ctypes structure.
"""
import ctypes

from .objects import COLOR_TO_IDX, T_BONUS, T_GOAL, T_WALL

MG_MAX_AGENTS = 8
MG_MAX_VIEW = 8
MG_AGENT_REC = 16
MG_ENV_REC = 16

F_GHOST, F_RESPAWN, F_REWARD_DECAY, F_SEE_THROUGH, F_BONUS_INITIAL, F_BONUS_RESET = 1, 2, 4, 8, 16, 32
GOAL_NONE, GOAL_FIXED, GOAL_RANDOM = 0, 1, 2

ERR_BAD_ACTION, ERR_PLACEMENT, ERR_STACK, ERR_TOGGLE, ERR_RENDER, ERR_PRESTIGE = 1, 2, 4, 8, 16, 32
AF_PLACED, AF_ACTIVE, AF_DONE = 1, 2, 4


class MgConfig(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32),
        ("height", ctypes.c_int32),
        ("n_agents", ctypes.c_int32),
        ("view_size", ctypes.c_int32),
        ("view_offset", ctypes.c_int32),
        ("view_tile_size", ctypes.c_int32),
        ("max_steps", ctypes.c_int32),
        ("n_clutter", ctypes.c_int32),
        ("n_bonus_tiles", ctypes.c_int32),
        ("goal_mode", ctypes.c_int32),
        ("flags", ctypes.c_uint32),
        ("plane_stride", ctypes.c_int32),
        ("goal_reward", ctypes.c_double),
        ("bonus_reward", ctypes.c_double),
        ("bonus_penalty", ctypes.c_double),
        ("agent_color", ctypes.c_uint8 * MG_MAX_AGENTS),
        ("spawn_delay", ctypes.c_int32 * MG_MAX_AGENTS),
        ("n_static_kinds", ctypes.c_uint8),
        ("kind_of_type", ctypes.c_uint8 * 15),
        ("hide_types", ctypes.c_uint32),
        ("spawn_top", ctypes.c_int32 * 2),
        ("spawn_size", ctypes.c_int32 * 2),
        ("spawn_max_tries", ctypes.c_int32),
        ("scenario", ctypes.c_int32),
        ("prestige_mask", ctypes.c_uint32),
        ("prestige_neg_mask", ctypes.c_uint32),
        ("prestige_beta", ctypes.c_double * MG_MAX_AGENTS),
        ("prestige_scale", ctypes.c_double * MG_MAX_AGENTS),
    ]

    def describe(self):
        return {f: (list(getattr(self, f)) if hasattr(getattr(self, f), "__len__") else getattr(self, f)) for f, _ in self._fields_}


class MgState(ctypes.Structure):
    _fields_ = [
        ("grid", ctypes.c_void_p),
        ("agents", ctypes.c_void_p),
        ("envrec", ctypes.c_void_p),
        ("cellbits", ctypes.c_void_p),
        ("n_envs", ctypes.c_int64),
        ("env_offset", ctypes.c_int64),
        ("seed", ctypes.c_uint64),
        ("pregen", ctypes.c_void_p),
        ("prestige", ctypes.c_void_p),
    ]


def plane_stride_for(width, height):
    """Bytes per plane per env: W*H rounded up to the 16-byte bulk-copy granularity."""
    return (width * height + 15) // 16 * 16


def make_config(
    width,
    height,
    agent_colors,
    view_size=7,
    view_offset=0,
    view_tile_size=8,
    max_steps=100,
    n_clutter=0,
    n_bonus_tiles=0,
    goal_mode=GOAL_FIXED,
    ghost_mode=True,
    respawn=False,
    reward_decay=True,
    see_through_walls=False,
    goal_reward=1.0,
    bonus_reward=1.0,
    bonus_penalty=0.0,
    bonus_initial_reward=True,
    bonus_reset_on_mistake=False,
    spawn_delay=None,
    hide_types=0,
    spawn_top=(0, 0),
    spawn_size=None,
    spawn_max_tries=None,
    scenario=0,
    prestige_beta=None,
    prestige_scale=None,
    allow_negative_prestige=None,
):
    n_agents = len(agent_colors)
    if not (1 <= n_agents <= MG_MAX_AGENTS):
        raise ValueError(f"n_agents must be in 1..{MG_MAX_AGENTS}, got {n_agents}")
    if not (3 <= view_size <= MG_MAX_VIEW):
        raise ValueError(f"view_size must be in 3..{MG_MAX_VIEW}, got {view_size}")
    if width < 3 or height < 3:
        raise ValueError("Grid needs width, height >= 3")  # marlgrid/base.py:98-99
    if width > 255 or height > 255:
        raise ValueError("width/height must fit a byte")
    if (2 * int(max_steps) + 1) * n_agents >= 65536:
        raise ValueError("(2*max_steps+1)*n_agents must stay below 65536: queue-arrival stamps are 16 bits wide (include/marlgrid_b200.h)")
    cfg = MgConfig()
    cfg.width, cfg.height, cfg.n_agents = int(width), int(height), n_agents
    cfg.view_size, cfg.view_offset, cfg.view_tile_size = int(view_size), int(view_offset), int(view_tile_size)
    cfg.max_steps, cfg.n_clutter, cfg.n_bonus_tiles = int(max_steps), int(n_clutter), int(n_bonus_tiles)
    cfg.goal_mode = int(goal_mode)
    flags = 0
    flags |= F_GHOST if ghost_mode else 0
    flags |= F_RESPAWN if respawn else 0
    flags |= F_REWARD_DECAY if bool(reward_decay) else 0
    flags |= F_SEE_THROUGH if see_through_walls else 0
    flags |= F_BONUS_INITIAL if bool(bonus_initial_reward) else 0
    flags |= F_BONUS_RESET if bonus_reset_on_mistake else 0
    cfg.flags = flags
    cfg.plane_stride = plane_stride_for(width, height)
    cfg.goal_reward, cfg.bonus_reward, cfg.bonus_penalty = float(goal_reward), float(bonus_reward), float(bonus_penalty)
    for i, c in enumerate(agent_colors):
        cfg.agent_color[i] = c if isinstance(c, int) else COLOR_TO_IDX[c]
    sd = list(spawn_delay) if spawn_delay is not None else [0] * n_agents
    for i in range(n_agents):
        cfg.spawn_delay[i] = int(sd[i])
    # RGB atlas kinds: 1 = Wall('worst'), 2 = Goal('green'), 3 = BonusTile('yellow'); every other
    # static type has a render() that raises in the reference (objects.py:236-277,298-395).
    for t in range(15):
        cfg.kind_of_type[t] = 0xFF
    cfg.kind_of_type[0] = 0
    cfg.kind_of_type[T_WALL] = 1
    cfg.kind_of_type[T_GOAL] = 2
    cfg.kind_of_type[T_BONUS] = 3
    cfg.n_static_kinds = 3
    cfg.hide_types = int(hide_types)
    # agent_spawn_kwargs (base.py:409-412 -> place_obj(top, size, max_tries), :690-696)
    top = (max(int(spawn_top[0]), 0), max(int(spawn_top[1]), 0))
    size = (int(width), int(height)) if spawn_size is None else (int(spawn_size[0]), int(spawn_size[1]))
    bottom = (min(top[0] + size[0], int(width)), min(top[1] + size[1], int(height)))
    if bottom[0] <= top[0] or bottom[1] <= top[1]:
        raise ValueError("agent_spawn_kwargs: empty spawn region (np_random.randint(top, bottom) with low >= high)")
    cfg.spawn_top[0], cfg.spawn_top[1] = top
    cfg.spawn_size[0], cfg.spawn_size[1] = (0, 0) if spawn_size is None else size
    cfg.spawn_max_tries = 0 if spawn_max_tries is None else int(max(1, min(int(spawn_max_tries), 100000)))
    cfg.scenario = int(scenario)
    pmask = nmask = 0
    for i, c in enumerate(agent_colors):
        if (c if not isinstance(c, int) else None) == "prestige" or c == COLOR_TO_IDX["prestige"]:
            pmask |= 1 << i
        cfg.prestige_beta[i] = 0.95 if prestige_beta is None else float(prestige_beta[i])
        cfg.prestige_scale[i] = 2.0 if prestige_scale is None else float(prestige_scale[i])
        if allow_negative_prestige is not None and allow_negative_prestige[i]:
            nmask |= 1 << i
    cfg.prestige_mask, cfg.prestige_neg_mask = pmask, nmask
    return cfg


def n_tiles(cfg):
    return (cfg.n_static_kinds + 1) * (1 + 4 * cfg.n_agents)
