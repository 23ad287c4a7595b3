"""Runs the UNMODIFIED reference (kandouss/marlgrid at /root/reference) under import shims.

TEST INFRASTRUCTURE ONLY -- used in the dev container (where /root/reference exists) to
(1) validate the C restatement in oracle/mg_oracle.c and (2) generate the golden fixtures
under tests/golden/ (see oracle/gen_golden.py).  Nothing here is imported by the product.

What is patched, and why -- none of it edits reference source:
  * sys.path gets oracle/shims (gym, gym_minigrid.rendering, pyglet stand-ins) + /root/reference
  * numpy aliases np.bool/np.float/np.int (removed in numpy 2; used at base.py:424,467,510,741)
  * marlgrid.agents.occlude_mask (agents.py:298-343) is wrapped so its input grid lives in a
    zero-padded buffer: the numba function reads row j == V one element out of bounds
    (agents.py:304-306); "out-of-bounds bytes read as 0" is the canonical behaviour
  * MultiGrid.tile_cache (base.py:85,225-243) is pre-warmed with ACTIVE agents of every
    colour x dir so an inactive agent can never poison the cache with a black tile
  * env.np_random is replaced by the Philox contract object (oracle/philox.py)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("MARLGRID_REFERENCE", "/root/reference")

_ref = None


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "marlgrid"))


def load_reference():
    """Import the reference package (once) and apply the determinism patches."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError("reference source not found at %s" % REFERENCE_ROOT)
    for alias, typ in (("bool", bool), ("float", float), ("int", int)):
        if not hasattr(np, alias):
            setattr(np, alias, typ)
    shims = os.path.join(HERE, "shims")
    for p in (REFERENCE_ROOT, shims):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import marlgrid  # noqa: F401  (the reference package)
    import marlgrid.agents as ref_agents
    import marlgrid.base as ref_base
    import marlgrid.envs as ref_envs
    import marlgrid.objects as ref_objects

    assert os.path.abspath(ref_base.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), ref_base.__file__

    orig = ref_agents.occlude_mask

    def occlude_mask_zero_padded(grid, agent_pos):
        buf = np.zeros(grid.size + 64, dtype=np.bool_)
        g = buf[: grid.size].reshape(grid.shape)
        g[...] = grid
        return orig(g, agent_pos)

    occlude_mask_zero_padded.raw = orig
    ref_agents.occlude_mask = occlude_mask_zero_padded

    class Ref:
        pass

    _ref = Ref()
    _ref.agents = ref_agents
    _ref.base = ref_base
    _ref.envs = ref_envs
    _ref.objects = ref_objects
    _ref.occlude_mask_raw = orig
    return _ref


def prewarm_tile_cache(tile_size):
    """Render every colour x dir agent tile while ACTIVE (SURVEY.md A.5 cache hazard)."""
    ref = load_reference()
    for color in ref.objects.COLORS.keys():
        if color in ("shadow",):
            continue
        ag = ref.agents.GridAgentInterface(color=color, view_size=7, view_tile_size=tile_size)
        ag.activate()
        for d in range(4):
            ag.dir = d
            ref.base.MultiGrid.cache_render_obj(ag, tile_size, 3)


def make_env(env_id=None, env_class=None, seed=1337, env_index=0, agents=None, **env_kwargs):
    """Build a reference env and splice in the Philox np_random."""
    from .philox import PhiloxNpRandom

    ref = load_reference()
    if env_id is not None:
        import gym

        env = gym.make(env_id)
    else:
        cls = getattr(ref.envs, env_class)
        if env_class == "DoorKeyEnv" and not hasattr(cls, "_rand_int"):
            # doorkey.py:26,34 calls self._rand_int, a gym-minigrid MiniGridEnv method that MultiGridEnv does not have: supplied
            # here (`np_random.randint(low, high)`, gym-minigrid's definition) so that the otherwise unmodified class can be built
            cls._rand_int = lambda self, low, high: self.np_random.randint(low, high)
            # ... and its constructor ends with reset() -> gen_obs() -> MultiGrid.render, which raises NameError for the Key tile
            # (objects.py:309 names an undefined point_in_circle): on this subclass the observation is the encoded variant
            # (gen_obs_grid + MultiGrid.encode, base.py:418-451,196-214), the only one the reference can produce for this world
            cls.gen_agent_obs = lambda self, agent: (lambda g, v: g.encode(v))(*self.gen_obs_grid(agent))
        ag = [ref.agents.GridAgentInterface(**kw) for kw in agents]
        env = cls(agents=ag, **env_kwargs)
    for a in env.agents:
        prewarm_tile_cache(a.view_tile_size)
    env.np_random = PhiloxNpRandom(seed, env_index)
    return env


def ref_reset(env):
    env.np_random.begin_reset()
    return env.reset()


def ref_step(env, actions):
    env.np_random.begin_step()
    return env.step(list(int(a) for a in actions))


def encoded_obs(env):
    """The BASELINE 'encoded obs': gen_obs_grid + MultiGrid.encode (base.py:418-451,196-214)."""
    out = []
    for agent in env.agents:
        grid, vis = env.gen_obs_grid(agent)
        out.append(grid.encode(vis))
    return np.stack(out).astype(np.uint8)


def rgb_obs(env):
    """The reference's native observation (base.py:453-471); values 0..255 (dtype int64 there)."""
    out = []
    for agent in env.agents:
        o = env.gen_agent_obs(agent)
        if isinstance(o, dict):
            o = o["pov"]
        o = np.asarray(o)
        assert o.min() >= 0 and o.max() <= 255
        out.append(o.astype(np.uint8))
    return np.stack(out)


def vis_masks(env):
    return np.stack([env.gen_obs_grid(a)[1] for a in env.agents]).astype(np.uint8)


def extract_state(env):
    """Walk the reference's object graph into the SoA layout the device uses."""
    ref = load_reference()
    W, H = env.grid.width, env.grid.height
    A = len(env.agents)
    planes = np.zeros((3, W, H), dtype=np.uint8)
    ax = np.full(A, -1, np.int32)
    ay = np.full(A, -1, np.int32)
    rank = np.full(A, -1, np.int32)  # position in the cell's agent queue (0 = head)
    idx_of = {id(a): k for k, a in enumerate(env.agents)}
    for i in range(W):
        for j in range(H):
            obj = env.grid.get(i, j)
            if obj is None:
                continue
            if obj.is_agent:
                queue = [obj] + list(obj.agents)
            else:
                planes[:, i, j] = obj.encode()
                queue = list(obj.agents)
            for r, q in enumerate(queue):
                k = idx_of[id(q)]
                assert ax[k] == -1, "agent present twice"
                ax[k], ay[k], rank[k] = i, j, r
    st = {
        "grid": planes,
        "agent_x": ax,
        "agent_y": ay,
        "agent_rank": rank,
        "agent_dir": np.array([a.dir for a in env.agents], np.int32),
        "agent_active": np.array([bool(a.active) for a in env.agents], np.uint8),
        "agent_done": np.array([bool(a.done) for a in env.agents], np.uint8),
        "agent_carry": np.array(
            [a.carrying.encode() if a.carrying is not None else (0, 0, 0) for a in env.agents], np.int32
        ).reshape(A, 3),
        "step_count": np.int32(env.step_count),
    }
    for k, a in enumerate(env.agents):
        if a.pos is not None:
            assert (ax[k], ay[k]) == tuple(int(v) for v in a.pos), (k, a.pos, ax[k], ay[k])
        else:
            assert ax[k] == -1
    return st
