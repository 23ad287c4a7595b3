"""BatchedMultiGridEnv: N independent MarlGrid envs stepped on one B200 by hand-written CUDA kernels.

Host-side mirror of the reference's env runtime, MultiGridEnv (marlgrid/base.py:334-653): same
constructor kwargs, `seed`, `reset`, `step`, `action_space`, `observation_space`, `num_agents`; the
per-cell Python loops are replaced by calls through the C ABI (include/marlgrid_b200.h) on
structure-of-arrays torch tensors that never leave the device.

    env.reset()            -> obs  uint8 [B, A, V, V, 3]            (obs_mode='encoded', MultiGrid.encode base.py:196-214)
                                   uint8 [B, A, V*ts, V*ts, 3]      (obs_mode='rgb',     MultiGrid.render base.py:301-331)
    env.step(actions[B,A]) -> (obs, rewards float64 [B, A], done bool [B], {})

With `autoreset=True` (default) envs whose episode ended are reset inside the same kernel launch and
the returned obs are the first obs of the new episode, which is what a caller loop
`obs, r, d, _ = env.step(a); if d: obs = env.reset()` produces with the reference.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import atlas as _atlas
from .config import (AF_ACTIVE, AF_DONE, AF_PLACED, ERR_BAD_ACTION, ERR_PLACEMENT, ERR_PRESTIGE, ERR_RENDER, ERR_STACK, ERR_TOGGLE,
                     MgConfig, MgState, n_tiles)
from .spaces import Box, Discrete, Tuple

# raw accessors (no torch.cuda.Stream / device objects on the per-step path)
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) or (lambda i: torch.cuda.current_stream(i).cuda_stream)
_cuda_get_device = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device


def compose_rich_obs(pov, agents, width, height, observe_rewards=True, observe_position=True, observe_orientation=True):
    """The reference's observation_style='rich' dict (marlgrid/base.py:461-471), batched: from the observation tensor
    `pov` [B, A, ...] and the agent records `agents` uint8 [B, A, 16] (x, y, dir, flags, ...).

    * 'reward' is 0: the reference reads `agent.step_reward`, which `step` only ever sets to 0 (base.py:464,519);
    * 'position' = pos / (width, height) as float64, (0, 0) for an agent that is not in the grid (base.py:466-467);
    * 'orientation' = the agent's dir (base.py:469-470).
    """
    out = {"pov": pov}
    if observe_rewards:
        out["reward"] = torch.zeros(agents.shape[:2], dtype=torch.int64, device=agents.device)
    if observe_position:
        placed = (agents[..., 3] & AF_PLACED) != 0
        xy = agents[..., 0:2].to(torch.float64) * placed[..., None]
        out["position"] = xy / torch.tensor([width, height], dtype=torch.float64, device=agents.device)
    if observe_orientation:
        out["orientation"] = agents[..., 2].to(torch.int64)
    return out


class BatchedMultiGridEnv:
    metadata = {}

    def __init__(self, cfg, num_envs=1, device="cuda", seed=1337, env_offset=0, obs_mode="encoded", autoreset=True,
                 check_errors=False, obs_buffers=2, pregen=True):
        if not isinstance(cfg, MgConfig):
            raise TypeError("cfg must be a marlgrid_b200.config.MgConfig")
        if obs_mode not in ("encoded", "rgb"):
            raise ValueError("obs_mode must be 'encoded' or 'rgb'")
        self._lib = _lib.load()  # raises if the CUDA extension is not built: there is no CPU path
        self.cfg = cfg
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("marlgrid_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.obs_mode = obs_mode
        self.autoreset = bool(autoreset)
        self.check_errors_each_step = bool(check_errors)
        self.env_offset = int(env_offset)
        _lib.check(self._lib.mg_config_validate(ctypes.byref(cfg)), "mg_config_validate")

        B, A, V, ts, S = self.num_envs, cfg.n_agents, cfg.view_size, cfg.view_tile_size, cfg.plane_stride
        dev = self.device
        self.grid = torch.empty((B, 3, S), dtype=torch.uint8, device=dev)
        self.agent_rec = torch.empty((B, A, 16), dtype=torch.uint8, device=dev)  # per-agent records (the reference keeps this state on env.agents[i])
        self.envrec = torch.empty((B, 4), dtype=torch.int32, device=dev)
        # derived bit-planes (include/marlgrid_b200.h), tile-transposed: word w of env e at [e // 32, w, e % 32]
        self.cellbits = torch.zeros(((B + 31) // 32, 44, 32), dtype=torch.int32, device=dev)
        # pre-generated next worlds, filled by the library's background generator (include/marlgrid_b200.h MgState.pregen)
        self.pregen = torch.zeros((B, 64), dtype=torch.int32, device=dev) if pregen else None
        # GridAgentInterface.prestige (agents.py:141-153): kept only for families with a 'prestige'-coloured agent, whose tile is
        # recoloured from it (agents.py:92-119); such families take the per-env step kernel + observe kernel
        self.prestige = torch.zeros((B, A), dtype=torch.float64, device=dev) if cfg.prestige_mask else None
        self.rewards = torch.zeros((B, A), dtype=torch.float64, device=dev)
        self.done = torch.zeros((B,), dtype=torch.bool, device=dev)  # the kernels write 0/1 bytes: no conversion pass per step
        # The observation tensor returned by step() is a view of a device buffer.  With obs_buffers = 2 (default) steps
        # alternate between two buffers, so the previous step's observation stays valid while the next one exists -- the
        # reference's `save_step(obs, act, next_obs, ...)` pattern works without copies; obs_buffers = 1 halves the memory.
        if obs_buffers not in (1, 2):
            raise ValueError("obs_buffers must be 1 or 2")
        obs_shape = (B, A, V, V, 3) if obs_mode == "encoded" else (B, A, V * ts, V * ts, 3)
        self._obs_bufs = [torch.zeros(obs_shape, dtype=torch.uint8, device=dev) for _ in range(obs_buffers)]
        self._obs_idx = 0
        self.obs = self._obs_bufs[0]
        if obs_mode == "encoded":
            self.atlas = None
        else:
            at = _atlas.build_atlas([int(c) for c in cfg.agent_color[:A]], ts, cfg.n_static_kinds)
            assert at.shape[0] == n_tiles(cfg)
            self.atlas = torch.from_numpy(at).to(dev)
        self._state = MgState()
        self._fast = None
        self._n_act = B * A
        self.seed(seed)
        self._sync_state_struct()
        with torch.cuda.device(dev):
            _lib.check(self._lib.mg_init(ctypes.byref(cfg), ctypes.byref(self._state), self._stream()), "mg_init")
            # MultiGridEnv.__init__ ends with self.reset() (base.py:368): a freshly constructed env can be stepped at once.
            # The episode counter (the Philox counter of the placement draws) is rewound afterwards, so an explicit
            # reset() as first call regenerates the same first world instead of consuming a second set of draws.
            _lib.check(self._lib.mg_reset(ctypes.byref(cfg), ctypes.byref(self._state), None, self._stream()), "mg_reset")
            self.envrec[:, 1] = 0

    def __del__(self):
        # a pass of the background world generator may still be reading / writing this env's tensors on the side stream
        try:
            if getattr(self, "pregen", None) is not None:
                self._lib.mg_pregen_drain()
        except Exception:  # noqa: BLE001 -- interpreter shutdown
            pass

    # ---- plumbing ------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _sync_state_struct(self):
        st = self._state
        st.grid, st.agents, st.envrec = self.grid.data_ptr(), self.agent_rec.data_ptr(), self.envrec.data_ptr()
        st.cellbits = self.cellbits.data_ptr()
        st.pregen = self.pregen.data_ptr() if self.pregen is not None else None
        st.prestige = self.prestige.data_ptr() if self.prestige is not None else None
        st.n_envs, st.env_offset, st.seed = self.num_envs, self.env_offset, self._seed

    def seed(self, seed=1337):
        """marlgrid/base.py:371-374.  The Philox key; the global env index is the counter."""
        self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        if getattr(self, "pregen", None) is not None:
            self._lib.mg_pregen_drain()  # no generator pass keyed with the old seed may still be writing world slots
        self._state.seed = self._seed
        return [seed]

    # ---- gym surface ---------------------------------------------------------------------------
    @property
    def num_agents(self):
        return self.cfg.n_agents

    @property
    def agents(self):
        """The per-agent interface objects, like the reference's `env.agents` (base.py:353,392-400).  Their dynamic state
        (pos / dir / carrying / active / done, agents.py:155-170) lives in the batched record tensor `env.agent_rec`."""
        ai = getattr(self, "agent_interfaces", None)
        if ai is None:  # constructed from a bare MgConfig: interfaces with the config's geometry
            from .agents import GridAgentInterface
            from .objects import IDX_TO_COLOR

            c = self.cfg
            ai = [GridAgentInterface(view_size=c.view_size, view_tile_size=c.view_tile_size, view_offset=c.view_offset,
                                     see_through_walls=bool(c.flags & 8), spawn_delay=int(c.spawn_delay[i]),
                                     color=IDX_TO_COLOR[int(c.agent_color[i])]) for i in range(c.n_agents)]
            self.agent_interfaces = ai
        return ai

    @property
    def action_space(self):
        """base.py:376-380: Tuple of the agents' action spaces (Discrete(7), or Discrete(3) with restrict_actions)."""
        return Tuple([a.action_space for a in self.agents])

    @property
    def observation_space(self):
        """base.py:382-386: Tuple of the agents' observation spaces (a Dict per agent for observation_style='rich').  With
        obs_mode='encoded' the image part is MultiGrid.encode's (V, V, 3) array instead of the RGB view."""
        if self.obs_mode == "rgb":
            return Tuple([a.observation_space for a in self.agents])
        shape = tuple(self.obs.shape[2:])
        out = []
        for a in self.agents:
            enc = Box(0, 255, shape, np.uint8)
            if a.observation_style == "rich":
                from .spaces import Dict

                out.append(Dict({**a.observation_space.spaces, "pov": enc}))
            else:
                out.append(enc)
        return Tuple(out)

    def _observe(self):
        cfg, st = ctypes.byref(self.cfg), ctypes.byref(self._state)
        if self.obs_mode == "encoded":
            _lib.check(self._lib.mg_obs_encode(cfg, st, self.obs.data_ptr(), self._stream()), "mg_obs_encode")
        else:
            _lib.check(self._lib.mg_obs_rgb(cfg, st, self.atlas.data_ptr(), self.obs.data_ptr(), self._stream()), "mg_obs_rgb")
        return self.obs

    def reset(self, mask=None, **kwargs):
        """Start a new episode in every env (or where mask[b] != 0); returns the observations."""
        with torch.cuda.device(self.device):
            m = None
            if mask is not None:
                m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            _lib.check(self._lib.mg_reset(ctypes.byref(self.cfg), ctypes.byref(self._state), m.data_ptr() if m is not None else None,
                                          self._stream()), "mg_reset")
            return self._observe()

    def _actions(self, actions):
        a = actions
        if not (torch.is_tensor(a) and a.dtype == torch.int32 and a.device == self.device and a.is_contiguous()
                and a.numel() == self.num_envs * self.cfg.n_agents):
            a = torch.as_tensor(actions, device=self.device)
            if a.dtype != torch.int32:
                a = a.to(torch.int32)
            a = a.reshape(self.num_envs, self.cfg.n_agents).contiguous()
        return a

    def _prepare_fast_step(self):
        """Everything of a step call that does not change from step to step, resolved once: a Python loop around
        env.step() is host-bound long before the 16 us kernel is (ctypes argument conversion, context managers)."""
        L = self._lib
        cfg, st = ctypes.byref(self.cfg), ctypes.byref(self._state)
        rew, done = ctypes.c_void_p(self.rewards.data_ptr()), ctypes.c_void_p(self.done.data_ptr())
        tails = []
        for buf in self._obs_bufs:  # one argument tail per observation buffer
            obs = ctypes.c_void_p(buf.data_ptr())
            tails.append((rew, done, obs) if self.obs_mode == "encoded" else (rew, done, ctypes.c_void_p(self.atlas.data_ptr()), obs))
        fn = L.mg_step_fused if self.obs_mode == "encoded" else L.mg_step_fused_rgb
        self._fast = (fn, cfg, st, tails, self.device.index)

    def step(self, actions):
        """One env.step for the whole batch (marlgrid/base.py:501-653) in a single kernel launch.

        The host side of this call is trimmed to the bone (a Python loop around env.step() is host-bound long before the
        kernel is): argument tuple resolved once, raw stream handle instead of a torch.cuda.Stream object."""
        a = actions
        if not (type(a) is torch.Tensor and a.dtype is torch.int32 and a.is_cuda and a.is_contiguous() and a.numel() == self._n_act
                and a.device == self.device):
            a = self._actions(actions)
        fast = self._fast
        if fast is None:
            self._prepare_fast_step()
            fast = self._fast
        fn, cfg, st, tails, dev_index = fast
        idx = self._obs_idx = (self._obs_idx + 1) % len(tails)
        self.obs = self._obs_bufs[idx]
        if _cuda_get_device() != dev_index:
            with torch.cuda.device(self.device):
                rc = fn(cfg, st, a.data_ptr(), *tails[idx], self.autoreset, _raw_stream(dev_index))
        else:
            rc = fn(cfg, st, a.data_ptr(), *tails[idx], self.autoreset, _raw_stream(dev_index))
        if rc:
            _lib.check(rc, "mg_step_fused")
        if self.check_errors_each_step:
            self.check_errors()
        return self.obs, self.rewards, self.done, {}

    def step_only(self, actions):
        """step without producing observations (mg_step); returns (rewards, done)."""
        with torch.cuda.device(self.device):
            a = self._actions(actions)
            _lib.check(self._lib.mg_step(ctypes.byref(self.cfg), ctypes.byref(self._state), a.data_ptr(), self.rewards.data_ptr(),
                                         self.done.data_ptr(), int(self.autoreset), self._stream()), "mg_step")
            return self.rewards, self.done

    def observe(self):
        with torch.cuda.device(self.device):
            return self._observe()

    def render(self, index=0, mode="rgb_array", **kwargs):
        """Whole-grid view of env `index` (marlgrid/base.py:714-795): uint8 image [H*32, W*32 + agent-view columns, 3].
        Same keyword arguments as the reference (highlight, tile_size, show_agent_views, ...); there is no window: every
        mode returns the array."""
        from . import render as _render

        return _render.render(self, index=index, **kwargs)

    def sync_derived(self):
        """Recompute the derived device state (occupancy bitboards, queue-head flags) after the planes or
        the agent records were edited from Python (e.g. `env.planes[...] = ...`)."""
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mg_sync_derived(ctypes.byref(self.cfg), ctypes.byref(self._state), self._stream()), "mg_sync_derived")

    def rollout(self, actions):
        """actions int32 [T, B, A]: T fused steps enqueued from C; returns the last (obs, rewards, done)."""
        with torch.cuda.device(self.device):
            if self.obs_mode != "encoded":
                raise NotImplementedError("rollout() drives the encoded-obs path")
            a = torch.as_tensor(actions, device=self.device).to(torch.int32).contiguous()
            T = a.shape[0]
            assert a.shape[1:] == (self.num_envs, self.cfg.n_agents)
            _lib.check(self._lib.mg_rollout_fused(ctypes.byref(self.cfg), ctypes.byref(self._state), a.data_ptr(), T, self.rewards.data_ptr(),
                                                  self.done.data_ptr(), self.obs.data_ptr(), int(self.autoreset), self._stream()), "mg_rollout_fused")
            return self.obs, self.rewards, self.done

    def rollout_all(self, actions, out=None):
        """On-device rollout loop (mg_rollout_persistent): actions int32 [T, B, A] -> (obs [T, B, A, V, V, 3], rewards
        [T, B, A], done [T, B]) of EVERY step; one kernel launch when the batch fits the resident CTAs (the tiles' state
        then stays in shared memory between the steps).  `out` = (obs, rewards, done) tensors to fill."""
        with torch.cuda.device(self.device):
            if self.obs_mode != "encoded":
                raise NotImplementedError("rollout_all() drives the encoded-obs path")
            a = torch.as_tensor(actions, device=self.device).to(torch.int32).contiguous()
            T, B, A, V = a.shape[0], self.num_envs, self.cfg.n_agents, self.cfg.view_size
            assert a.shape[1:] == (B, A)
            if out is None:
                out = (torch.empty((T, B, A, V, V, 3), dtype=torch.uint8, device=self.device),
                       torch.empty((T, B, A), dtype=torch.float64, device=self.device), torch.empty((T, B), dtype=torch.bool, device=self.device))
            obs, rew, done = out
            _lib.check(self._lib.mg_rollout_persistent(ctypes.byref(self.cfg), ctypes.byref(self._state), a.data_ptr(), T, rew.data_ptr(),
                                                       done.data_ptr(), obs.data_ptr(), int(self.autoreset), self._stream()), "mg_rollout_persistent")
            if T:
                self.obs.copy_(obs[-1]); self.rewards.copy_(rew[-1]); self.done.copy_(done[-1])
            return obs, rew, done

    def rollout_policy(self, policy, first_actions, n_steps, out=None):
        """Closed-loop rollout on the device (mg_rollout_policy; the loop of the reference's README.md:43-57 without the host):
        step 0 plays `first_actions` int32 [B, A], step t + 1 the actions `policy` (marlgrid_b200.policy.LinearPolicy) chooses
        from step t's observations.  Returns (obs [T, B, A, V, V, 3], rewards [T, B, A], done [T, B], actions [T, B, A]) of
        every step; one kernel launch when the batch fits the resident CTAs of the rollout kernel.  `out` = the four tensors."""
        with torch.cuda.device(self.device):
            if self.obs_mode != "encoded":
                raise NotImplementedError("rollout_policy() drives the encoded-obs path")
            T, B, A, V = int(n_steps), self.num_envs, self.cfg.n_agents, self.cfg.view_size
            if policy.n_agents != A or policy.view_size != V:
                raise ValueError(f"policy is for {policy.n_agents} agents with view {policy.view_size}, env has {A} / {V}")
            if out is None:
                out = (torch.empty((T, B, A, V, V, 3), dtype=torch.uint8, device=self.device), torch.empty((T, B, A), dtype=torch.float64, device=self.device),
                       torch.empty((T, B), dtype=torch.bool, device=self.device), torch.empty((T, B, A), dtype=torch.int32, device=self.device))
            obs, rew, done, act = out
            assert act.shape == (T, B, A) and act.dtype == torch.int32 and act.is_contiguous()
            if T:
                act[0].copy_(torch.as_tensor(first_actions, device=self.device).to(torch.int32).reshape(B, A))
            pol, keep = policy.device_struct(self.device)
            _lib.check(self._lib.mg_rollout_policy(ctypes.byref(self.cfg), ctypes.byref(self._state), ctypes.byref(pol), T, act.data_ptr(), rew.data_ptr(),
                                                   done.data_ptr(), obs.data_ptr(), int(self.autoreset), self._stream()), "mg_rollout_policy")
            if T:
                self.obs.copy_(obs[-1]); self.rewards.copy_(rew[-1]); self.done.copy_(done[-1])
            return obs, rew, done, act

    def policy_act(self, policy, obs=None, out=None):
        """Actions int32 [B, A] the LinearPolicy chooses from `obs` (default: the observations of the last step / reset) --
        `agents.action_step(obs)` of the reference's loop as one kernel launch (mg_policy_act)."""
        with torch.cuda.device(self.device):
            if self.obs_mode != "encoded":
                raise NotImplementedError("LinearPolicy reads encoded observations")
            obs = self.obs if obs is None else obs
            assert obs.dtype == torch.uint8 and obs.is_contiguous() and obs.numel() == self.obs.numel()
            if out is None:
                out = torch.empty((self.num_envs, self.cfg.n_agents), dtype=torch.int32, device=self.device)
            pol, keep = policy.device_struct(self.device)
            _lib.check(self._lib.mg_policy_act(ctypes.byref(self.cfg), ctypes.byref(self._state), ctypes.byref(pol), obs.data_ptr(), out.data_ptr(), self._stream()), "mg_policy_act")
            return out

    def random_actions(self, counter, n_actions=7, seed=0, out=None):
        """Uniform synthetic policy on the device; `counter` selects the draw."""
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty((self.num_envs, self.cfg.n_agents), dtype=torch.int32, device=self.device)
            _lib.check(self._lib.mg_random_actions(out.data_ptr(), out.numel(), n_actions, seed, counter, self._stream()), "mg_random_actions")
            return out

    # ---- errors the reference would have raised --------------------------------------------------
    @property
    def err(self):
        return (self.envrec[:, 3] >> 16) & 0xFFFF

    def check_errors(self, clear=True):
        """Device code cannot raise; it accumulates MG_ERR_* bits per env.  Raise like the reference."""
        bits = int(np.bitwise_or.reduce(self.err.cpu().numpy())) if self.num_envs else 0
        if bits and clear:
            self.envrec[:, 3] &= 0xFFFF
        if bits & ERR_BAD_ACTION:
            raise ValueError("Environment can't handle action (marlgrid/base.py:619-620)")
        if bits & ERR_PLACEMENT:
            raise RecursionError("Rejection sampling failed in place_obj. (marlgrid/base.py:706)")
        if bits & ERR_STACK:
            raise AssertionError("agent left a cell that cannot be overlapped (marlgrid/base.py:558)")
        if bits & ERR_TOGGLE:
            raise TypeError("Box.toggle() takes 1 positional argument but 3 were given (marlgrid/objects.py:381)")
        if bits & ERR_PRESTIGE:
            raise AttributeError("'GridAgentInterface' object has no attribute 'rew' (allow_negative_prestige, marlgrid/agents.py:146-148)")
        if bits & ERR_RENDER:
            raise NameError("object has no working render() in the reference (marlgrid/objects.py:274-277,309-321,370)")

    # ---- SoA views (decoded) -------------------------------------------------------------------
    @property
    def planes(self):
        """uint8 [B, 3, W, H] view: type / colour / state planes, index [x][y] like MultiGrid.grid (base.py:91).
        After writing through this view call sync_derived()."""
        W, H = self.cfg.width, self.cfg.height
        return self.grid[:, :, : W * H].unflatten(2, (W, H))

    @property
    def grid_type(self):
        return self.planes[:, 0]

    @property
    def grid_color(self):
        return self.planes[:, 1]

    @property
    def grid_state(self):
        return self.planes[:, 2]

    @property
    def agent_pos(self):
        return self.agent_rec[:, :, 0:2]

    @property
    def agent_dir(self):
        return self.agent_rec[:, :, 2]

    @property
    def agent_flags(self):
        """MG_AF_* bits (bit 7 of the stored byte is device-derived queue-head state and is masked off)."""
        return self.agent_rec[:, :, 3] & 0x7F

    @property
    def agent_placed(self):
        return (self.agent_rec[:, :, 3] & AF_PLACED) != 0

    @property
    def agent_active(self):
        return (self.agent_rec[:, :, 3] & AF_ACTIVE) != 0

    @property
    def agent_done(self):
        return (self.agent_rec[:, :, 3] & AF_DONE) != 0

    @property
    def agent_carrying(self):
        return self.agent_rec[:, :, 4:7]

    @property
    def agent_stamp(self):
        return self.agent_rec[:, :, 8:12].contiguous().view(torch.int32)[..., 0]

    @property
    def step_count(self):
        return self.envrec[:, 0]

    @property
    def episode(self):
        return self.envrec[:, 1]

    # ---- checkpoint / resume (the RNG is counter-based: resume is exact) -----------------------
    def state_dict(self):
        sd = {"grid": self.grid.clone(), "agents": self.agent_rec.clone(), "envrec": self.envrec.clone(), "seed": self._seed,
              "env_offset": self.env_offset}
        if self.prestige is not None:
            sd["prestige"] = self.prestige.clone()
        return sd

    def load_state_dict(self, sd):
        self.grid.copy_(sd["grid"])
        self.agent_rec.copy_(sd["agents"])
        self.envrec.copy_(sd["envrec"])
        if self.prestige is not None and "prestige" in sd:
            self.prestige.copy_(sd["prestige"])
        if self.pregen is not None:
            self._lib.mg_pregen_drain()  # episode numbers may go backwards: no generator pass may be in flight (mg_pregen.cu)
            self.pregen.zero_()
        self._seed = int(sd["seed"])
        self.env_offset = int(sd["env_offset"])
        self._sync_state_struct()
        self.sync_derived()

    def unbatched(self, index=0):
        """gym-style view of env `index`: lists of per-agent obs, like the reference's return values."""
        return UnbatchedView(self, index)

    def __repr__(self):
        c = self.cfg
        return f"<{type(self).__name__} {c.width}x{c.height} agents={c.n_agents} num_envs={self.num_envs} obs={self.obs_mode} on {self.device}>"


class UnbatchedView:
    """The reference's single-env surface on top of a batched env (requires num_envs == 1 to step)."""

    def __init__(self, env, index=0):
        self.env = env
        self.index = index

    @property
    def num_agents(self):
        return self.env.num_agents

    @property
    def action_space(self):
        return self.env.action_space

    @property
    def observation_space(self):
        return self.env.observation_space

    def seed(self, seed=1337):
        return self.env.seed(seed)

    def _per_agent(self, obs):
        n = self.env.num_agents
        if isinstance(obs, dict):  # observation_style='rich': one dict per agent, like the reference
            return [{k: v[self.index, a] for k, v in obs.items()} for a in range(n)]
        return [obs[self.index, a] for a in range(n)]

    def reset(self, **kw):
        return self._per_agent(self.env.reset())

    def render(self, mode="rgb_array", **kwargs):
        return self.env.render(index=self.index, mode=mode, **kwargs)

    def step(self, actions):
        if self.env.num_envs != 1:
            raise ValueError("UnbatchedView.step needs num_envs == 1")
        if len(actions) != self.env.num_agents:
            raise AssertionError("len(actions) == len(self.agents)")  # base.py:508
        obs, rew, done, info = self.env.step(torch.as_tensor(list(int(a) for a in actions), dtype=torch.int32).view(1, -1))
        self.env.check_errors()
        return self._per_agent(obs), rew[0], bool(done[0].item()), info
