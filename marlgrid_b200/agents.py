"""Agent-side interface: GridAgentInterface (configuration), LearningAgent and IndependentLearners.

* GridAgentInterface mirrors the constructor of the reference class (marlgrid/agents.py:19-35): it is
  the per-agent *configuration* (view geometry, colour, spaces).  The dynamic state the reference
  keeps on the same object (pos/dir/carrying/active/done, agents.py:155-170) lives in the batched SoA
  `agents` tensor of the env instead (include/marlgrid_b200.h).
* IndependentLearners / LearningAgent do not exist in the reference's code at the surveyed commit;
  they are specified only by its README (README.md:21-64).  They are supplied here from that usage:
  `agents.action_step(obs_array)`, `agents.save_step(obs, act, next_obs, rew, done)`,
  `with agents.episode(): ...`, learners implementing action_step/save_step/start_episode/end_episode.
"""
import contextlib

import numpy as np
import torch

from .objects import Actions, COLOR_TO_IDX
from .spaces import Box, Dict, Discrete


class GridAgentInterface:
    actions = Actions  # marlgrid/agents.py:10-17 (an IntEnum: agent.actions.forward == 2)

    def __init__(
        self,
        view_size=7,
        view_tile_size=5,
        view_offset=0,
        observation_style="image",
        observe_rewards=False,
        observe_position=False,
        observe_orientation=False,
        restrict_actions=False,
        see_through_walls=False,
        hide_item_types=(),
        prestige_beta=0.95,
        prestige_scale=2,
        allow_negative_prestige=False,
        spawn_delay=0,
        color="red",
        **kwargs,
    ):
        if observation_style not in ("image", "rich"):
            raise ValueError(f"{type(self).__name__} kwarg 'observation_style' must be one of 'image', 'rich'.")  # agents.py:78
        if color not in COLOR_TO_IDX:
            raise ValueError(f"unknown colour {color!r}")
        self.view_size = view_size
        self.view_tile_size = view_tile_size
        self.view_offset = view_offset
        self.observation_style = observation_style
        self.observe_rewards = observe_rewards
        self.observe_position = observe_position
        self.observe_orientation = observe_orientation
        self.restrict_actions = restrict_actions
        self.see_through_walls = see_through_walls
        self.hide_item_types = list(hide_item_types)
        self.prestige_beta = prestige_beta if prestige_beta <= 1 else 0.95
        self.prestige_scale = prestige_scale
        self.allow_negative_prestige = allow_negative_prestige
        self.spawn_delay = spawn_delay
        self.color = color
        self.init_kwargs = kwargs
        image_space = Box(low=0, high=255, shape=(view_tile_size * view_size, view_tile_size * view_size, 3), dtype="uint8")
        if observation_style == "image":
            self.observation_space = image_space
        else:
            sp = {"pov": image_space}
            if observe_rewards:
                sp["reward"] = Box(low=-np.inf, high=np.inf, shape=(), dtype=np.float32)
            if observe_position:
                sp["position"] = Box(low=0, high=1, shape=(2,), dtype=np.float32)
            if observe_orientation:
                sp["orientation"] = Discrete(n=4)
            self.observation_space = Dict(sp)
        self.action_space = Discrete(3) if restrict_actions else Discrete(len(self.actions))
        self.metadata = {"color": color, "view_size": view_size, "view_tile_size": view_tile_size}

    def clone(self):
        return type(self)(
            view_size=self.view_size, view_tile_size=self.view_tile_size, view_offset=self.view_offset,
            observation_style=self.observation_style, observe_rewards=self.observe_rewards,
            observe_position=self.observe_position, observe_orientation=self.observe_orientation,
            restrict_actions=self.restrict_actions, see_through_walls=self.see_through_walls,
            hide_item_types=self.hide_item_types, prestige_beta=self.prestige_beta, prestige_scale=self.prestige_scale,
            allow_negative_prestige=self.allow_negative_prestige, spawn_delay=self.spawn_delay, color=self.color, **self.init_kwargs,
        )


class LearningAgent(GridAgentInterface):
    """Base class of README.md:21-25: subclass and implement action_step / save_step."""

    def action_step(self, obs):
        raise NotImplementedError

    def save_step(self, *transition_values):
        raise NotImplementedError

    def start_episode(self):
        pass

    def end_episode(self):
        pass


class IndependentLearners(list):
    """A list of learners that is also the `agents` argument of an env (README.md:29-36).

    obs_array / reward_array are indexed [agent] for a single env or [batch, agent] for a batched env;
    each learner receives its own slice (tensors stay on the device).
    """

    def __init__(self, *learners):
        super().__init__(learners)

    def _per_agent(self, x, batched):
        n = len(self)
        if isinstance(x, (list, tuple)):
            return [x[k] for k in range(n)]
        if torch.is_tensor(x) or isinstance(x, np.ndarray):
            if batched:
                return [x[:, k] for k in range(n)]
            return [x[k] for k in range(n)]
        return [x] * n

    def action_step(self, obs_array):
        batched = (torch.is_tensor(obs_array) or isinstance(obs_array, np.ndarray)) and obs_array.ndim == 5
        obs = self._per_agent(obs_array, batched)
        acts = [agent.action_step(o) for agent, o in zip(self, obs)]
        if batched:
            return torch.stack([torch.as_tensor(a) for a in acts], dim=1)
        return acts

    def save_step(self, obs, act, next_obs, rew, done):
        batched = (torch.is_tensor(obs) or isinstance(obs, np.ndarray)) and obs.ndim == 5
        o, a, n, r = (self._per_agent(v, batched) for v in (obs, act, next_obs, rew))
        d = self._per_agent(done, False) if isinstance(done, (list, tuple)) else [done] * len(self)
        for k, agent in enumerate(self):
            agent.save_step(o[k], a[k], n[k], r[k], d[k])

    def start_episode(self):
        for agent in self:
            if hasattr(agent, "start_episode"):
                agent.start_episode()

    def end_episode(self):
        for agent in self:
            if hasattr(agent, "end_episode"):
                agent.end_episode()

    @contextlib.contextmanager
    def episode(self):
        self.start_episode()
        try:
            yield self
        finally:
            self.end_episode()
