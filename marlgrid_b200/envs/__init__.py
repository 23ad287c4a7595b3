"""Scenario classes and the env registry -- drop-in for `marlgrid.envs` (marlgrid/envs/__init__.py).

Same class names, constructor kwargs, registered ids and `env_from_config` as the reference; every
constructor additionally accepts the batching arguments `num_envs`, `device`, `obs_mode`
('rgb' = the reference's native observation, 'encoded' = MultiGrid.encode), `env_offset`,
`autoreset`.  The scenario generators themselves (`_gen_grid`: marlgrid/envs/empty.py:9-16,
cluttered.py:25-36, goalcycle.py:30-51) run inside the reset kernel; the classes below only
translate kwargs into an MgConfig.
"""
import inspect
import random
import sys

from ..agents import GridAgentInterface
from ..config import GOAL_FIXED, GOAL_NONE, GOAL_RANDOM, make_config
from ..objects import hide_mask
from ..env import BatchedMultiGridEnv, compose_rich_obs

this_module = sys.modules[__name__]
registered_envs = []
registry = {}


class MultiGridEnv(BatchedMultiGridEnv):
    """Constructor surface of marlgrid/base.py:335-347 on top of the batched device env."""

    def __init__(
        self,
        agents=[],
        grid_size=None,
        width=None,
        height=None,
        max_steps=100,
        reward_decay=True,
        seed=1337,
        respawn=False,
        ghost_mode=True,
        agent_spawn_kwargs={},
        num_envs=1,
        device="cuda",
        obs_mode="rgb",
        env_offset=0,
        autoreset=True,
        check_errors=False,
        obs_buffers=2,
        pregen=True,
    ):
        if grid_size is not None:
            assert width is None and height is None  # base.py:349-351
            width, height = grid_size, grid_size
        # agent_spawn_kwargs (base.py:346): forwarded to place_obj for every agent placement -- at reset (base.py:409-412), after
        # a spawn delay (:505) and at respawn (:642).  `top` / `size` / `max_tries` (base.py:690-696) run inside the kernels;
        # `reject_fn` is a Python callable and cannot.
        spawn = dict(agent_spawn_kwargs or {})
        if "reject_fn" in spawn and spawn["reject_fn"] is not None:
            raise NotImplementedError("agent_spawn_kwargs['reject_fn'] (a Python callable per placement try) cannot run inside the reset kernel")
        spawn.pop("reject_fn", None)
        if set(spawn) - {"top", "size", "max_tries"}:
            raise TypeError(f"place_obj() got an unexpected keyword argument {sorted(set(spawn) - {'top', 'size', 'max_tries'})[0]!r}")  # base.py:690
        self.agent_spawn_kwargs = spawn
        self.agent_interfaces = []
        for a in agents:  # add_agent base.py:392-400
            if isinstance(a, dict):
                self.agent_interfaces.append(GridAgentInterface(**a))
            elif isinstance(a, GridAgentInterface):
                self.agent_interfaces.append(a)
            else:
                raise ValueError(
                    "To add an agent to a marlgrid environment, call add_agent with either a GridAgentInterface object "
                    " or a dictionary that can be used to initialize one.")
        if not self.agent_interfaces:
            raise ValueError("a batched MarlGrid env needs at least one agent")
        ai = self.agent_interfaces
        if len({tuple(sorted(a.hide_item_types)) for a in ai}) != 1:
            raise ValueError("all agents of a batched env must share hide_item_types (one mask per env family)")
        for field in ("view_size", "view_tile_size", "view_offset", "see_through_walls", "observation_style", "observe_rewards",
                      "observe_position", "observe_orientation"):
            if len({getattr(a, field) for a in ai}) != 1:
                raise ValueError(f"all agents of a batched env must share {field} (observations are one tensor)")
        self.width, self.height = width, height
        self.max_steps, self.reward_decay, self.respawn, self.ghost_mode = max_steps, reward_decay, respawn, ghost_mode
        cfg = make_config(
            width=width, height=height, agent_colors=[a.color for a in ai],
            view_size=ai[0].view_size, view_offset=ai[0].view_offset, view_tile_size=ai[0].view_tile_size,
            max_steps=max_steps, ghost_mode=ghost_mode, respawn=respawn, reward_decay=reward_decay,
            see_through_walls=ai[0].see_through_walls, spawn_delay=[a.spawn_delay for a in ai],
            hide_types=hide_mask(ai[0].hide_item_types),
            prestige_beta=[a.prestige_beta for a in ai], prestige_scale=[a.prestige_scale for a in ai],
            allow_negative_prestige=[a.allow_negative_prestige for a in ai],
            **{"spawn_top": tuple(spawn.get("top", (0, 0))), "spawn_size": spawn.get("size"), "spawn_max_tries": spawn.get("max_tries"),
               **self._scenario()},
        )
        super().__init__(cfg, num_envs=num_envs, device=device, seed=seed, env_offset=env_offset, obs_mode=obs_mode,
                         autoreset=autoreset, check_errors=check_errors, obs_buffers=obs_buffers, pregen=pregen)

    def _scenario(self):
        raise NotImplementedError

    # observation_style='rich' (marlgrid/base.py:461-471, agents.py:66-76): the observation becomes a dict of batched tensors
    def _style(self, obs):
        a0 = self.agent_interfaces[0]
        if a0.observation_style != "rich":
            return obs
        return compose_rich_obs(obs, self.agent_rec, self.width, self.height, a0.observe_rewards, a0.observe_position, a0.observe_orientation)

    def reset(self, *args, **kwargs):
        return self._style(super().reset(*args, **kwargs))

    def observe(self):
        return self._style(super().observe())

    def step(self, actions):
        obs, rew, done, info = super().step(actions)
        return self._style(obs), rew, done, info


class EmptyMultiGrid(MultiGridEnv):
    """marlgrid/envs/empty.py: border walls + green goal at (W-2, H-2)."""
    mission = "get to the green square"

    def _scenario(self):
        return dict(goal_mode=GOAL_FIXED)


class ClutteredMultiGrid(MultiGridEnv):
    """marlgrid/envs/cluttered.py: + n_clutter random interior walls (+ optionally a random goal)."""
    mission = "get to the green square"

    def __init__(self, *args, n_clutter=None, clutter_density=None, randomize_goal=False, **kwargs):
        if (n_clutter is None) == (clutter_density is None):
            raise ValueError("Must provide n_clutter xor clutter_density in environment config.")  # cluttered.py:10-11
        self._clutter = (n_clutter, clutter_density)
        self.randomize_goal = randomize_goal
        super().__init__(*args, **kwargs)

    def _scenario(self):
        n_clutter, density = self._clutter
        if density is not None:
            n_clutter = int(density * (self.width - 2) * (self.height - 2))  # cluttered.py:15-16
        self.n_clutter = n_clutter
        return dict(goal_mode=GOAL_RANDOM if self.randomize_goal else GOAL_FIXED, n_clutter=n_clutter)


class DoorKeyEnv(MultiGridEnv):
    """marlgrid/envs/doorkey.py: a vertical wall at a random column with a locked yellow door, the yellow key somewhere on the
    left, the goal in the bottom-right corner.  The reference's class calls `self._rand_int`, which its base class lacks, and its
    constructor renders a Key tile, whose render() raises: it cannot be instantiated there.  Here `_rand_int(lo, hi)` is
    `np_random.randint(lo, hi)` (gym-minigrid's definition) and observations are the encoded ones (Key / Door have no working
    RGB tile in the reference, objects.py:309,370: obs_mode='rgb' sets MG_ERR_RENDER like any such object)."""
    mission = "use the key to open the door and then get to the goal"

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("obs_mode", "encoded")
        super().__init__(*args, **kwargs)
        self.agent_spawn_kwargs = {}  # doorkey.py:40

    def _scenario(self):
        return dict(goal_mode=GOAL_FIXED, scenario=1, spawn_top=(0, 0), spawn_size=None, spawn_max_tries=None)


class ClutteredGoalCycleEnv(MultiGridEnv):
    """marlgrid/envs/goalcycle.py: clutter + n_bonus_tiles BonusTiles rewarded when visited in cyclic order."""
    mission = "Cycle between yellow goal tiles."

    def __init__(self, *args, reward=1, penalty=0.0, n_clutter=None, clutter_density=None, n_bonus_tiles=3, initial_reward=True,
                 cycle_reset=False, reset_on_mistake=False, reward_decay=False, **kwargs):
        if (n_clutter is None) == (clutter_density is None):
            raise ValueError("Must provide n_clutter xor clutter_density in environment config.")  # goalcycle.py:10-11
        self._clutter = (n_clutter, clutter_density)
        self.reward, self.penalty, self.initial_reward = reward, penalty, initial_reward
        self.n_bonus_tiles, self.reset_on_mistake = n_bonus_tiles, reset_on_mistake
        super().__init__(*args, **{**kwargs, "reward_decay": reward_decay})  # goalcycle.py:14

    def _scenario(self):
        n_clutter, density = self._clutter
        if density is not None:
            n_clutter = int(density * (self.width - 2) * (self.height - 2))
        self.n_clutter = n_clutter
        return dict(goal_mode=GOAL_NONE, n_clutter=n_clutter, n_bonus_tiles=self.n_bonus_tiles, bonus_reward=self.reward,
                    bonus_penalty=self.penalty, bonus_initial_reward=self.initial_reward, bonus_reset_on_mistake=self.reset_on_mistake)


def register_marl_env(env_name, env_class, n_agents, grid_size, view_size, view_tile_size=8, view_offset=0, agent_color=None,
                      env_kwargs={}):
    """marlgrid/envs/__init__.py:20-55.  (Like the reference, the registered agents always get
    view_tile_size=8 -- the argument is ignored there, envs/__init__.py:42.)"""
    colors = ["red", "blue", "purple", "orange", "olive", "pink"]
    assert n_agents <= len(colors)

    def factory(**overrides):
        agents = [
            GridAgentInterface(color=c if agent_color is None else agent_color, view_size=view_size, view_tile_size=8, view_offset=view_offset)
            for c in colors[:n_agents]
        ]
        return env_class(agents=agents, grid_size=grid_size, **{**env_kwargs, **overrides})

    env_class_name = f"env_{len(registered_envs)}"
    setattr(this_module, env_class_name, factory)
    registered_envs.append(env_name)
    registry[env_name] = factory
    try:  # also visible to gym.make when a gym is installed
        from gym.envs.registration import register as gym_register  # type: ignore

        gym_register(env_name, entry_point=f"marlgrid_b200.envs:{env_class_name}")
    except Exception:  # noqa: BLE001
        pass


def make(env_name, **overrides):
    """gym.make(env_name) stand-in; overrides are env kwargs such as num_envs=..., obs_mode='encoded'."""
    if env_name not in registry:
        raise KeyError(f"unknown env id {env_name!r}; registered: {registered_envs}")
    return registry[env_name](**overrides)


def env_from_config(env_config, randomize_seed=True):
    """marlgrid/envs/__init__.py:58-67."""
    possible_envs = {k: v for k, v in globals().items() if inspect.isclass(v) and issubclass(v, MultiGridEnv)}
    env_class = possible_envs[env_config["env_class"]]
    env_kwargs = {k: v for k, v in env_config.items() if k != "env_class"}
    if randomize_seed:
        env_kwargs["seed"] = env_kwargs.get("seed", 0) + random.randint(0, 1337 * 1337)
    return env_class(**env_kwargs)


# marlgrid/envs/__init__.py:70-121 (note: the '15x15' single-agent id really is 11x11 / view 5 there)
register_marl_env("MarlGrid-1AgentCluttered15x15-v0", ClutteredMultiGrid, n_agents=1, grid_size=11, view_size=5, env_kwargs={"n_clutter": 30})
register_marl_env("MarlGrid-3AgentCluttered11x11-v0", ClutteredMultiGrid, n_agents=3, grid_size=11, view_size=7, env_kwargs={"clutter_density": 0.15})
register_marl_env("MarlGrid-3AgentCluttered15x15-v0", ClutteredMultiGrid, n_agents=3, grid_size=15, view_size=7, env_kwargs={"clutter_density": 0.15})
register_marl_env("MarlGrid-2AgentEmpty9x9-v0", EmptyMultiGrid, n_agents=2, grid_size=9, view_size=7)
register_marl_env("MarlGrid-3AgentEmpty9x9-v0", EmptyMultiGrid, n_agents=3, grid_size=9, view_size=7)
register_marl_env("MarlGrid-4AgentEmpty9x9-v0", EmptyMultiGrid, n_agents=4, grid_size=9, view_size=7)
register_marl_env("Goalcycle-demo-solo-v0", ClutteredGoalCycleEnv, n_agents=1, grid_size=13, view_size=7, view_tile_size=5, view_offset=1,
                  env_kwargs={"clutter_density": 0.1, "n_bonus_tiles": 3})
