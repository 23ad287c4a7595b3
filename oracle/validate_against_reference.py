"""Lock-step validation of the C restatement (oracle/mg_oracle.c) against the UNMODIFIED reference.

Run in the dev container (needs /root/reference):   python -m oracle.validate_against_reference
TEST INFRASTRUCTURE ONLY.  For every scenario the reference env (under the shims, fed the Philox
contract draws) and the C oracle are stepped with the same actions; after every reset/step the
script compares: encoded obs, RGB obs, float64 rewards (bit pattern), done, grid planes, agent
pos/dir/active/done/carrying and the queue order of stacked agents.
"""
import sys
import time

import numpy as np

from marlgrid_b200 import atlas as product_atlas
from marlgrid_b200.config import GOAL_FIXED, GOAL_NONE, GOAL_RANDOM, make_config
from marlgrid_b200.objects import COLOR_TO_IDX, hide_mask

from . import mg_oracle as mo
from . import philox as px
from . import reference_harness as rh


def config_from_ref_env(env):
    """MgConfig for a constructed reference env (reads the reference's own attributes)."""
    cls = type(env).__mro__
    names = [c.__name__ for c in cls]
    ag = env.agents
    kw = dict(
        width=env.width, height=env.height,
        agent_colors=[a.color for a in ag],
        view_size=ag[0].view_size, view_offset=ag[0].view_offset, view_tile_size=ag[0].view_tile_size,
        max_steps=env.max_steps, ghost_mode=bool(env.ghost_mode), respawn=bool(env.respawn),
        reward_decay=bool(env.reward_decay), see_through_walls=bool(ag[0].see_through_walls),
        spawn_delay=[a.spawn_delay for a in ag],
        hide_types=hide_mask(getattr(ag[0], "hide_item_types", [])),
    )
    sk = dict(getattr(env, "agent_spawn_kwargs", {}) or {})  # place_obj(agent, **agent_spawn_kwargs), base.py:409-412
    assert set(sk) <= {"top", "size", "max_tries"}, sk
    kw.update(spawn_top=tuple(sk.get("top", (0, 0))), spawn_size=sk.get("size"), spawn_max_tries=sk.get("max_tries"))
    kw.update(prestige_beta=[a.prestige_beta for a in ag], prestige_scale=[a.prestige_scale for a in ag],
              allow_negative_prestige=[bool(a.allow_negative_prestige) for a in ag])
    if "DoorKeyEnv" in names:
        kw.update(goal_mode=GOAL_FIXED, scenario=1)
    elif "ClutteredGoalCycleEnv" in names:
        kw.update(goal_mode=GOAL_NONE, n_clutter=env.n_clutter, n_bonus_tiles=env.n_bonus_tiles,
                  bonus_reward=env.reward, bonus_penalty=env.penalty,
                  bonus_initial_reward=env.initial_reward, bonus_reset_on_mistake=env.reset_on_mistake)
    elif "ClutteredMultiGrid" in names:
        kw.update(goal_mode=GOAL_RANDOM if env.randomize_goal else GOAL_FIXED, n_clutter=env.n_clutter)
    elif "EmptyMultiGrid" in names:
        kw.update(goal_mode=GOAL_FIXED)
    else:
        raise ValueError(names)
    return make_config(**kw)


def reference_atlas(agent_colors, ts):
    """The tile atlas as rendered by the reference's own render_tile (base.py:275-299)."""
    ref = rh.load_reference()
    rh.prewarm_tile_cache(ts)
    MG = ref.base.MultiGrid
    O = ref.objects
    A = len(agent_colors)
    ags = []
    for c in agent_colors:
        row = []
        for d in range(4):
            a = ref.agents.GridAgentInterface(color=c, view_size=7, view_tile_size=ts)
            a.activate()
            a.dir = d
            row.append(a)
        ags.append(row)
    statics = [None, O.Wall(), O.Goal(color="green", reward=1), O.BonusTile(color="yellow", reward=1)]
    tiles = []
    for k, s in enumerate(statics):
        for slot in range(1 + 4 * A):
            if slot == 0:
                obj = s
                if obj is not None:
                    obj.agents = []
            else:
                q, d = (slot - 1) // 4, (slot - 1) % 4
                if s is None:
                    obj = ags[q][d]
                    obj.agents = []
                else:
                    obj = s
                    obj.agents = [ags[q][d]]
            tiles.append(np.asarray(MG.render_tile(obj, tile_size=ts, top_agent=None)).astype(np.uint8))
    return np.stack([np.stack([np.ascontiguousarray(ref.base.rotate_grid(t, k)) for k in range(4)]) for t in tiles])


class LockStep:
    def __init__(self, name, seed=1337, env_index=0, rgb=True, **mk):
        self.name = name
        self.env = rh.make_env(seed=seed, env_index=env_index, **mk)
        self.cfg = config_from_ref_env(self.env)
        self.ob = mo.OracleBatch(self.cfg, 1, seed=seed, env_offset=env_index)
        # the reference's agent dirs persist from construction (0); so do the oracle's (zero init)
        self.rgb = rgb
        if not rgb:
            # encoded-obs variant of the reference: its step() always renders RGB
            # (base.py:457-460), which raises NameError for Key/Ball/Door tiles (objects.py:309,321,370)
            env = self.env
            env.gen_agent_obs = lambda agent: (lambda g, v: g.encode(v))(*env.gen_obs_grid(agent))
        self.atlas = product_atlas.build_atlas([COLOR_TO_IDX[a.color] for a in self.env.agents], self.cfg.view_tile_size) if rgb else None
        self.n_checked = 0
        self.inject = None  # optional callable(self) run after every reset (adds Key/Ball/Box/Door objects)
        self.events = {"reward>0": 0, "reward<0": 0, "stacked": 0, "agent_done": 0, "carrying": 0, "raised": 0, "episodes_done": 0}
        self.trace = {"actions": [], "enc": [], "rgb": [], "rew": [], "done": [], "reset_after": []}

    def compare(self, tag):
        env, ob = self.env, self.ob
        st = rh.extract_state(env)
        W, H = self.cfg.width, self.cfg.height
        pl = ob.planes()[0]
        assert np.array_equal(pl, st["grid"]), f"{self.name} {tag}: grid planes differ\n{pl[0].T}\n{st['grid'][0].T}"
        placed = (ob.agent_flags[0] & 1).astype(bool)
        assert np.array_equal(placed, st["agent_x"] >= 0), f"{self.name} {tag}: placed {placed} {st['agent_x']}"
        assert np.array_equal(ob.agent_x[0][placed], st["agent_x"][placed]) and np.array_equal(ob.agent_y[0][placed], st["agent_y"][placed]), \
            f"{self.name} {tag}: pos"
        assert np.array_equal(ob.agent_dir[0], st["agent_dir"]), f"{self.name} {tag}: dir"
        assert np.array_equal((ob.agent_flags[0] >> 1) & 1, st["agent_active"]), f"{self.name} {tag}: active"
        assert np.array_equal((ob.agent_flags[0] >> 2) & 1, st["agent_done"]), f"{self.name} {tag}: done flags"
        assert np.array_equal(ob.agent_rank()[0], st["agent_rank"]), f"{self.name} {tag}: queue order {ob.agent_rank()[0]} vs {st['agent_rank']}"
        assert np.array_equal(ob.agent_carry[0].astype(np.int32), st["agent_carry"]), f"{self.name} {tag}: carry"
        assert int(ob.step_count[0]) == int(st["step_count"]), f"{self.name} {tag}: step_count"
        pr = np.array([float(a.prestige) for a in env.agents], np.float64)
        assert np.array_equal(ob.prestige[0].view(np.uint64), pr.view(np.uint64)), f"{self.name} {tag}: prestige {ob.prestige[0]} vs {pr}"
        enc_ref = rh.encoded_obs(env)
        enc = ob.obs_encode()[0]
        assert np.array_equal(enc, enc_ref), f"{self.name} {tag}: encoded obs differ at {np.argwhere(enc != enc_ref)[:5]}"
        if self.rgb:
            rgb_ref = rh.rgb_obs(env)
            rgb = ob.obs_rgb(self.atlas)[0]
            assert np.array_equal(rgb, rgb_ref), f"{self.name} {tag}: rgb obs differ at {np.argwhere(rgb != rgb_ref)[:5]}"
            self.trace["rgb"].append(rgb_ref)
        self.trace["enc"].append(enc_ref)
        self.n_checked += len(env.agents)
        self.events["stacked"] += int((st["agent_rank"] > 0).sum())
        self.events["agent_done"] += int(st["agent_done"].sum())
        self.events["carrying"] += int((st["agent_carry"][:, 0] > 0).sum())

    def reset(self):
        rh.ref_reset(self.env)
        self.ob.reset()
        if self.inject is not None:
            self.inject(self)
        self.compare("reset")

    def put_static(self, x, y, obj):
        """Place a reference WorldObj on an empty cell of BOTH worlds (grid.set, base.py:149-152)."""
        assert self.env.grid.get(x, y) is None
        self.env.grid.set(x, y, obj)
        self.ob.planes()[0][:, x, y] = obj.encode()

    def step(self, actions, t):
        err_before = int(self.ob.err[0])
        try:
            o, r, d, _ = rh.ref_step(self.env, actions)
        except (TypeError, ValueError, AssertionError, RecursionError, AttributeError) as exc:
            # the reference raised mid-step: the device contract is an error bit (include/marlgrid_b200.h MG_ERR_*)
            self.ob.step(np.asarray(actions, np.int32)[None], autoreset=False)
            want = {TypeError: 8, ValueError: 1, AssertionError: 4, RecursionError: 2, AttributeError: 32}[type(exc)]
            got = int(self.ob.err[0])
            assert got & want, f"{self.name} step {t}: reference raised {type(exc).__name__} but oracle err={got}"
            self.events["raised"] += 1
            # the reference aborted mid-step, the oracle completed it: agent dirs (which survive the
            # reset that follows, agents.py:161-170) are re-synchronised from the reference
            self.ob.agents[0, :, 2] = [a.dir for a in self.env.agents]
            self.last_raise = (np.asarray(actions, np.int32), want)
            return "raised"
        rew, done = self.ob.step(np.asarray(actions, np.int32)[None], autoreset=False)
        assert int(self.ob.err[0]) == err_before, f"{self.name} step {t}: oracle flagged err {int(self.ob.err[0])} but the reference did not raise"
        r = np.asarray(r, np.float64)
        self.events["reward>0"] += int((r > 0).sum())
        self.events["reward<0"] += int((r < 0).sum())
        assert np.array_equal(rew[0].view(np.uint64), r.view(np.uint64)), f"{self.name} step {t}: rewards {rew[0]} vs {r}"
        assert bool(done[0]) == bool(d), f"{self.name} step {t}: done {done[0]} vs {d}"
        self.trace["actions"].append(np.asarray(actions, np.int32))
        self.trace["rew"].append(r)
        self.trace["done"].append(bool(d))
        self.compare(f"step {t}")
        return bool(d)

    def run(self, episodes, steps, rng, p_forward=0.5, n_actions=7):
        for ep in range(episodes):
            self.reset()
            for t in range(steps):
                A = len(self.env.agents)
                act = rng.randint(0, n_actions, size=A)
                fw = rng.rand(A) < p_forward
                act[fw] = 2
                d = self.step(act, t)
                if d == "raised":
                    self.ob.envrec[0, 3] &= 0xFFFF  # clear error bits, abandon the episode
                    break
                self.trace["reset_after"].append(d)
                if d:
                    self.events["episodes_done"] += 1
                    break


def agents_cfg(n, colors=("red", "blue", "purple", "orange", "olive", "pink"), **kw):
    return [dict(color=colors[i % len(colors)], **{"view_size": 7, "view_tile_size": 8, **kw}) for i in range(n)]


# added in round 2 (own RNG stream in gen_golden.py, so the older fixtures stay reproducible): agent_spawn_kwargs (base.py:346,
# 409-412,505,642 -> place_obj(top, size, max_tries) base.py:690-699) and the DoorKey generator (doorkey.py:15-41; the class is
# made constructible by supplying the `_rand_int` it calls, see reference_harness.make_env)
EXTRA = [
    dict(name="Empty-spawnbox", env_class="EmptyMultiGrid", agents=agents_cfg(3), grid_size=9, agent_spawn_kwargs=dict(top=(1, 1), size=(3, 4))),
    dict(name="Cluttered-spawnbox-respawn", env_class="ClutteredMultiGrid", agents=agents_cfg(3), grid_size=6, n_clutter=2, respawn=True,
         max_steps=90, agent_spawn_kwargs=dict(top=(2, -2), size=(9, 5), max_tries=500)),
    dict(name="Empty-spawnbox-delay", env_class="EmptyMultiGrid", agents=agents_cfg(3, spawn_delay=4), grid_size=8, max_steps=40,
         agent_spawn_kwargs=dict(top=(2, 2), size=(2, 2))),
    dict(name="Empty6x6-prestige", env_class="EmptyMultiGrid", grid_size=6, max_steps=60, respawn=True,
         agents=[dict(color="prestige", view_size=7, view_tile_size=8), dict(color="red", view_size=7, view_tile_size=8),
                 dict(color="prestige", view_size=7, view_tile_size=8, prestige_beta=0.9, prestige_scale=0.7)]),
    dict(name="Goalcycle-prestige-ts11", env_class="ClutteredGoalCycleEnv", grid_size=9, n_clutter=4, n_bonus_tiles=3, penalty=-0.5, max_steps=80,
         agents=[dict(color="prestige", view_size=5, view_tile_size=11, view_offset=1), dict(color="prestige", view_size=5, view_tile_size=11, view_offset=1, prestige_scale=1.0)]),
    dict(name="Empty5x5-prestige-negative (AttributeError)", env_class="EmptyMultiGrid", grid_size=5, max_steps=40,
         agents=[dict(color="prestige", view_size=5, view_tile_size=8, allow_negative_prestige=True)]),
    dict(name="DoorKey8x8x2", env_class="DoorKeyEnv", agents=agents_cfg(2), grid_size=8, max_steps=120, interactive=True, no_inject=True),
    dict(name="DoorKey6x6x1", env_class="DoorKeyEnv", agents=agents_cfg(1), grid_size=6, max_steps=80, interactive=True, no_inject=True),
    dict(name="human_player.py config x3 agents", env_class="ClutteredGoalCycleEnv", agents=agents_cfg(3, view_offset=1, view_tile_size=11), grid_size=13, max_steps=250, clutter_density=0.15,
         respawn=True, ghost_mode=True, reward_decay=False, n_bonus_tiles=3, initial_reward=True, penalty=-1.5),  # examples/human_player.py:33-55 (encoded / RGB views)
]

SCENARIOS = [
    dict(name="2AgentEmpty9x9", env_id="MarlGrid-2AgentEmpty9x9-v0"),
    dict(name="3AgentEmpty9x9", env_id="MarlGrid-3AgentEmpty9x9-v0"),
    dict(name="4AgentEmpty9x9", env_id="MarlGrid-4AgentEmpty9x9-v0"),
    dict(name="3AgentCluttered11x11", env_id="MarlGrid-3AgentCluttered11x11-v0"),
    dict(name="3AgentCluttered15x15", env_id="MarlGrid-3AgentCluttered15x15-v0"),
    dict(name="1AgentCluttered(11x11,V5)", env_id="MarlGrid-1AgentCluttered15x15-v0"),
    dict(name="Empty5x5x4-crowded", env_class="EmptyMultiGrid", agents=agents_cfg(4), grid_size=5, max_steps=40),
    dict(name="Empty6x6x6-crowded", env_class="EmptyMultiGrid", agents=agents_cfg(6), grid_size=6, max_steps=60),
    dict(name="Cluttered9x9x3-dense", env_class="ClutteredMultiGrid", agents=agents_cfg(3), grid_size=9, clutter_density=0.2),
    dict(name="Cluttered-randgoal", env_class="ClutteredMultiGrid", agents=agents_cfg(3), grid_size=10, n_clutter=8, randomize_goal=True),
    dict(name="Cluttered-rect-12x8", env_class="ClutteredMultiGrid", agents=agents_cfg(2), width=12, height=8, n_clutter=6),
    dict(name="Empty-noghost", env_class="EmptyMultiGrid", agents=agents_cfg(4), grid_size=6, ghost_mode=False, max_steps=50),
    dict(name="Empty-nodecay", env_class="EmptyMultiGrid", agents=agents_cfg(2), grid_size=6, reward_decay=False, max_steps=30),
    dict(name="Empty-offset1", env_class="EmptyMultiGrid", agents=agents_cfg(3, view_offset=1), grid_size=8),
    dict(name="Empty-seethrough", env_class="EmptyMultiGrid", agents=agents_cfg(2, see_through_walls=True), grid_size=8),
    dict(name="Empty-respawn", env_class="EmptyMultiGrid", agents=agents_cfg(3), grid_size=5, respawn=True, max_steps=60),
    dict(name="Empty-spawndelay", env_class="EmptyMultiGrid", agents=[dict(color="red", view_size=7, view_tile_size=8), dict(color="blue", view_size=7, view_tile_size=8, spawn_delay=3), dict(color="purple", view_size=7, view_tile_size=8, spawn_delay=7)], grid_size=6, max_steps=40),
    dict(name="Cluttered-hide-walls", env_class="ClutteredMultiGrid", agents=agents_cfg(3, hide_item_types=["Wall"]), grid_size=8, n_clutter=6, max_steps=40),
    dict(name="Empty-hide-goal-agents", env_class="EmptyMultiGrid", agents=agents_cfg(4, hide_item_types=["Goal", "Agent"]), grid_size=5, max_steps=50),
    dict(name="Goalcycle-demo-solo", env_id="Goalcycle-demo-solo-v0"),
    dict(name="Goalcycle-3agents", env_class="ClutteredGoalCycleEnv", agents=agents_cfg(3, view_offset=1), grid_size=9, clutter_density=0.1, n_bonus_tiles=3, penalty=-1.5, respawn=True, max_steps=80),
    dict(name="Goalcycle-noinit-reset", env_class="ClutteredGoalCycleEnv", agents=agents_cfg(2), grid_size=7, n_clutter=2, n_bonus_tiles=4, penalty=0.25, reward=2, initial_reward=False, reset_on_mistake=True, max_steps=80),
]


INTERACTIVE = [
    dict(name="Empty8x8x3+keys/doors/balls", env_class="EmptyMultiGrid", agents=agents_cfg(3), grid_size=8, interactive=True, max_steps=150),
    dict(name="Empty7x7x4-noghost+objects", env_class="EmptyMultiGrid", agents=agents_cfg(4), grid_size=7, ghost_mode=False, interactive=True, max_steps=150),
    dict(name="Empty7x7x2+box (TypeError)", env_class="EmptyMultiGrid", agents=agents_cfg(2), grid_size=7, interactive=True, with_box=True, max_steps=150),
    dict(name="Empty8x8x4+objects, hidden Wall/Agent/Key/Door", env_class="EmptyMultiGrid", agents=agents_cfg(4, hide_item_types=["Wall", "Agent", "Key", "Door"]), grid_size=8, interactive=True, max_steps=120),
    dict(name="Empty6x6x4 crowded, hidden Agent", env_class="EmptyMultiGrid", agents=agents_cfg(4, hide_item_types=["Agent"]), grid_size=6, max_steps=60),
    dict(name="Empty6x6x2 bad action (ValueError)", env_class="EmptyMultiGrid", agents=agents_cfg(2), grid_size=6, n_actions=8, max_steps=30),
]


def inject_interactive(ls):
    """Scatter Key/Ball/Box/Door objects so pickup/drop/toggle (base.py:590-613) do something."""
    O = rh.load_reference().objects
    rng = np.random.RandomState(ls.env.np_random.episode)
    free = [(x, y) for x in range(1, ls.cfg.width - 1) for y in range(1, ls.cfg.height - 1) if ls.env.grid.get(x, y) is None]
    rng.shuffle(free)
    objs = [O.Key("blue"), O.Key("red"), O.Ball("green"), O.Door("blue", 3), O.Door("red", 2), O.Door("yellow", 1), O.Ball("purple")]
    if getattr(ls, "with_box", False):
        objs.append(O.Box(3))
    for (x, y), obj in zip(free, objs):
        ls.put_static(x, y, obj)


def validate_los(n=20000, seed=0):
    ref = rh.load_reference()
    rng = np.random.RandomState(seed)
    total = 0
    for V, positions in ((7, [(3, 6), (3, 5)]), (5, [(2, 4), (2, 3)]), (3, [(1, 2)]), (8, [(4, 7), (4, 6)])):
        for (ax, ay) in positions:
            dens = rng.rand(n, 1, 1)
            t = (rng.rand(n, V, V) > dens * 0.6)
            got = mo.los_batch(t.astype(np.uint8), ax, ay)
            for i in range(n):
                want = ref.agents.occlude_mask(np.ascontiguousarray(t[i]), (ax, ay))
                if not np.array_equal(got[i].astype(bool), want):
                    raise AssertionError(f"LOS mismatch V={V} pos={(ax, ay)} case {i}\n{t[i].astype(int)}\n{got[i]}\n{want.astype(int)}")
            total += n
    return total


def validate_philox():
    rng = np.random.RandomState(1)
    for _ in range(2000):
        ctr = rng.randint(0, 2 ** 32, size=4, dtype=np.uint64)
        key = rng.randint(0, 2 ** 32, size=2, dtype=np.uint64)
        assert px.philox4x32_10(ctr, key) == mo.philox(ctr, key)
    for A in range(1, 9):
        for t in range(200):
            assert px.shuffle_perm(1337 + A, 12345678901 + t, t, A) == mo.order(1337 + A, 12345678901 + t, t, A)


def validate_atlas():
    n = 0
    for ts in (8, 5, 11, 32):
        colors = ["red", "blue", "purple", "orange", "olive", "pink"]
        ra = reference_atlas(colors, ts)
        pa = product_atlas.build_atlas([COLOR_TO_IDX[c] for c in colors], ts)
        assert ra.shape == pa.shape, (ra.shape, pa.shape)
        assert np.array_equal(ra, pa), f"atlas ts={ts} differs in tiles {sorted(set(np.argwhere(ra != pa)[:, 0]))}"
        n += ra.shape[0] * 4
    return n


def main(quick=False):
    t0 = time.time()
    validate_philox()
    print("philox: python == C on 2000 random (ctr,key) + 1600 shuffles")
    n = validate_atlas()
    print(f"atlas: product build_atlas == reference render_tile on {n} oriented tiles (ts=8,5,11,32)")
    n = validate_los(2000 if quick else 20000)
    print(f"LOS: C restatement == numba occlude_mask (zero-padded) on {n} grids")
    rng = np.random.RandomState(7)
    total = 0
    for sc in SCENARIOS + INTERACTIVE + EXTRA:
        sc = dict(sc)
        name = sc.pop("name")
        interactive = sc.pop("interactive", False)
        with_box = sc.pop("with_box", False)
        no_inject = sc.pop("no_inject", False)
        n_actions = sc.pop("n_actions", 7)
        ls = LockStep(name, seed=1337 + total, env_index=total, rgb=not interactive, **sc)
        if interactive and not no_inject:
            ls.inject = inject_interactive
            ls.with_box = with_box
        ls.run(episodes=3 if quick else 12, steps=ls.cfg.max_steps + 5, rng=rng, p_forward=0.35 if interactive else 0.5, n_actions=n_actions)
        total += ls.n_checked
        ev = " ".join(f"{k}={v}" for k, v in ls.events.items() if v)
        print(f"  {name:32s} ok  ({ls.n_checked} agent-observations; {ev})")
    print(f"lock-step: {total} agent-observations (encoded+RGB), rewards bits, done, state: 0 mismatches  [{time.time() - t0:.0f}s]")


if __name__ == "__main__":
    main(quick="--quick" in sys.argv)
