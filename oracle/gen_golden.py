"""Freeze outputs of the UNMODIFIED reference as golden fixtures under tests/golden/.

Run in the dev container (needs /root/reference):   python -m oracle.gen_golden
TEST INFRASTRUCTURE ONLY.  The reference cannot travel to the GPU box, so its behaviour on seeded
trajectories is committed as small .npz files together with this generating script:

  traj_<scenario>.npz   event stream (0 = reset, 1 = step(actions), 2 = overwrite planes, 3 = step on which
                        the reference raises: expect the MG_ERR_* bit, then take `dir` from the fixture) and, after
                        every event, what the reference shows: encoded obs, RGB obs (first episodes),
                        rewards (float64), done, grid planes, agent pos/dir/flags/queue-rank/carrying
  los.npz               transparency grids + occlude_mask results (numba, zero-padded input)
  atlas_ts<k>.npz       tiles rendered by the reference's MultiGrid.render_tile, all orientations
Every trajectory is produced by oracle.validate_against_reference.LockStep, i.e. the C oracle is
checked against the reference on exactly these trajectories while they are recorded.
"""
import json
import os

import numpy as np

from . import reference_harness as rh
from . import validate_against_reference as val

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class Recorder(val.LockStep):
    def __init__(self, *a, rgb_events=40, **kw):
        super().__init__(*a, **kw)
        self.ev = []
        self.rgb_events = rgb_events

    def _snap(self, kind, actions, rew, done):
        st = rh.extract_state(self.env)
        A = len(self.env.agents)
        rec = dict(
            kind=kind,
            actions=np.zeros(A, np.int32) if actions is None else np.asarray(actions, np.int32),
            rew=np.zeros(A, np.float64) if rew is None else np.asarray(rew, np.float64),
            done=bool(done),
            enc=rh.encoded_obs(self.env),
            grid=st["grid"], x=st["agent_x"], y=st["agent_y"], dir=st["agent_dir"],
            flags=((st["agent_x"] >= 0) * 1 + st["agent_active"] * 2 + st["agent_done"] * 4).astype(np.uint8),
            rank=st["agent_rank"], carry=st["agent_carry"], step_count=int(st["step_count"]),
        )
        if self.rgb and len(self.ev) < self.rgb_events:
            rec["rgb"] = rh.rgb_obs(self.env)
        self.ev.append(rec)

    def reset(self):
        rh.ref_reset(self.env)
        self.ob.reset()
        self._snap(0, None, None, False)
        if self.inject is not None:
            self.inject(self)
            self._snap(2, None, None, False)
        self.compare("reset")

    def step(self, actions, t):
        d = super().step(actions, t)
        if d != "raised":
            self._snap(1, actions, self.trace["rew"][-1], d)
        else:  # kind 3: a step on which the reference raises; only actions, err and the post-raise dirs are meaningful
            rec = dict(self.ev[-1])  # the reference's object graph may be inconsistent after the raise
            rec.pop("rgb", None)
            rec.update(kind=3, actions=np.asarray(actions, np.int32), err=self.last_raise[1],
                       dir=np.array([a.dir for a in self.env.agents], np.int32))
            self.ev.append(rec)
        return d

    def save(self, fname, make_kwargs):
        ev = self.ev
        n_rgb = sum(1 for e in ev if "rgb" in e)
        out = dict(
            meta=np.frombuffer(json.dumps(make_kwargs).encode(), dtype=np.uint8),
            kind=np.array([e["kind"] for e in ev], np.uint8),
            actions=np.stack([e["actions"] for e in ev]),
            rew=np.stack([e["rew"] for e in ev]),
            done=np.array([e["done"] for e in ev], np.uint8),
            enc=np.stack([e["enc"] for e in ev]),
            grid=np.stack([e["grid"] for e in ev]),
            x=np.stack([e["x"] for e in ev]), y=np.stack([e["y"] for e in ev]), dir=np.stack([e["dir"] for e in ev]),
            flags=np.stack([e["flags"] for e in ev]), rank=np.stack([e["rank"] for e in ev]), carry=np.stack([e["carry"] for e in ev]),
            step_count=np.array([e["step_count"] for e in ev], np.int32),
            err=np.array([e.get("err", 0) for e in ev], np.int32),
        )
        if n_rgb:
            out["rgb"] = np.stack([e["rgb"] for e in ev[:n_rgb]])
        np.savez_compressed(os.path.join(OUT, fname), **out)
        return len(ev)


def cfg_kwargs(cfg, seed, env_index):
    """make_config kwargs that rebuild `cfg` (JSON-able), plus the RNG identity of the run."""
    from marlgrid_b200.config import F_BONUS_INITIAL, F_BONUS_RESET, F_GHOST, F_RESPAWN, F_REWARD_DECAY, F_SEE_THROUGH

    return dict(
        seed=seed, env_index=env_index,
        config=dict(
            width=cfg.width, height=cfg.height, agent_colors=[int(c) for c in cfg.agent_color[: cfg.n_agents]],
            view_size=cfg.view_size, view_offset=cfg.view_offset, view_tile_size=cfg.view_tile_size,
            max_steps=cfg.max_steps, n_clutter=cfg.n_clutter, n_bonus_tiles=cfg.n_bonus_tiles, goal_mode=cfg.goal_mode,
            ghost_mode=bool(cfg.flags & F_GHOST), respawn=bool(cfg.flags & F_RESPAWN), reward_decay=bool(cfg.flags & F_REWARD_DECAY),
            see_through_walls=bool(cfg.flags & F_SEE_THROUGH), goal_reward=cfg.goal_reward, bonus_reward=cfg.bonus_reward,
            bonus_penalty=cfg.bonus_penalty, bonus_initial_reward=bool(cfg.flags & F_BONUS_INITIAL),
            bonus_reset_on_mistake=bool(cfg.flags & F_BONUS_RESET), spawn_delay=[int(s) for s in cfg.spawn_delay[: cfg.n_agents]],
            hide_types=int(cfg.hide_types),
            spawn_top=[int(cfg.spawn_top[0]), int(cfg.spawn_top[1])],
            spawn_size=None if (cfg.spawn_size[0] == 0 and cfg.spawn_size[1] == 0) else [int(cfg.spawn_size[0]), int(cfg.spawn_size[1])],
            spawn_max_tries=int(cfg.spawn_max_tries) or None, scenario=int(cfg.scenario),
            prestige_beta=[float(b) for b in cfg.prestige_beta[: cfg.n_agents]], prestige_scale=[float(b) for b in cfg.prestige_scale[: cfg.n_agents]],
            allow_negative_prestige=[bool((cfg.prestige_neg_mask >> i) & 1) for i in range(cfg.n_agents)],
        ),
    )


def gen_trajectories(only=None):
    """only: substring filter on the scenario name (the RNG / seed schedule of the other scenarios is unchanged: scenarios
    added later get their own stream below)."""
    rng = np.random.RandomState(2024)
    k = 0
    for sc in val.SCENARIOS + val.INTERACTIVE[:3]:
        if "hide" in sc["name"]:  # added after the first fixtures were frozen: own stream, so the older files stay reproducible
            continue
        if only is not None and only not in sc["name"]:
            k += 1
            continue
        sc = dict(sc)
        name = sc.pop("name")
        interactive = sc.pop("interactive", False)
        with_box = sc.pop("with_box", False)
        sc.pop("n_actions", None)
        seed, env_index = 1337 + 17 * k, 1000 * k + 3
        k += 1
        rec = Recorder(name, seed=seed, env_index=env_index, rgb=not interactive, **sc)
        if interactive:
            rec.inject = val.inject_interactive
            rec.with_box = with_box
        rec.run(episodes=3, steps=rec.cfg.max_steps + 3, rng=rng, p_forward=0.35 if interactive else 0.5)
        fname = "traj_" + "".join(ch if ch.isalnum() else "_" for ch in name).strip("_") + ".npz"
        n = rec.save(fname, cfg_kwargs(rec.cfg, seed, env_index))
        print(f"  {fname:48s} {n} events; {rec.events}")


def gen_hide_trajectories():
    """hide_item_types scenarios (base.py:441-449), recorded with their own RNG stream."""
    rng = np.random.RandomState(4048)
    for k, sc in enumerate(s for s in val.SCENARIOS if "hide" in s["name"]):
        sc = dict(sc)
        name = sc.pop("name")
        seed, env_index = 7001 + 13 * k, 500 * k + 9
        rec = Recorder(name, seed=seed, env_index=env_index, rgb=True, **sc)
        rec.run(episodes=3, steps=rec.cfg.max_steps + 3, rng=rng, p_forward=0.5)
        fname = "traj_" + "".join(ch if ch.isalnum() else "_" for ch in name).strip("_") + ".npz"
        n = rec.save(fname, cfg_kwargs(rec.cfg, seed, env_index))
        print(f"  {fname:48s} {n} events; {rec.events}")


def gen_extra_trajectories():
    """agent_spawn_kwargs and DoorKey scenarios (validate_against_reference.EXTRA), recorded with their own RNG stream."""
    rng = np.random.RandomState(90210)
    for k, sc in enumerate(val.EXTRA):
        sc = dict(sc)
        name = sc.pop("name")
        interactive = sc.pop("interactive", False)
        sc.pop("no_inject", None)
        seed, env_index = 9001 + 11 * k, 321 * k + 5
        rec = Recorder(name, seed=seed, env_index=env_index, rgb=not interactive, **sc)
        rec.run(episodes=4, steps=rec.cfg.max_steps + 3, rng=rng, p_forward=0.4)
        fname = "traj_" + "".join(ch if ch.isalnum() else "_" for ch in name).strip("_") + ".npz"
        n = rec.save(fname, cfg_kwargs(rec.cfg, seed, env_index))
        print(f"  {fname:48s} {n} events; {rec.events}")


def gen_los():
    ref = rh.load_reference()
    rng = np.random.RandomState(5)
    out = {}
    for V, positions in ((7, [(3, 6), (3, 5)]), (5, [(2, 4), (2, 3)])):
        for (ax, ay) in positions:
            n = 1500
            dens = rng.rand(n, 1, 1)
            t = rng.rand(n, V, V) > dens * 0.6
            # hand cases of SURVEY.md A.4: column-1 wall vs column-5 wall asymmetry, all transparent, all opaque
            t[0] = True
            t[1] = False
            t[2] = True; t[2][1, : V - 1] = False
            t[3] = True; t[3][V - 2, : V - 1] = False
            m = np.stack([ref.agents.occlude_mask(np.ascontiguousarray(t[i]), (ax, ay)) for i in range(n)])
            out[f"t_V{V}_{ax}_{ay}"] = np.packbits(t, axis=None)
            out[f"m_V{V}_{ax}_{ay}"] = np.packbits(m, axis=None)
            out[f"n_V{V}_{ax}_{ay}"] = np.array([n])
    np.savez_compressed(os.path.join(OUT, "los.npz"), **out)
    print("  los.npz")


def gen_atlas():
    colors = ["red", "blue", "purple", "orange", "olive", "pink"]
    for ts in (8, 5, 11):
        np.savez_compressed(os.path.join(OUT, f"atlas_ts{ts}.npz"), atlas=val.reference_atlas(colors, ts),
                            colors=np.array([val.COLOR_TO_IDX[c] for c in colors], np.uint8))
        print(f"  atlas_ts{ts}.npz")


class RenderRecorder(val.LockStep):
    """Short runs that freeze the reference's whole-grid view env.render(mode='rgb_array') (base.py:714-795) after every event."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.ev = []
        rh.prewarm_tile_cache(32)  # TILE_PIXELS: the same cache-poisoning hazard as for the agent views

    def _snap(self, kind, actions):
        A = len(self.env.agents)
        self.ev.append(dict(kind=kind, actions=np.zeros(A, np.int32) if actions is None else np.asarray(actions, np.int32),
                            img=np.asarray(self.env.render(mode="rgb_array")).astype(np.uint8)))

    def reset(self):
        rh.ref_reset(self.env)
        self.ob.reset()
        self._snap(0, None)
        self.compare("reset")

    def step(self, actions, t):
        d = super().step(actions, t)
        self._snap(1, actions)
        return d


def gen_rich():
    """rich_<scenario>.npz: the observation dicts of observation_style='rich' agents (base.py:461-471: pov, reward, position,
    orientation) after a reset and every step, recorded from the reference."""
    rng = np.random.RandomState(31337)
    agents = [dict(color=c, view_size=7, view_tile_size=8, observation_style="rich", observe_rewards=True, observe_position=True,
                   observe_orientation=True) for c in ("red", "blue", "purple")]
    seed, env_index = 2718, 44
    env = rh.make_env(env_class="EmptyMultiGrid", agents=agents, grid_size=7, max_steps=30, seed=seed, env_index=env_index)
    ev = []

    def snap(kind, actions, obs):
        ev.append(dict(kind=kind, actions=np.zeros(3, np.int32) if actions is None else np.asarray(actions, np.int32),
                       pov=np.stack([np.asarray(o["pov"]).astype(np.uint8) for o in obs]), reward=np.array([o["reward"] for o in obs], np.float64),
                       position=np.stack([np.asarray(o["position"], np.float64) for o in obs]), orientation=np.array([o["orientation"] for o in obs], np.int64)))

    snap(0, None, rh.ref_reset(env))
    for t in range(70):
        act = rng.randint(0, 7, size=3)
        act[rng.rand(3) < 0.5] = 2
        obs, rew, done, _ = rh.ref_step(env, act)
        snap(1, act, obs)
        if done:
            snap(0, None, rh.ref_reset(env))
    cfg = val.config_from_ref_env(env)
    np.savez_compressed(os.path.join(OUT, "rich_Empty7x7x3.npz"), meta=np.frombuffer(json.dumps(cfg_kwargs(cfg, seed, env_index)).encode(), dtype=np.uint8),
                        **{k: np.stack([e[k] for e in ev]) for k in ("kind", "actions", "pov", "reward", "position", "orientation")})
    print(f"  rich_Empty7x7x3.npz                              {len(ev)} events")


def gen_render_extra():
    """render_<scenario>.npz for scenarios of validate_against_reference.EXTRA (own RNG stream): 'prestige' agents in the whole-grid view."""
    rng = np.random.RandomState(1177)
    picks = {"Empty6x6-prestige": 40}
    for k, sc in enumerate(s for s in val.EXTRA if s["name"] in picks):
        sc = dict(sc)
        name = sc.pop("name")
        seed, env_index = 515 + 5 * k, 19 * k + 2
        rec = RenderRecorder(name, seed=seed, env_index=env_index, rgb=True, **sc)
        rec.reset()
        for t in range(picks[name]):
            act = rng.randint(0, 7, size=len(rec.env.agents))
            act[rng.rand(len(act)) < 0.6] = 2
            if rec.step(act, t) is True:
                rec.reset()
        fname = "render_" + "".join(ch if ch.isalnum() else "_" for ch in name).strip("_") + ".npz"
        np.savez_compressed(os.path.join(OUT, fname), meta=np.frombuffer(json.dumps(cfg_kwargs(rec.cfg, seed, env_index)).encode(), dtype=np.uint8),
                            kind=np.array([e["kind"] for e in rec.ev], np.uint8), actions=np.stack([e["actions"] for e in rec.ev]),
                            img=np.stack([e["img"] for e in rec.ev]))
        print(f"  {fname:48s} {len(rec.ev)} frames {rec.ev[0]['img'].shape}; {rec.events}")


def gen_render():
    """render_<scenario>.npz: event stream (0 = reset, 1 = step(actions)) + the reference's rendered frame after every event."""
    rng = np.random.RandomState(77)
    picks = {"3AgentCluttered11x11": 12, "4AgentEmpty9x9": 10, "Empty-offset1": 8, "Empty-seethrough": 6, "Goalcycle-demo-solo": 8, "Empty5x5x4-crowded": 45}
    for k, sc in enumerate(s for s in val.SCENARIOS if s["name"] in picks):
        sc = dict(sc)
        name = sc.pop("name")
        seed, env_index = 4242 + 5 * k, 77 * k + 1
        rec = RenderRecorder(name, seed=seed, env_index=env_index, rgb=True, **sc)
        rec.reset()
        for t in range(picks[name]):
            act = rng.randint(0, 7, size=len(rec.env.agents))
            act[rng.rand(len(act)) < 0.5] = 2
            if rec.step(act, t) is True:
                rec.reset()
        fname = "render_" + "".join(ch if ch.isalnum() else "_" for ch in name).strip("_") + ".npz"
        np.savez_compressed(os.path.join(OUT, fname), meta=np.frombuffer(json.dumps(cfg_kwargs(rec.cfg, seed, env_index)).encode(), dtype=np.uint8),
                            kind=np.array([e["kind"] for e in rec.ev], np.uint8), actions=np.stack([e["actions"] for e in rec.ev]),
                            img=np.stack([e["img"] for e in rec.ev]))
        print(f"  {fname:48s} {len(rec.ev)} frames {rec.ev[0]['img'].shape}")


if __name__ == "__main__":
    import sys

    os.makedirs(OUT, exist_ok=True)
    if "render" in sys.argv[1:]:
        gen_render()
    elif "hide" in sys.argv[1:]:
        gen_hide_trajectories()
    elif "extra" in sys.argv[1:]:
        gen_extra_trajectories()
        gen_render_extra()
        gen_rich()
    else:
        gen_los()
        gen_atlas()
        gen_trajectories()
        gen_hide_trajectories()
        gen_extra_trajectories()
        gen_render()
        gen_render_extra()
        gen_rich()
