"""CPU suite: host-side logic, the C-ABI surface and the failure mode without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol(cuda_lib):
    """Every function declared in include/marlgrid_b200.h is exported by libmarlgrid_b200.so (no compute calls)."""
    hdr = open(os.path.join(ROOT, "include", "marlgrid_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", hdr))
    names -= {"mg_stream_t"}
    assert len(names) >= 20
    raw = ctypes.CDLL(os.path.join(ROOT, "marlgrid_b200", "libmarlgrid_b200.so"))
    for n in sorted(names):
        assert hasattr(raw, n), f"{n} declared in the header but not exported"
    assert cuda_lib.mg_version() == 1
    assert b"sm_100a" in cuda_lib.mg_build_info()


def test_config_struct_matches_c_layout(cuda_lib):
    from marlgrid_b200.config import MgConfig, make_config

    assert cuda_lib.mg_sizeof_config() == ctypes.sizeof(MgConfig)
    cfg = make_config(15, 15, ["red", "blue", "purple"], n_clutter=25)
    assert cuda_lib.mg_config_validate(ctypes.byref(cfg)) == 0
    assert cuda_lib.mg_obs_bytes_per_env(ctypes.byref(cfg), 0) == 3 * 147
    assert cuda_lib.mg_obs_bytes_per_env(ctypes.byref(cfg), 1) == 3 * 56 * 56 * 3
    assert cfg.plane_stride == 240 and cfg.plane_stride % 16 == 0
    bad = make_config(15, 15, ["red"], n_clutter=1)
    bad.plane_stride = 100
    assert cuda_lib.mg_config_validate(ctypes.byref(bad)) == -1
    with pytest.raises(ValueError):
        make_config(2, 9, ["red"])  # Grid needs width, height >= 3 (base.py:98-99)
    with pytest.raises(ValueError):
        make_config(9, 9, ["red"] * 9)


def test_registry_ids_match_reference():
    """Same ids and resolved scenario parameters as marlgrid/envs/__init__.py:70-121 (incl. its quirks)."""
    from marlgrid_b200 import envs

    assert envs.registered_envs == [
        "MarlGrid-1AgentCluttered15x15-v0", "MarlGrid-3AgentCluttered11x11-v0", "MarlGrid-3AgentCluttered15x15-v0",
        "MarlGrid-2AgentEmpty9x9-v0", "MarlGrid-3AgentEmpty9x9-v0", "MarlGrid-4AgentEmpty9x9-v0", "Goalcycle-demo-solo-v0",
    ]


def test_env_construction_fails_loudly_without_gpu():
    """No CPU fallback: on a box without CUDA the product refuses to build an env."""
    import torch

    from marlgrid_b200 import envs

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(Exception) as ei:
        envs.make("MarlGrid-2AgentEmpty9x9-v0")
    assert not isinstance(ei.value, (ImportError, AttributeError)), ei.value
    with pytest.raises(ValueError):
        envs.make("MarlGrid-2AgentEmpty9x9-v0", device="cpu")


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "marlgrid_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libmg_oracle" not in src, f


def test_scenario_kwargs_resolve_like_reference():
    from marlgrid_b200.agents import GridAgentInterface
    from marlgrid_b200.config import GOAL_FIXED, GOAL_NONE, GOAL_RANDOM
    from marlgrid_b200.envs import ClutteredGoalCycleEnv, ClutteredMultiGrid, EmptyMultiGrid

    class Probe(Exception):
        pass

    def cfg_of(cls, **kw):
        # intercept before any device work: BatchedMultiGridEnv.__init__ receives the finished MgConfig
        import marlgrid_b200.envs as E

        orig = E.BatchedMultiGridEnv.__init__

        def grab(self, cfg, **k):
            raise Probe(cfg, k)

        E.BatchedMultiGridEnv.__init__ = grab
        try:
            cls(**kw)
        except Probe as p:
            return p.args
        finally:
            E.BatchedMultiGridEnv.__init__ = orig

    ag = [GridAgentInterface(color=c, view_size=7, view_tile_size=8) for c in ("red", "blue", "purple")]
    cfg, k = cfg_of(ClutteredMultiGrid, agents=ag, grid_size=15, clutter_density=0.15, num_envs=5)
    assert (cfg.width, cfg.height, cfg.n_agents, cfg.n_clutter, cfg.goal_mode) == (15, 15, 3, 25, GOAL_FIXED)  # int(.15*13*13)
    assert k["num_envs"] == 5 and list(cfg.agent_color[:3]) == [0, 3, 5]
    cfg, _ = cfg_of(ClutteredMultiGrid, agents=ag, grid_size=11, n_clutter=7, randomize_goal=True)
    assert (cfg.n_clutter, cfg.goal_mode) == (7, GOAL_RANDOM)
    cfg, _ = cfg_of(EmptyMultiGrid, agents=ag[:2], width=9, height=7, max_steps=50, ghost_mode=False, respawn=True, reward_decay=False)
    assert (cfg.width, cfg.height, cfg.max_steps, cfg.flags & 7, cfg.goal_mode) == (9, 7, 50, 2, GOAL_FIXED)
    cfg, _ = cfg_of(ClutteredGoalCycleEnv, agents=[dict(color="prestige", view_size=7, view_offset=1, view_tile_size=11)], grid_size=13,
                    clutter_density=0.15, n_bonus_tiles=3, penalty=-1.5, max_steps=250, respawn=True, obs_mode="encoded")
    assert (cfg.goal_mode, cfg.n_bonus_tiles, cfg.bonus_penalty, cfg.view_offset, cfg.view_tile_size) == (GOAL_NONE, 3, -1.5, 1, 11)
    assert not (cfg.flags & 4)  # goal-cycle envs default reward_decay=False (goalcycle.py:9,14)
    assert cfg.prestige_mask == 1 and cfg.prestige_beta[0] == 0.95 and cfg.prestige_scale[0] == 2.0
    # agent_spawn_kwargs -> place_obj(top, size, max_tries) (base.py:346,409-412,690-696): the box is clipped like the reference's
    from marlgrid_b200.envs import DoorKeyEnv

    cfg, _ = cfg_of(EmptyMultiGrid, agents=ag, grid_size=9, agent_spawn_kwargs=dict(top=(5, -2), size=(9, 5), max_tries=500))
    assert (tuple(cfg.spawn_top), tuple(cfg.spawn_size), cfg.spawn_max_tries, cfg.scenario) == ((5, 0), (9, 5), 500, 0)
    cfg, _ = cfg_of(EmptyMultiGrid, agents=ag, grid_size=9)
    assert (tuple(cfg.spawn_top), tuple(cfg.spawn_size), cfg.spawn_max_tries) == ((0, 0), (0, 0), 0)  # size None = the whole grid
    with pytest.raises(ValueError, match="empty spawn region"):
        EmptyMultiGrid(agents=ag, grid_size=9, agent_spawn_kwargs=dict(top=(9, 0)))
    with pytest.raises(NotImplementedError, match="reject_fn"):
        EmptyMultiGrid(agents=ag, grid_size=9, agent_spawn_kwargs=dict(reject_fn=lambda pos: False))
    with pytest.raises(TypeError):
        EmptyMultiGrid(agents=ag, grid_size=9, agent_spawn_kwargs=dict(rand_dir=True))
    cfg, k = cfg_of(DoorKeyEnv, agents=ag[:2], grid_size=8)
    assert cfg.scenario == 1 and k["obs_mode"] == "encoded" and tuple(cfg.spawn_size) == (0, 0)
    with pytest.raises(ValueError):
        ClutteredMultiGrid(agents=ag, grid_size=9)  # n_clutter xor clutter_density (cluttered.py:10-11)
    with pytest.raises(ValueError):
        EmptyMultiGrid(agents=[ag[0], GridAgentInterface(view_size=5)], grid_size=9)  # views must be uniform
    with pytest.raises(ValueError):
        EmptyMultiGrid(agents=[object()], grid_size=9)


def test_agent_interface_spaces():
    from marlgrid_b200.agents import GridAgentInterface

    a = GridAgentInterface(view_size=7, view_tile_size=8, color="blue")
    assert a.observation_space.shape == (56, 56, 3) and a.action_space.n == 7
    assert GridAgentInterface(restrict_actions=True).action_space.n == 3
    r = GridAgentInterface(observation_style="rich", observe_position=True, observe_orientation=True)
    assert set(r.observation_space.spaces) == {"pov", "position", "orientation"}
    with pytest.raises(ValueError):
        GridAgentInterface(observation_style="nope")
    c = a.clone()
    assert (c.view_size, c.view_tile_size, c.color) == (7, 8, "blue")


def test_independent_learners_protocol():
    """README.md:21-64 usage: action_step / save_step / episode() with per-agent slices."""
    import torch

    from marlgrid_b200 import IndependentLearners, LearningAgent

    log = []

    class L(LearningAgent):
        def __init__(self, k):
            super().__init__(color=["red", "blue"][k])
            self.k = k

        def action_step(self, obs):
            return torch.full((obs.shape[0],), self.k, dtype=torch.int32) if obs.ndim == 4 else self.k

        def save_step(self, *tr):
            log.append((self.k, tuple(getattr(t, "shape", None) for t in tr[:4])))

        def start_episode(self):
            log.append((self.k, "start"))

        def end_episode(self):
            log.append((self.k, "end"))

    agents = IndependentLearners(L(0), L(1))
    obs = torch.zeros((5, 2, 7, 7, 3), dtype=torch.uint8)
    with agents.episode():
        act = agents.action_step(obs)
        assert tuple(act.shape) == (5, 2) and act[:, 1].eq(1).all()
        agents.save_step(obs, act, obs, torch.zeros(5, 2, dtype=torch.float64), torch.zeros(5, dtype=torch.bool))
    assert log[0] == (0, "start") and log[-1] == (1, "end")
    assert (0, (torch.Size([5, 7, 7, 3]), torch.Size([5]), torch.Size([5, 7, 7, 3]), torch.Size([5]))) in log
    single = [np.zeros((56, 56, 3)), np.zeros((56, 56, 3))]
    assert agents.action_step(single) == [0, 1]


def test_rich_observation_dict_composition():
    """observation_style='rich' (marlgrid/base.py:461-471): checked in the dev container against the unmodified reference
    (reward is always 0 there: base.py:464,519; position = pos / (W, H) float64, (0, 0) when not placed; orientation = dir)."""
    import numpy as np
    import torch

    from marlgrid_b200.env import compose_rich_obs

    agents = torch.zeros((2, 3, 16), dtype=torch.uint8)
    agents[0, 0, :4] = torch.tensor([6, 6, 0, 3], dtype=torch.uint8)   # reference: reset obs of 2AgentEmpty9x9, agent at (6, 6)
    agents[0, 1, :4] = torch.tensor([1, 6, 3, 3], dtype=torch.uint8)
    agents[1, 2, :4] = torch.tensor([5, 2, 1, 0], dtype=torch.uint8)   # not placed: pos is None -> (0, 0)
    pov = torch.zeros((2, 3, 56, 56, 3), dtype=torch.uint8)
    r = compose_rich_obs(pov, agents, 9, 9)
    assert set(r) == {"pov", "reward", "position", "orientation"} and r["pov"] is pov
    assert r["position"].dtype == torch.float64 and tuple(r["position"].shape) == (2, 3, 2)
    assert np.array_equal(r["position"][0, 0].numpy(), np.array([6, 6]) / np.array([9, 9], dtype=float))
    assert np.array_equal(r["position"][0, 1].numpy(), np.array([1, 6]) / np.array([9, 9], dtype=float))
    assert np.array_equal(r["position"][1, 2].numpy(), np.zeros(2))
    assert int(r["orientation"][0, 1]) == 3 and int(r["orientation"][1, 2]) == 1 and int(r["reward"].abs().sum()) == 0
    assert set(compose_rich_obs(pov, agents, 9, 9, observe_rewards=False, observe_orientation=False)) == {"pov", "position"}


def test_marlgrid_import_alias():
    """compat/marlgrid: code written against the reference (`import marlgrid.envs`, README.md:29-36) resolves to this package."""
    import subprocess
    import sys

    code = ("import marlgrid, marlgrid.envs, marlgrid.agents, marlgrid.base, marlgrid.objects;"
            "from marlgrid.agents import GridAgentInterface;"
            "assert marlgrid.envs.registered_envs[1] == 'MarlGrid-3AgentCluttered11x11-v0';"
            "assert marlgrid.IndependentLearners is marlgrid.agents.IndependentLearners;"
            "assert marlgrid.envs.ClutteredMultiGrid.__mro__[1] is marlgrid.base.MultiGridEnv; print('alias ok')")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/")
    assert out.returncode == 0 and "alias ok" in out.stdout, out.stderr


def test_hide_item_types_mask():
    """GridAgentInterface.hide_item_types (agents.py:30): WorldObj.type strings -> MgConfig.hide_types bits; 'Agent' is the
    type string of agents (objects.py:137-138)."""
    import pytest

    from marlgrid_b200.agents import GridAgentInterface
    from marlgrid_b200.config import make_config
    from marlgrid_b200.objects import hide_mask

    assert hide_mask([]) == 0 and hide_mask(["Wall"]) == 1 << 8 and hide_mask(["Goal", "Agent"]) == (1 << 4) | (1 << 13)
    with pytest.raises(ValueError):
        hide_mask(["Unicorn"])
    assert GridAgentInterface(hide_item_types=["Wall"]).clone().hide_item_types == ["Wall"]
    assert make_config(9, 9, ["red"], hide_types=hide_mask(["Door"])).hide_types == 1 << 11


def test_prestige_tile_arithmetic():
    """agents.py:92-119 + blend_tiles base.py:260-273 on the white atlas tile: the host-side composition used by env.render()
    (the device kernels' version is checked against the reference's frames in the GPU suite)."""
    from marlgrid_b200.atlas import build_atlas
    from marlgrid_b200.render import prestige_colour, prestige_tile

    assert list(prestige_colour(0.0, 2.0, False)) == [255, 0, 0] and list(prestige_colour(50.0, 2.0, False)) == [0, 0, 255]
    assert list(prestige_colour(0.0, 2.0, True)) == [127, 0, 127]
    at = build_atlas([12], 8)[:, 0]  # one agent, coloured 'prestige' = white (objects.py:24)
    col = prestige_colour(1.0, 2.0, False)
    t = prestige_tile(at, 0, 1, 5, col)  # on an empty cell, facing right
    alpha = at[1][..., 0].astype(np.int64)
    assert np.array_equal(t[..., 0], (alpha * col[0]) >> 8) and not t[..., 1].any() and np.array_equal(t[..., 2], (alpha * col[2]) >> 8)
    g = prestige_tile(at, 2, 1, 5, col)  # over the Goal: green where the triangle is absent
    assert tuple(g[0, 0]) == (0, 255, 0) and g[..., 1].min() < 255


def test_grid_recorder_on_a_host_env(tmp_path):
    """GridRecorder (utils/video.py:55-154) around any env with reset / step / render: frame bookkeeping and export."""
    import numpy as np

    from marlgrid_b200.utils.video import GridRecorder

    class Env:  # a stand-in with the gym surface the recorder uses
        max_steps = 5

        def __init__(self):
            self.t = 0

        def reset(self):
            self.t = 0
            return "obs"

        def step(self, action):
            self.t += 1
            return "obs", 0.0, self.t >= 5, {}

        def render(self, mode="rgb_array"):
            return np.full((4, 6, 3), self.t, np.uint8)

    rec = GridRecorder(Env(), save_root=str(tmp_path), max_steps=None, auto_save_interval=2)
    assert rec.max_steps == 6 and not rec.recording
    for episode in range(3):
        rec.reset()
        for _ in range(5):
            rec.step(0)
    # auto_save_interval: episodes whose reset count is >= 2 past the last save are filmed and exported at the next reset
    saved = sorted(os.listdir(str(tmp_path)))
    assert any(n.startswith("frames_") for n in saved) and any(n.startswith("video_") for n in saved), saved
    frames_dir = os.path.join(str(tmp_path), [n for n in saved if n.startswith("frames_")][0])
    assert len(os.listdir(frames_dir)) == 6  # five pre-step frames + the final frame appended at reset


def test_shard_ranges_partition_the_batch():
    from marlgrid_b200.sharding import shard_range

    for total, world in ((1048576, 8), (65536, 1), (10, 3), (7, 8)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (o1, c1), (o2, _) in zip(spans, spans[1:]):
            assert o1 + c1 == o2
    assert shard_range(1048576, 3, 8) == (3 * 131072, 131072)


def test_policy_packing_and_cpu_statement():
    """marlgrid_b200.policy packs int8 weights as [A][NW][8][4] for the kernels; the oracle's vectorised Philox equals the
    scalar statement of the RNG contract; the CPU policy statement takes the lowest maximal action and explores as specified."""
    from marlgrid_b200.policy import LinearPolicy, epsilon_to_u32, n_obs_words, pack_bias, pack_weights
    from oracle import philox, policy_oracle

    A, V, K = 3, 7, 5
    n = V * V * 3
    rng = np.random.RandomState(4)
    w = rng.randint(-128, 128, size=(A, K, n)).astype(np.int8)
    pw = pack_weights(w, V)
    assert pw.shape == (A, n_obs_words(V), 8, 4) and pw.dtype == np.int8
    for a, k, i in [(0, 0, 0), (1, 3, 77), (2, 4, n - 1)]:
        assert pw[a, i // 4, k, i % 4] == w[a, k, i]
    assert not pw[:, :, K:, :].any() and pw[:, -1, :, (n % 4):].any() == 0 if n % 4 else True
    assert pack_bias([[1, 2, 3, 4, 5]] * A, A, K).tolist() == [[1, 2, 3, 4, 5, 0, 0, 0]] * A
    assert epsilon_to_u32(0.0) == 0 and epsilon_to_u32(1.0) == 0xFFFFFFFF and epsilon_to_u32(0.5) == 1 << 31
    with pytest.raises(ValueError):
        pack_weights(np.full((A, K, n), 300), V)
    with pytest.raises(ValueError):
        LinearPolicy(np.zeros((A, 8, n), np.int8))
    # vectorised Philox == scalar statement
    c = rng.randint(0, 2**32, size=(4, 50), dtype=np.uint64)
    got = policy_oracle.philox4x32_10_np(c[0], c[1], c[2], c[3], 0x12345678, 0x9ABCDEF0)
    for j in range(50):
        assert tuple(int(x[j]) for x in got) == philox.philox4x32_10(tuple(int(c[q][j]) for q in range(4)), (0x12345678, 0x9ABCDEF0))
    # the policy statement: ties -> lowest k; epsilon = 1 -> always the uniform draw
    obs = rng.randint(0, 12, size=(6, A, V, V, 3)).astype(np.uint8)
    zero = np.zeros((A, K, n), np.int8)
    act = policy_oracle.linear_policy_actions(obs, zero, np.array([[0, 5, 5, 1, 5]] * A, np.int32), K, 0, 0, np.arange(6), np.zeros(6, np.int64))
    assert (act == 1).all()
    g, t = np.arange(6) + 100, np.arange(6) + 9
    act = policy_oracle.linear_policy_actions(obs, zero, np.zeros((A, K), np.int32), K, 0xFFFFFFFF, 99, g, t)
    for b in range(6):
        for a in range(A):
            r = philox.philox4x32_10((int(g[b]), 0, int(t[b]), 0x20000000 | a), (99, 0))
            assert act[b, a] == (philox.mulhi32(r[1], K) if r[0] < 0xFFFFFFFF else 0)
    want = (obs.reshape(6, A, -1).astype(np.int64)[:, :, None, :] * w.astype(np.int64)[None]).sum(-1)
    assert np.array_equal(policy_oracle.linear_policy_actions(obs, w, np.zeros((A, K), np.int32), K, 0, 0, g, t), want.argmax(-1))


def test_linear_q_learner_protocol_and_quantisation():
    """marlgrid_b200.learners: the learner protocol of README.md:21-25 on batched tensors, TD(0) updates that fit a synthetic
    batch, and the int8 hand-off to the rollout kernel (the quantised policy picks the float model's greedy action)."""
    import torch

    from marlgrid_b200 import IndependentLearners
    from marlgrid_b200.learners import LinearQLearner, quantized_policy
    from oracle import policy_oracle

    rng = np.random.RandomState(0)
    learners = IndependentLearners(*[LinearQLearner(view_size=7, epsilon=0.0, seed=k, color=c) for k, c in enumerate(("red", "blue"))])
    obs = torch.from_numpy(rng.randint(0, 14, size=(64, 2, 7, 7, 3)).astype(np.uint8))
    act = learners.action_step(obs)
    assert act.shape == (64, 2) and act.dtype == torch.int32 and int(act.max()) < 7
    nxt = torch.from_numpy(rng.randint(0, 14, size=(64, 2, 7, 7, 3)).astype(np.uint8))
    rew = torch.from_numpy((rng.rand(64, 2) < 0.2).astype(np.float64))
    with learners.episode():
        learners.save_step(obs, act, nxt, rew, torch.zeros(64, dtype=torch.bool))
        assert len(learners[0].buffer) == 1
    assert learners[0].buffer == []  # end_episode consumed it in an update
    l0 = learners[0]
    first = l0.update(obs[:, 0], act[:, 0], nxt[:, 0], rew[:, 0], torch.zeros(64, dtype=torch.bool))
    for _ in range(200):
        last = l0.update(obs[:, 0], act[:, 0], nxt[:, 0], rew[:, 0], torch.ones(64, dtype=torch.bool))  # done: the target is the reward
    assert last < 0.1 * first
    pol = quantized_policy(list(learners), epsilon=0.0)
    assert pol.weights.dtype == np.int8 and np.abs(pol.weights).max(axis=(1, 2)).tolist() == [127, 127]
    greedy = torch.stack([l.q_values(obs[:, k]).argmax(-1) for k, l in enumerate(learners)], dim=1).numpy()
    got = policy_oracle.linear_policy_actions(obs.numpy(), pol.weights, pol.bias, 7, 0, 0, np.arange(64), np.zeros(64, np.int64))
    assert (got == greedy).mean() > 0.9  # (int8 rounding may flip near-ties)


def test_linear_q_trainer_iteration_on_a_stand_in_env():
    """LinearQTrainer.iterate (host logic): quantise -> rollout_policy -> one TD update per agent on the transitions
    (obs[t], act[t+1], rew[t+1], obs[t+1], done[t+1]); checked on a stand-in env that records what it is asked for."""
    import types

    import torch

    from marlgrid_b200.learners import LinearQLearner, LinearQTrainer

    B, A, V = 32, 2, 5
    calls = []

    class StandIn:
        num_envs, num_agents, device = B, A, "cpu"
        cfg = types.SimpleNamespace(view_size=V)

        def policy_act(self, pol):
            calls.append(("act", pol.n_actions, pol.epsilon_u32))
            return torch.zeros((B, A), dtype=torch.int32)

        def rollout_policy(self, pol, first, n_steps, out=None):
            calls.append(("rollout", n_steps, pol.weights.shape, pol.seed))
            g = torch.Generator().manual_seed(len(calls))
            obs, rew, done, act = out
            obs.copy_(torch.randint(0, 14, obs.shape, generator=g, dtype=torch.uint8))
            rew.copy_((torch.rand(rew.shape, generator=g) < 0.1).double())
            done.copy_(torch.rand(done.shape, generator=g) < 0.05)
            act.copy_(torch.randint(0, 7, act.shape, generator=g, dtype=torch.int32))
            act[0] = first
            return out

    learners = [LinearQLearner(view_size=V, seed=k, color=c) for k, c in enumerate(("red", "blue"))]
    tr = LinearQTrainer(StandIn(), learners, horizon=8, epsilon=0.25, seed=100)
    w0 = [l.W.detach().clone() for l in learners]
    s1 = tr.iterate()
    s2 = tr.iterate()
    assert [c[0] for c in calls] == ["act", "rollout", "act", "rollout", "act"]  # the first actions once, then every rollout continues the last
    assert calls[1][1:] == (8, (A, 7, V * V * 3), 100) and calls[3][3] == 101 and calls[0][2] == 1 << 30
    assert all(l.updates == 2 for l in learners) and all(not torch.equal(w, l.W.detach()) for w, l in zip(w0, learners))
    assert set(s1) == {"reward_per_env_step", "loss", "episodes"} and len(s2["loss"]) == A and all(np.isfinite(s2["loss"]))
