"""world_size-2 gloo test (CPU): the env-index sharding contract of the N>1 path.

Each rank runs ITS shard (oracle port, standing in for the device kernels on a box without GPUs) with
env_offset from marlgrid_b200.sharding; the gathered shards must equal the unsharded batch bit for bit,
and the only collectives used are the off-path statistics / result gathers.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, steps, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from marlgrid_b200.config import make_config
    from marlgrid_b200.sharding import global_stats, shard_range
    from oracle import mg_oracle

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    cfg = make_config(11, 11, ["red", "blue", "purple"], n_clutter=12)
    off, cnt = shard_range(total, rank, world)
    ob = mg_oracle.OracleBatch(cfg, cnt, seed=77, env_offset=off)
    ob.reset()
    rng = np.random.RandomState(123)  # same global action stream on every rank; each takes its slice
    ret = np.zeros(cnt)
    for t in range(steps):
        act = rng.randint(0, 7, size=(total, 3)).astype(np.int32)
        obs, rew, done = ob.step(act[off: off + cnt], autoreset=True, with_obs=True)
        ret += rew.sum(axis=1)
    mean_return = global_stats(ret.sum(), cnt, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, (off, obs, ob.grid.copy(), ob.envrec.copy()))
    dist.barrier()
    if rank == 0:
        out.put((mean_return, gathered))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shards_equal_unsharded_batch(oracle):
    import torch.multiprocessing as mp

    from marlgrid_b200.config import make_config

    total, steps, world = 101, 130, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    mean_return, gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = make_config(11, 11, ["red", "blue", "purple"], n_clutter=12)
    full = oracle.OracleBatch(cfg, total, seed=77)
    full.reset()
    rng = np.random.RandomState(123)
    ret = np.zeros(total)
    for t in range(steps):
        act = rng.randint(0, 7, size=(total, 3)).astype(np.int32)
        obs, rew, done = full.step(act, autoreset=True, with_obs=True)
        ret += rew.sum(axis=1)
    gathered.sort(key=lambda g: g[0])
    assert np.array_equal(np.concatenate([g[1] for g in gathered]), obs)
    assert np.array_equal(np.concatenate([g[2] for g in gathered]), full.grid)
    assert np.array_equal(np.concatenate([g[3] for g in gathered]), full.envrec)
    assert abs(mean_return - ret.mean()) < 1e-12


def _synthetic_transitions(n, seed):
    rng = np.random.RandomState(seed)
    obs = rng.randint(0, 14, size=(n, 7, 7, 3)).astype(np.uint8)
    nxt = rng.randint(0, 14, size=(n, 7, 7, 3)).astype(np.uint8)
    act = rng.randint(0, 7, size=(n,)).astype(np.int32)
    rew = (rng.rand(n) < 0.1).astype(np.float64)
    done = rng.rand(n) < 0.05
    return obs, act, nxt, rew, done


def _learner_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from marlgrid_b200.learners import LinearQLearner

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    learner = LinearQLearner(view_size=7, seed=3)
    o, a, n, r, d = (torch.from_numpy(x) for x in _synthetic_transitions(512, 9))
    half = slice(rank * 256, (rank + 1) * 256)  # each rank learns from its own envs' transitions
    for _ in range(5):
        learner.update(o[half], a[half], n[half], r[half], d[half])
    w = [None] * world
    dist.all_gather_object(w, (learner.W.detach().numpy().copy(), learner.b.detach().numpy().copy()))
    if rank == 0:
        out.put(w)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_learner_gradient_allreduce():
    """LinearQLearner.update under torch.distributed (gloo here, NCCL on the GPU box): both ranks end with the same weights,
    equal to one process learning from the whole batch -- the system's only collective, off the env data path."""
    import torch
    import torch.multiprocessing as mp

    from marlgrid_b200.learners import LinearQLearner

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_learner_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    w = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(w[0][0], w[1][0]) and np.array_equal(w[0][1], w[1][1])
    single = LinearQLearner(view_size=7, seed=3)
    o, a, n, r, d = (torch.from_numpy(x) for x in _synthetic_transitions(512, 9))
    for _ in range(5):
        single.update(o, a, n, r, d)
    assert np.allclose(w[0][0], single.W.detach().numpy(), atol=2e-5) and np.allclose(w[0][1], single.b.detach().numpy(), atol=2e-5)
