"""Times the UNMODIFIED Python reference (kandouss/marlgrid) on the host cores.  BENCH INFRASTRUCTURE ONLY.

Protocol of BASELINE.md section 3 / SURVEY.md 8(d): `multiprocessing.Pool(n_cores)`, one reference env per process
(`env = gym.make(id); env.seed(s); env.reset()`), uniform random actions from `np.random.RandomState(s)`, `reset()` on
`done`, warm-up steps, then timed steps; per-core mean and n-core sum.  Two variants:
  * "rgb"     -- the reference's native `step()` (base.py:501-653 -> gen_agent_obs -> MultiGrid.render, base.py:453-460,301-331)
  * "encoded" -- `env.gen_agent_obs` replaced, on the instance, by `gen_obs_grid` + `MultiGrid.encode`
                 (base.py:418-451,196-214): the observation BASELINE.json's encoded configs name (the reference never calls
                 `encode` itself)
The reference source is imported from oracle/_ref (staged, sha256-verified copy; see oracle/stage_reference.py) or, in the
dev container, straight from /root/reference -- unmodified either way; only the import shims of oracle/shims (gym,
gym_minigrid.rendering, pyglet stand-ins) and the numpy 1.x aliases are added.  Timing uses the RAW reference: the
line-of-sight zero-padding patch of the parity runs is NOT applied here.

    python -m oracle.reference_bench --env-id MarlGrid-3AgentCluttered15x15-v0 --variant encoded --procs 16 --seconds 10
prints one JSON object.
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    from oracle import stage_reference

    staged = os.path.join(HERE, "_ref")
    if os.path.isdir(os.path.join(staged, "marlgrid")):
        if not stage_reference.verify():
            raise RuntimeError("oracle/_ref does not match its manifest: re-run `python -m oracle.stage_reference`")
        return staged, "oracle/_ref (staged unmodified copy, sha256-verified)"
    if os.path.isdir("/root/reference/marlgrid"):
        return "/root/reference", "/root/reference"
    return None, "unavailable"


def _import_reference(root):
    import numpy as np

    for alias, typ in (("bool", bool), ("float", float), ("int", int)):  # numpy >= 1.24 dropped them (base.py:424,467,510,741)
        if not hasattr(np, alias):
            setattr(np, alias, typ)
    for p in (root, os.path.join(HERE, "shims")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import gym
    import marlgrid.envs  # noqa: F401  (registers the env ids)
    import marlgrid.base as base

    assert os.path.abspath(base.__file__).startswith(os.path.abspath(root)), base.__file__
    return gym


def _worker(args):
    root, env_id, variant, seed, warmup, seconds = args
    import numpy as np

    gym = _import_reference(root)
    env = gym.make(env_id)
    env.seed(seed)
    env.reset()
    if variant == "encoded":
        def encoded(agent, _env=env):
            grid, vis = _env.gen_obs_grid(agent)
            return grid.encode(vis)

        env.gen_agent_obs = encoded  # instance attribute: the class and the source stay untouched
    rng = np.random.RandomState(seed)
    A = len(env.agents)

    def run(n_steps=None, budget=None):
        t0 = time.perf_counter()
        n = 0
        while True:
            _obs, _rew, done, _ = env.step(rng.randint(0, 7, size=A))
            if done:
                env.reset()
            n += 1
            if n_steps is not None and n >= n_steps:
                break
            if budget is not None and (n & 15) == 0 and time.perf_counter() - t0 >= budget:
                break
        return n, time.perf_counter() - t0

    run(n_steps=warmup)  # includes numba's JIT of occlude_mask and the tile cache fill
    n, dt = run(budget=seconds)
    return n, dt


def measure(env_id, variant, procs, seconds, warmup=200):
    import multiprocessing as mp

    root, where = reference_root()
    if root is None:
        return {"unavailable": "reference source not staged (oracle/_ref) and /root/reference absent"}
    ctx = mp.get_context("spawn")  # the parent may hold a CUDA context
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_worker, [(root, env_id, variant, 1000 + i, warmup, seconds) for i in range(procs)])
    rates = [n / dt for n, dt in res]
    return {
        "value": sum(rates), "unit": "env-steps/s", "cores": procs, "kind": "reference", "per_core_mean": sum(rates) / len(rates),
        "env_id": env_id, "variant": variant, "steps_timed": int(sum(n for n, _ in res)), "seconds_per_proc": seconds, "warmup_steps": warmup,
        "wall_s": time.perf_counter() - t0, "source": where,
        "model": "multiprocessing.Pool(%d), one env per process, uniform random actions, reset() on done (BASELINE.md section 3)" % procs,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env-id", default="MarlGrid-3AgentCluttered15x15-v0")
    ap.add_argument("--variant", default="encoded", choices=["encoded", "rgb"])
    ap.add_argument("--procs", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--warmup", type=int, default=200)
    a = ap.parse_args()
    print(json.dumps(measure(a.env_id, a.variant, a.procs, a.seconds, a.warmup)), flush=True)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(HERE))
    main()
