#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched MarlGrid hot path (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
  torchrun --nproc-per-node N ... bench.py --gpus N ...         # N ranks, one per GPU, env-index sharded
  python bench.py --impl reference ...                          # CPU arm: the unmodified Python reference on the host cores

A "step" is one env.step() of the whole batch: MarlGrid-3AgentCluttered15x15-v0, 65 536 envs per GPU,
encoded observations, uniform random actions, auto-reset (BASELINE.json configs[2]).  One JSON line is
printed by rank 0.

Timing (`value`): W warm-up steps (+ an untimed clock ramp), then REPEATS x exactly K steps back to back, every
repeat between its own pair of CUDA events on the launching stream, the whole series between barrier +
synchronize; `value` = B*K / median repeat (max over ranks per repeat), REPEATS chosen so that the timed region
lasts >= 60 ms whatever K is.  The steps visit `--replicas` (6) independent env families of 65 536 envs round
robin, so a family's state has been evicted from the 126 MB L2 by the time it is stepped again ("inputs larger
than L2"; no flush kernel between launches); `sustained` is the mean over all repeats (all-reset steps included).
`e2e` drives the C-ABI host-buffer engine (mg_engine_step): pinned host actions in, obs/rewards/done out, copies
inside the timer; its last step is checked against the oracle outside the timer.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENV_ID = "MarlGrid-3AgentCluttered15x15-v0"
# BASELINE.json configs: name -> (env id, batch, obs mode, agents, algorithmic bytes per env-step [SURVEY.md 8(d)])
WORKLOADS = {
    "cfg3": ("MarlGrid-3AgentCluttered15x15-v0", 65536, "encoded", 3, 1272),
    "cfg2": ("MarlGrid-3AgentCluttered11x11-v0", 4096, "encoded", 3, 957),
    "cfg4": ("MarlGrid-4AgentEmpty9x9-v0", 262144, "rgb", 4, 38070),
    "cfg5_share": ("MarlGrid-3AgentCluttered15x15-v0", 131072, "encoded", 3, 1272),
}
ALGO_BYTES_PER_ENV_STEP = 1272  # 743 read + 522 write + 7 amortised reset (whole env.step)
FALLBACK_HBM_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md fallback
MIN_TIMED_MS = 60.0             # total timed region of the headline figure, whatever --steps is


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: an NVML polling thread (5 ms period; the
    timed region of this bench lasts ~0.1 s, too short for `nvidia-smi -lms` to even start), nvidia-smi as fallback."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index
        self.sm, self.power, self.reason_bits = [], [], 0
        self.sm_max = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.source = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.source = None

    def _poll_nvml(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def _one_shot_smi(self):
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self._physical_index()}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=20).stdout.strip().splitlines()
            parts = [p.strip() for p in out[0].split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(parts[0]), "sm_max_mhz": float(parts[1]), "samples": 1, "source": "nvidia-smi (single sample after the timed region)",
                    "reasons": sorted(n for n, v in zip(names, parts[3:7]) if v.lower().startswith("active"))}
        except Exception:  # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml and nvidia-smi unavailable"]}

    def stop(self):
        if self.source != "nvml":
            return self._one_shot_smi()
        self.stop_flag.set()
        self.thread.join(timeout=2)
        nv = self.nv
        names = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                 ("hw_power_brake", nv.nvmlClocksEventReasonHwPowerBrakeSlowdown)]
        reasons = sorted(n for n, b in names if self.reason_bits & b)
        sm = sorted(self.sm)
        if not sm:
            return self._one_shot_smi()
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max, "samples": len(sm), "source": "nvml, 5 ms period",
                "power_w_max": max(self.power) if self.power else None, "reasons": reasons}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:  # noqa: BLE001
        pass
    return "unknown"


def cpu_port_throughput(n_envs, threads, target_s=10.0, seed=1337):
    """The oracle (C port of the reference algorithm, oracle/mg_oracle.c) on a bounded sample of the same
    workload: n_envs envs, as many steps as fit ~target_s seconds (probed first), all `threads` host threads."""
    import numpy as np

    from marlgrid_b200.config import GOAL_FIXED, make_config
    from oracle import mg_oracle

    cfg = make_config(15, 15, ["red", "blue", "purple"], view_size=7, view_tile_size=8, n_clutter=int(0.15 * 13 * 13), goal_mode=GOAL_FIXED)
    ob = mg_oracle.OracleBatch(cfg, n_envs, seed=seed, threads=threads)
    ob.reset()
    rng = np.random.RandomState(0)
    act = rng.randint(0, 7, size=(16, n_envs, 3)).astype(np.int32)
    t0 = time.perf_counter()
    ob.rollout(act, n_steps=20)  # warm-up + probe
    probe = (time.perf_counter() - t0) / 20
    n_steps = int(min(50000, max(100, target_s / max(probe, 1e-6))))
    t0 = time.perf_counter()
    ob.rollout(act, n_steps=n_steps)
    dt = time.perf_counter() - t0
    return n_envs * n_steps / dt, dt, n_steps


def cpu_port_baseline(target_s):
    thr = host_threads()
    n_envs = 65536 if thr >= 32 else 8192
    v, dt, n_steps = cpu_port_throughput(n_envs, thr, target_s=target_s)
    return {"value": v, "unit": "env-steps/s", "cores": thr, "kind": "port", "cpu": cpu_model(),
            "sample": f"{n_envs} envs x {n_steps} steps of the same workload ({dt:.1f} s), multithreaded C port of the reference step+reset+encode (oracle/mg_oracle.c)"}


def python_reference_baseline(env_id, variant, seconds):
    """The UNMODIFIED Python reference (oracle/_ref, staged by oracle/stage_reference.py) on all host cores, in a child
    process (the parent may hold a CUDA context): oracle/reference_bench.py, protocol of BASELINE.md section 3."""
    procs = host_threads()
    try:
        out = subprocess.run([sys.executable, "-m", "oracle.reference_bench", "--env-id", env_id, "--variant", variant, "--procs", str(procs),
                              "--seconds", str(seconds)], cwd=ROOT, capture_output=True, text=True, timeout=60 + 6 * seconds)
        for line in reversed(out.stdout.strip().splitlines()):
            if line.startswith("{"):
                d = json.loads(line)
                if "value" in d:
                    d["cpu"] = cpu_model()
                    d["sample"] = (f"{procs} processes x {seconds:.0f} s of {env_id} ({variant} observations; {d['steps_timed']} env.step calls), "
                                   f"unmodified reference from {d['source']}")
                return d
        return {"unavailable": "oracle.reference_bench printed no result: " + (out.stderr.strip().splitlines() or ["?"])[-1][:200]}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}


def workload_config(args, world):
    return {
        "workload": f"{ENV_ID} batch={args.batch_per_gpu}/GPU x {world} GPU(s), encoded obs [B,3,7,7,3] u8, uniform random actions, auto-reset "
                    f"(BASELINE.json configs[2]{'; the sharded family of configs[4], whose own 131072/GPU share is other_configs.cfg5_share' if world > 1 else ''})",
        "env_id": ENV_ID, "batch_per_gpu": args.batch_per_gpu, "global_batch": args.batch_per_gpu * world, "n_agents": 3,
        "parallelism": f"env-index sharding x{world}, no collective on the data path",
        "l2": f"inputs larger than L2: {getattr(args, 'replicas', 6)} independent env families of batch_per_gpu envs stepped round robin (each step touches ~51 MB, "
              "a family is revisited after > 126 MB of other traffic); repeats of K steps back to back, one CUDA event pair per repeat, median repeat",
    }


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path -- the unmodified Python package staged in
    oracle/_ref -- on all host cores (encoded-observation variant of BASELINE.md section 3 for this workload); the
    multithreaded C port is timed next to it (`cpu_port`, the harder baseline).  A bench "step" of this arm = one env.step
    on each of the `cores` envs (one per process)."""
    if rank != 0:
        return
    py = python_reference_baseline(ENV_ID, "encoded", seconds=20.0)
    port = cpu_port_baseline(target_s=12.0)
    if "value" in py:
        v, cores, kind, sample = py["value"], py["cores"], "reference", py["sample"]
    else:  # oracle/_ref was not staged on this box: fall back to the port, and say so
        v, cores, kind, sample = port["value"], port["cores"], "port", port["sample"] + f" [python reference unavailable: {py.get('unavailable')}]"
    line = {
        "impl": "reference", "metric": "env-steps/s", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * cores / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample, "cpu": cpu_model(),
                         "per_core_mean": py.get("per_core_mean"), "model": py.get("model")},
        "cpu_port": port,
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "agent_steps_per_s": 3 * v,
    }
    print(json.dumps(line), flush=True)


def pin_to_gpu_numa_node(local_rank):
    """Run this rank's host thread -- and therefore first-touch its pinned buffers -- on the CPUs NVML reports as local to its
    GPU (e2e at N > 1: all ranks copying through one NUMA node was a suspect for the flat 1 -> 8 GPU curve of round 1)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local_rank
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                idx = int(ids[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            before = len(os.sched_getaffinity(0))
            os.sched_setaffinity(0, allowed)
            return {"cpus_local_to_gpu": len(allowed), "cpus_before": before}
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}
    return {"cpus_local_to_gpu": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=65536)
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--replicas", type=int, default=6, help="independent env families visited round robin in the timed loop (working set > L2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline figure + e2e only: no other_configs / desync / CPU arms")
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS), help="another BASELINE config alone (informational line; the driver's line is cfg3)")
    ap.add_argument("--profile", action="store_true", help="for runs under ncu: --quick with 2 repeats and no clock ramp")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.steps = max(args.steps, 1)
    args.quick = args.quick or args.profile

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if world > 1:
        # communicator lines (rank / nranks) stay visible for the driver, but on stderr: stdout carries the JSON line.  Set before
        # torch is imported: NCCL reads its debug settings once.
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):  # (this image presets NCCL_DEBUG=VERSION)
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

    import ctypes

    import numpy as np
    import torch

    from marlgrid_b200 import _lib, envs
    from marlgrid_b200.config import MgState

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    B, K, W = args.batch_per_gpu, args.steps, args.warmup
    L = _lib.load()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(values):
        t = torch.tensor(values, dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    # ---- another BASELINE config: `steps` env.step() calls back to back over n_fams families, one CUDA event pair -------
    peak_gbs, _ = measured_peak()

    def timed_rollout(make_env, n_fams, steps, name):
        env_id, Bo, mode, Ao, algo = WORKLOADS[name]
        fs = [make_env(r) for r in range(n_fams)]
        for f in fs:
            f.reset()
        acts = [fs[0].random_actions(t, seed=rank) for t in range(16)]
        for t in range(max(3, n_fams)):
            fs[t % n_fams].step(acts[t % 16])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for t in range(steps):
            fs[t % n_fams].step(acts[t % 16])
        e1.record()
        host_s = time.perf_counter() - t0
        barrier()
        ms = max_over_ranks([e0.elapsed_time(e1) / steps])[0]
        py = None
        if n_fams == 1 and mode == "encoded":  # launch-bound sizes: the same steps enqueued from C (mg_rollout_fused), without the Python call per step
            py = {"ms_per_step": ms, "host_issue_ms_per_step": 1e3 * host_s / steps, "note": "env.step() called in a Python loop"}
            tape = torch.stack(acts)
            fs[0].rollout(tape)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps // 16):
                fs[0].rollout(tape)
            e1.record()
            barrier()
            ms = max_over_ranks([e0.elapsed_time(e1) / (steps // 16 * 16)])[0]
        ach = algo * Bo / (ms * 1e-3) / 1e9
        del fs
        torch.cuda.empty_cache()
        return {"env_id": env_id, "batch_per_gpu": Bo, "obs": mode, "value": world * Bo / (ms * 1e-3), "agent_steps_per_s": world * Ao * Bo / (ms * 1e-3),
                "ms_per_step": ms, "steps": steps, "host_issue_ms_per_step": 1e3 * host_s / steps, "python_loop": py,
                "roofline": {"achieved": ach, "peak": peak_gbs, "frac": ach / peak_gbs, "algorithmic_bytes_per_launch": algo * Bo}}

    def mk(name, **kw):
        env_id, Bo, mode, _, _ = WORKLOADS[name]
        return lambda r: envs.make(env_id, num_envs=Bo, obs_mode=mode, seed=1337, env_offset=(rank * 8 + r) * Bo, device=dev, **kw)


    if args.workload != "cfg3":
        nf, st = {"cfg2": (1, 400), "cfg5_share": (3, 120), "cfg4": (1, 12)}[args.workload]
        res = timed_rollout(mk(args.workload, **({"obs_buffers": 1} if args.workload == "cfg4" else {})), nf, max(args.steps, 3) if args.profile else st, args.workload)
        if rank == 0:
            print(json.dumps(dict(res, metric="env-steps/s", workload=args.workload, n_gpus=world)), flush=True)
        if dist is not None:
            dist.destroy_process_group()
        return

    # R independent env families of B envs each, stepped round robin: between two visits of a family the other R-1 steps
    # touch (R-1) x ~51 MB (bit-plane lines, records, actions in; observations, records, rewards out), more than the 126 MB
    # L2 holds, so every timed step finds its inputs in HBM -- "inputs larger than L2", no flush kernel between launches.
    R = args.replicas
    fams = [envs.make(ENV_ID, num_envs=B, obs_mode="encoded", seed=1337, env_offset=(rank * R + r) * B, device=dev) for r in range(R)]
    env = fams[0]
    A = env.num_agents
    for f in fams:
        f.reset()
    POOL = 128
    actions = torch.empty((POOL, B, A), dtype=torch.int32, device=dev)
    for t in range(POOL):
        env.random_actions(t, seed=rank, out=actions[t])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    PP = ctypes.c_void_p * R
    rotations = []  # the round robin continues across calls: rotation r starts at family r
    for r0 in range(R):
        order = [fams[(r0 + i) % R] for i in range(R)]
        rotations.append(((MgState * R)(*[f._state for f in order]), PP(*[f.rewards.data_ptr() for f in order]),
                          PP(*[f.done.data_ptr() for f in order]), PP(*[f.obs.data_ptr() for f in order])))
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    cfg_ref = ctypes.byref(env.cfg)
    step_no = [0]  # global step counter of the round robin

    def run_steps(n):
        while n > 0:
            m = min(n, POOL)
            st, rp, dp, op = rotations[step_no[0] % R]
            _lib.check(L.mg_rollout_fused_rr(cfg_ref, st, R, actions.data_ptr(), m, rp, dp, op, 1, stream), "mg_rollout_fused_rr")
            step_no[0] += m
            n -= m

    # ---- warm-up: W steps, then an untimed clock ramp (a fresh box idles at ~800 MHz) -----------------------------
    run_steps(W)
    ramp_steps = 0
    if not args.profile:
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 0.25:
            run_steps(120)
            ramp_steps += 120
            torch.cuda.synchronize()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed: REPEATS x exactly K cold steps, back to back, one event pair per repeat ------------------------------
    est_us = 16.0 * max(1.0, B / 65536.0)
    REPEATS = int(min(3000, max(5, -(-MIN_TIMED_MS * 1e3 // (K * est_us)))))
    if args.profile:
        REPEATS = 2
    launches0 = L.mg_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(REPEATS + 1)]
    barrier()
    wall0 = time.perf_counter()
    evs[0].record()
    for i in range(REPEATS):
        run_steps(K)  # ONE launch per step: fused step + auto-reset + observe kernel
        evs[i + 1].record()
    barrier()
    wall_cold = time.perf_counter() - wall0
    launches = L.mg_launch_count() - launches0
    rep_ms = max_over_ranks([evs[i].elapsed_time(evs[i + 1]) for i in range(REPEATS)])
    clocks = sampler.stop() if rank == 0 else None
    srt_rep = sorted(rep_ms)
    median_ms, total_ms = srt_rep[len(srt_rep) // 2], sum(rep_ms)

    extras = {}
    if not args.quick:
        # ---- secondary: per-step events with a 256 MiB L2 flush before each step (distribution of single cold steps) ----
        KF = 100
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(KF)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(KF)]
        for t in range(KF):
            flush.fill_(t & 0xFF)
            starts[t].record()
            env.step(actions[t % POOL])
            stops[t].record()
        barrier()
        srt = sorted(s.elapsed_time(e) for s, e in zip(starts, stops))
        extras["step_ms_flushed"] = {"min": srt[0], "median": srt[len(srt) // 2], "p99": srt[min(len(srt) - 1, int(0.99 * len(srt)))], "max": srt[-1], "steps": len(srt),
                                     "note": "secondary: single steps, 256 MiB L2 flush before each, per-step CUDA events (~2 us resolution)"}

        # ---- the all-reset step: every episode of the family ends on the same step (uniform random actions keep them in lock step) ----
        reset_us = []
        for _ in range(3):
            env.reset()
            env.rollout(actions[:99])
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            env.step(actions[99])
            e1.record()
            torch.cuda.synchronize()
            assert float(env.done.float().mean().item()) > 0.99  # (an env whose agents all reached the goal earlier is on another schedule)
            reset_us.append(1e3 * e0.elapsed_time(e1))
        extras["all_reset_us"] = sorted(max_over_ranks(reset_us))[1]

        # ---- K warm steps back to back on ONE family (state L2-resident: what a rollout loop at this batch sees) ----
        KW = max(K, 500)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        done_steps = 0
        while done_steps < KW:
            n = min(POOL, KW - done_steps)
            env.rollout(actions[:n])
            done_steps += n
        e1.record()
        barrier()
        warm_ms = max_over_ranks([e0.elapsed_time(e1) / KW])[0]
        extras["warm"] = {"value": world * B / (warm_ms * 1e-3), "ms_per_step": warm_ms, "steps": KW,
                          "note": "steps back to back on one family, state L2-resident, launched from mg_rollout_fused"}

        # ---- desynchronised episodes: the steady state of a long-running batch.  The step counters of a fresh family are
        # spread uniformly over an episode, so every step ~1 % of the envs (in ~27 % of the tiles) time out and are regenerated
        # inside the kernel, instead of all of them on every 100th step.
        env.reset()
        env.envrec[:, 0] = torch.randint(0, 100, (B,), device=dev, dtype=torch.int32)
        env.rollout(actions[:100])  # warm-up; also spreads the episode numbers
        KD = 400
        pg_stats = (ctypes.c_uint64 * 2)()
        L.mg_pregen_stats(pg_stats, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        done_steps = 0
        while done_steps < KD:
            n = min(POOL, KD - done_steps)
            env.rollout(actions[:n])
            done_steps += n
        e1.record()
        torch.cuda.synchronize()
        d_us = max_over_ranks([1e3 * e0.elapsed_time(e1) / KD])[0]
        sc = env.step_count
        L.mg_pregen_stats(pg_stats, 0)
        extras["desync"] = {"us_per_step": d_us, "vs_lockstep_warm": d_us / (1e3 * warm_ms), "steps": KD,
                            "pregen_hit_rate": float(pg_stats[0]) / max(1.0, float(pg_stats[0] + pg_stats[1])),
                            "envs_finishing_per_step": float((sc == 0).sum().item()),
                            "note": "one family (L2-resident) whose step counters are spread uniformly over the episode length: ~655 envs finish and are "
                                    "regenerated in every step; ratio against `warm` (same regime, all episodes in lock step, no reset in the window average)"}

        # ---- on-device rollout loop: 100 steps per launch (mg_rollout_persistent), every step's outputs written to its own slice ----
        TP = 100
        env.reset()  # episodes back in lock step (the desynchronised family above is a different regime)
        pout = (torch.empty((TP, B, A, 7, 7, 3), dtype=torch.uint8, device=dev), torch.empty((TP, B, A), dtype=torch.float64, device=dev),
                torch.empty((TP, B), dtype=torch.bool, device=dev))
        env.rollout_all(actions[:TP], out=pout)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            env.rollout_all(actions[:TP], out=pout)
        e1.record()
        barrier()
        persistent_ms = max_over_ranks([e0.elapsed_time(e1) / (5 * TP)])[0]
        del pout
        extras["rollout_persistent"] = {"value": world * B / (persistent_ms * 1e-3), "ms_per_step": persistent_ms, "steps_per_launch": TP,
                                        "note": "on-device rollout loop (mg_rollout_persistent): 100 steps per launch on a fixed action tape, tile state resident in "
                                                "shared memory between steps, every step's obs / rewards / done written to HBM (informational: open-loop)"}

        # ---- closed loop on the device: the policy (int8 linear layer + epsilon-greedy) evaluated inside the rollout kernel ----
        from marlgrid_b200.policy import LinearPolicy

        pol = LinearPolicy.random(A, 7, n_actions=7, epsilon=0.1, seed=5)
        qout = (torch.empty((TP, B, A, 7, 7, 3), dtype=torch.uint8, device=dev), torch.empty((TP, B, A), dtype=torch.float64, device=dev),
                torch.empty((TP, B), dtype=torch.bool, device=dev), torch.empty((TP, B, A), dtype=torch.int32, device=dev))
        env.rollout_policy(pol, actions[0], TP, out=qout)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            env.rollout_policy(pol, actions[0], TP, out=qout)
        e1.record()
        barrier()
        policy_ms = max_over_ranks([e0.elapsed_time(e1) / (5 * TP)])[0]
        del qout
        extras["rollout_policy"] = {"value": world * B / (policy_ms * 1e-3), "ms_per_step": policy_ms, "steps_per_launch": TP,
                                    "note": "closed-loop rollout (mg_rollout_policy): step t+1 plays the actions an int8 linear policy (7 actions, epsilon 0.1) chose from "
                                            "step t's observations, evaluated on the observation tile inside the rollout kernel; every step's obs / rewards / done / actions written"}

        from marlgrid_b200.agents import GridAgentInterface

        # ---- env kwargs off the specialised shape (general fused kernel / step kernel + observe kernel): parity-tested, here timed ----
        feats, local_ms, local_launches = {}, [], []
        feat_cases = (("hide_item_types=['Goal']", {"hide_item_types": ["Goal"]}, {}), ("ghost_mode=False", {}, {"ghost_mode": False}),
                      ("respawn=True", {}, {"respawn": True}), ("see_through_walls=True", {"see_through_walls": True}, {}))
        for name, akw, ekw in feat_cases:
            try:  # (no collective inside the try: a rank that failed alone must not leave the others waiting)
                fe = envs.ClutteredMultiGrid(agents=[GridAgentInterface(color=c, view_size=7, view_tile_size=8, **akw) for c in ("red", "blue", "purple")],
                                             grid_size=15, clutter_density=0.15, num_envs=B, obs_mode="encoded", seed=1337, env_offset=rank * B, device=dev, **ekw)
                fe.reset()
                fe.rollout(actions[:20])
                torch.cuda.synchronize()
                l0 = L.mg_launch_count()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fe.rollout(actions[:100])
                fe.rollout(actions[:100])
                e1.record()
                torch.cuda.synchronize()
                local_ms.append(e0.elapsed_time(e1) / 200)
                local_launches.append((L.mg_launch_count() - l0) / 200.0)
                del fe
            except Exception as ex:  # noqa: BLE001 -- informational section: report, do not fail the bench line
                local_ms.append(float("nan"))
                local_launches.append(repr(ex))
        barrier()
        for (name, _, _), f_ms, nl in zip(feat_cases, max_over_ranks(local_ms), local_launches):
            feats[name] = {"value": world * B / (f_ms * 1e-3), "ms_per_step": f_ms, "launches_per_step": nl} if f_ms == f_ms else {"error": str(nl)}
        extras["feature_paths"] = dict(feats, note="cfg3 with one env kwarg changed, 200 warm steps enqueued from C on one family (compare `warm`)")

        # ---- the same warm steps through the Python surface, env.step(actions) called in a Python loop -------------------
        KP_ = max(K, 500)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for t in range(KP_):
            env.step(actions[t % POOL])
        e1.record()
        host_issue_s = time.perf_counter() - t0  # time the host needed to enqueue the steps
        barrier()
        py_ms = e0.elapsed_time(e1) / KP_
        extras["python_api"] = {"value": world * B / (py_ms * 1e-3), "ms_per_step": py_ms, "host_issue_ms_per_step": 1e3 * host_issue_s / KP_,
                                "note": "env.step(actions) in a Python loop on one family (rank 0's figures; host-bound when host_issue_ms_per_step ~ ms_per_step)"}
    del fams[1:]
    del rotations
    torch.cuda.empty_cache()

    # ---- other BASELINE configs, measured in the same run (device-timed, max over ranks) --------------------------------
    other = {}
    if not args.quick:
        # cfg2: 4 096 envs -- 128 tiles, launch / latency bound; the state of one family is L2-resident whatever one does
        other["cfg2"] = timed_rollout(mk("cfg2"), 1, 400, "cfg2")
        other["cfg2"]["note"] = "BASELINE configs[1]: 128 tiles on 148 SMs -- one launch per step enqueued from C (mg_rollout_fused), launch / latency bound; state L2-resident (1 MB)"
        # cfg5's per-GPU share: 131 072 envs; 3 families round robin (a step touches ~100 MB)
        other["cfg5_share"] = timed_rollout(mk("cfg5_share"), 3, 120, "cfg5_share")
        other["cfg5_share"]["note"] = "BASELINE configs[4] per-GPU share (131 072 envs/GPU; x N GPUs = the sharded batch), 3 families round robin"
        # cfg4: RGB, 262 144 envs, 9.9 GB of observations per step (>> L2)
        other["cfg4"] = timed_rollout(mk("cfg4", obs_buffers=1), 1, 12, "cfg4")
        other["cfg4"]["note"] = "BASELINE configs[3]; RGB observations [B,4,56,56,3] u8, every step writes 9.9 GB"

    # ---- e2e: host buffers through the C ABI engine ----------------------------------------------
    numa = pin_to_gpu_numa_node(local_rank) if world > 1 else None
    h = ctypes.c_void_p()
    _lib.check(L.mg_engine_create(ctypes.byref(h), ctypes.byref(env.cfg), B, rank * B, 1337, local_rank, 0, None, 0), "mg_engine_create")
    obs_bytes, rew_bytes, act_bytes = B * A * 147, B * A * 8, B * A * 4
    p_obs, p_rew, p_done = L.mg_host_alloc(obs_bytes), L.mg_host_alloc(rew_bytes), L.mg_host_alloc(B)
    # 8 different action batches, each in its own pinned buffer (the policy's output as it would sit in host memory):
    # every step's H2D copy reads a different one
    host_actions = np.random.RandomState(rank).randint(0, 7, size=(8, B * A)).astype(np.int32)
    p_acts = []
    for i in range(8):
        pa = L.mg_host_alloc(act_bytes)
        np.ctypeslib.as_array(ctypes.cast(pa, ctypes.POINTER(ctypes.c_int32)), shape=(B * A,))[:] = host_actions[i]
        p_acts.append(pa)
    _lib.check(L.mg_engine_reset(h, p_obs), "mg_engine_reset")
    KE = args.e2e_steps
    WE = 5
    for t in range(WE):
        _lib.check(L.mg_engine_step(h, p_acts[t % 8], p_obs, p_rew, p_done, 1), "mg_engine_step")
    barrier()
    t0 = time.perf_counter()
    for t in range(KE):
        _lib.check(L.mg_engine_step(h, p_acts[(WE + t) % 8], p_obs, p_rew, p_done, 1), "mg_engine_step")
    barrier()
    e2e_s = time.perf_counter() - t0
    # the copy ceiling of this box: the same transfers (same buffers, same slicing, same streams) without the kernel
    barrier()
    t0 = time.perf_counter()
    for t in range(KE):
        _lib.check(L.mg_engine_copy_only(h, p_acts[t % 8], p_obs, p_rew, p_done), "mg_engine_copy_only")
    barrier()
    copy_s = time.perf_counter() - t0
    e2e_ms, copy_ms = max_over_ranks([e2e_s * 1e3, copy_s * 1e3])
    # outside the timer: the last step's results against the oracle replaying the same actions (rank 0)
    e2e_parity = None
    if rank == 0 and not args.no_cpu_baseline and not args.quick:
        from oracle import mg_oracle

        # the copy-only loop overwrote the host buffers with the device's (unchanged) last results: still the last step's outputs
        obs_last = np.ctypeslib.as_array(ctypes.cast(p_obs, ctypes.POINTER(ctypes.c_uint8)), shape=(B, A, 7, 7, 3))
        rew_last = np.ctypeslib.as_array(ctypes.cast(p_rew, ctypes.POINTER(ctypes.c_double)), shape=(B, A))
        done_last = np.ctypeslib.as_array(ctypes.cast(p_done, ctypes.POINTER(ctypes.c_uint8)), shape=(B,))
        ob = mg_oracle.OracleBatch(env.cfg, B, seed=1337, env_offset=0, threads=host_threads())
        ob.reset()
        o2 = r2 = d2 = None
        for t in range(WE + KE):
            last = t == WE + KE - 1
            res = ob.step(host_actions[t % 8].reshape(B, A), autoreset=True, with_obs=last)
            if last:
                o2, r2, d2 = res
        ok = bool(np.array_equal(obs_last, o2) and np.array_equal(rew_last.view(np.uint64), r2.view(np.uint64)) and np.array_equal(done_last, d2))
        e2e_parity = {"ok": ok, "checked": f"obs / reward bits / done of step {WE + KE} of all {B} envs == oracle replay of the same host actions",
                      "obs_checksum": int(obs_last.reshape(-1)[::4099].astype(np.int64).sum())}
    L.mg_engine_destroy(h)

    if rank == 0:
        peak, peak_src = measured_peak()
        value = world * B * K / (median_ms * 1e-3)
        sustained = world * B * K * REPEATS / (total_ms * 1e-3)
        e2e_value = world * B * KE / (e2e_ms * 1e-3)
        d2h = obs_bytes + rew_bytes + B
        avg_step_s = (median_ms / K) * 1e-3  # one launch per step, back to back: the kernel's average launch duration (launch gaps included)
        achieved = ALGO_BYTES_PER_ENV_STEP * B / avg_step_s / 1e9
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
                traffic, traffic_src = tj.get("fused_kernel_dram_bytes_per_launch"), tj.get("source")
        except Exception:  # noqa: BLE001
            pass
        cpu = cpu_py = None
        if not args.no_cpu_baseline and not args.quick:
            cpu_py = python_reference_baseline(ENV_ID, "encoded", seconds=8.0)
            cpu_port = cpu_port_baseline(target_s=8.0)
            if "value" in cpu_py:
                cpu = {"value": cpu_py["value"], "unit": "env-steps/s", "cores": cpu_py["cores"], "kind": "reference", "sample": cpu_py["sample"],
                       "cpu": cpu_py["cpu"], "per_core_mean": cpu_py["per_core_mean"], "model": cpu_py["model"]}
            else:
                cpu = dict(cpu_port, note=f"python reference unavailable: {cpu_py.get('unavailable')}")
            extras["cpu_port"] = cpu_port
            extras["cpu_baseline_python"] = cpu_py
            extras["cpu_baseline_python_rgb"] = python_reference_baseline(ENV_ID, "rgb", seconds=5.0)
        line = {
            "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": median_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config(args, world),
            "agent_steps_per_s": value * A,
            "timing": {"repeats": REPEATS, "steps_per_repeat": K, "timed_region_ms": total_ms, "repeat_ms": {"min": srt_rep[0], "median": median_ms, "max": srt_rep[-1]},
                       "ramp_steps": ramp_steps,
                       "note": "value = B*K / median repeat; each repeat = exactly K launches between its own CUDA event pair, repeats enqueued back to back"},
            "sustained": {"value": sustained, "ms_per_step": total_ms / (K * REPEATS),
                          "note": "mean over all repeats: includes the all-reset steps (every 100th step of a family regenerates all its envs)"},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": act_bytes, "d2h_bytes_per_step": d2h,
                    "steps": KE, "api": "mg_engine_step (C ABI, pinned host buffers, synchronous; batch cut into 4 env ranges whose D2H copies overlap the next range's kernel)",
                    "copy_ceiling": {"value": world * B * KE / (copy_ms * 1e-3), "gbs": world * (d2h + act_bytes) * KE / (copy_ms * 1e-3) / 1e9,
                                     "note": "the same H2D / D2H transfers (same pinned buffers, slices and streams) without the kernel, all ranks concurrently"},
                    "frac_of_copy_ceiling": copy_ms / e2e_ms, "parity": e2e_parity, "numa": numa},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "fused2_kernel<OBS=1,V=7,A=3> (env.step + auto-reset + egocentric encode: the only launch of a step)",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * B, "avg_launch_ms": avg_step_s * 1e3,
                         "peak_source": peak_src,
                         "note": "algorithmic bytes = SURVEY.md 8(d) 1272 B/env-step x envs per launch; launch duration = median K-step repeat / K (back to back over env "
                                 "families larger than L2, launch gaps included)",
                         "sustained_frac": ALGO_BYTES_PER_ENV_STEP * B / (total_ms / (K * REPEATS) * 1e-3) / 1e9 / peak},
            "other_configs": other,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "wall_s": {"cold_loop": wall_cold},
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
