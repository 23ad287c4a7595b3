#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched MarlGrid hot path (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
  torchrun --nproc-per-node N ... bench.py --gpus N ...         # N ranks, one per GPU, env-index sharded
  python bench.py --impl reference ...                          # CPU arm: the oracle port on the host cores

A "step" is one env.step() of the whole batch: MarlGrid-3AgentCluttered15x15-v0, 65 536 envs per GPU,
encoded observations, uniform random actions, auto-reset (BASELINE.json configs[2]; 8 GPUs = the
sharded family of configs[4]).  One JSON line is printed by rank 0.

Timing: W warm-up steps, then K steps back to back between ONE pair of CUDA events on the launching
stream.  The steps visit `--replicas` (6) independent env families of 65 536 envs round robin, so a
family's state has been evicted from the 126 MB L2 by the time it is stepped again ("inputs larger
than L2"; no flush kernel between launches): `value` = B*K / device time, max over ranks.
`step_ms_flushed` is the distribution of single steps timed with a 256 MiB flush before each;
`warm` repeats K steps on one family (state L2-resident: what a rollout loop at this batch sees).  `e2e` drives the C-ABI host
buffer engine (mg_engine_step): pinned host actions in, obs/rewards/done out, copies inside the timer.
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
# other BASELINE.json configs, runnable for the record with --workload (the driver's line is always cfg3):
#   name -> (env id, default batch, obs mode, agents, algorithmic bytes per env-step [SURVEY.md 8(d)])
WORKLOADS = {
    "cfg3": ("MarlGrid-3AgentCluttered15x15-v0", 65536, "encoded", 3, 1272),
    "cfg2": ("MarlGrid-3AgentCluttered11x11-v0", 4096, "encoded", 3, 957),
    "cfg4": ("MarlGrid-4AgentEmpty9x9-v0", 262144, "rgb", 4, 38070),
}
ALGO_BYTES_PER_ENV_STEP = 1272  # 743 read + 522 write + 7 amortised reset (whole env.step)
ALGO_BYTES_OBS_KERNEL = 1164    # SURVEY.md 8(d) obs-kernel-only figure: read 3*W*H + 16*A = 723, write obs 441
FALLBACK_HBM_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md fallback


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


def cpu_port_throughput(n_envs, threads, target_s=12.0, seed=1337):
    """The oracle (CPU port of the reference algorithm, oracle/mg_oracle.c) on a bounded sample of the same
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


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm's CPU implementation (oracle port), all host threads."""
    if rank != 0:
        return
    threads = host_threads()
    n_envs = 65536 if threads >= 32 else 8192  # a bench "step" of this arm = one env.step over this sample of the batch
    v, dt, n_steps = cpu_port_throughput(n_envs, threads, target_s=30.0)
    sample = (f"{n_envs} envs x {n_steps} steps ({dt:.1f} s) of the same workload, C port of the reference step+reset+encode "
              f"(oracle/mg_oracle.c), {threads} threads")
    line = {
        "impl": "reference", "metric": "env-steps/s", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * n_envs / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "agent_steps_per_s": 3 * v,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {
        "workload": f"{ENV_ID} batch={args.batch_per_gpu}/GPU x {world} GPU(s), encoded obs [B,3,7,7,3] u8, uniform random actions, auto-reset "
                    f"(BASELINE.json configs[2]{'; sharded family of configs[4]' if world > 1 else ''})",
        "env_id": ENV_ID, "batch_per_gpu": args.batch_per_gpu, "global_batch": args.batch_per_gpu * world, "n_agents": 3,
        "parallelism": f"env-index sharding x{world}, no collective on the data path",
        "l2": f"inputs larger than L2: {getattr(args, 'replicas', 6)} independent env families of batch_per_gpu envs stepped round robin (each step touches ~51 MB, "
              "a family is revisited after > 126 MB of other traffic); K steps back to back, one CUDA event pair",
    }


def run_other_workload(args):
    """Device-side throughput of another BASELINE config (cold, per-step events) -- informational line."""
    import torch

    from marlgrid_b200 import envs

    env_id, B, mode, A, algo = WORKLOADS[args.workload]
    if args.batch_per_gpu != 65536:
        B = args.batch_per_gpu
    env = envs.make(env_id, num_envs=B, obs_mode=mode, seed=1337)
    env.reset()
    acts = [env.random_actions(t) for t in range(32)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=env.device)
    for t in range(args.warmup):
        env.step(acts[t % 32])
    K = args.steps
    st = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    en = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    torch.cuda.synchronize()
    for t in range(K):
        flush.fill_(t & 0xFF)
        st[t].record()
        env.step(acts[t % 32])
        en[t].record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in zip(st, en))
    tot = sum(ms)
    peak, src = measured_peak()
    ach = algo * B / (tot / K * 1e-3) / 1e9
    print(json.dumps({"metric": "env-steps/s", "workload": args.workload, "env_id": env_id, "batch": B, "obs": mode, "value": B * K / (tot * 1e-3),
                      "agent_steps_per_s": A * B * K / (tot * 1e-3), "ms_per_step": tot / K, "step_ms_median": ms[K // 2], "steps": K,
                      "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                   "algorithmic_bytes_per_launch": algo * B, "peak_source": src}}), flush=True)


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
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload != "cfg3":
        return run_other_workload(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

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
        os.environ["NCCL_DEBUG"] = "WARN"  # NCCL's version banner goes to stdout: rank 0 must print ONE JSON line
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    B, K, W = args.batch_per_gpu, args.steps, args.warmup
    L = _lib.load()

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
    states = (MgState * R)(*[f._state for f in fams])
    PP = ctypes.c_void_p * R
    rew_p, done_p, obs_p = PP(*[f.rewards.data_ptr() for f in fams]), PP(*[f.done.data_ptr() for f in fams]), PP(*[f.obs.data_ptr() for f in fams])
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def rollout_rr(n_steps):
        done_steps = 0
        while done_steps < n_steps:  # chunks of POOL steps (a multiple of R: the round robin continues seamlessly)
            n = min(POOL - POOL % R, n_steps - done_steps)
            _lib.check(L.mg_rollout_fused_rr(ctypes.byref(env.cfg), states, R, actions.data_ptr(), n, rew_p, done_p, obs_p, 1, stream), "mg_rollout_fused_rr")
            done_steps += n

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up -------------------------------------------------------------------------------
    rollout_rr(max(W, R) // R * R)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed: K cold steps back to back, one event pair -----------------------------------------
    K = max(K // R, 1) * R
    launches0 = L.mg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.perf_counter()
    e0.record()
    rollout_rr(K)  # ONE launch per step: fused step + auto-reset + observe kernel
    e1.record()
    barrier()
    wall_cold = time.perf_counter() - wall0
    launches = L.mg_launch_count() - launches0
    cold_total_ms = float(e0.elapsed_time(e1))

    # ---- secondary: per-step events with a 256 MiB L2 flush before each step (distribution of single cold steps) ----
    KF = min(K, 300)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(KF)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(KF)]
    for t in range(KF):
        flush.fill_(t & 0xFF)
        starts[t].record()
        env.step(actions[t % POOL])
        stops[t].record()
    barrier()
    cold_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]

    # ---- timed: K warm steps back to back on ONE family (state L2-resident: what a rollout loop at this batch sees) ----
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done_steps = 0
    while done_steps < K:
        n = min(POOL, K - done_steps)
        env.rollout(actions[:n])
        done_steps += n
    e1.record()
    barrier()
    warm_total_ms = e0.elapsed_time(e1)

    # ---- on-device rollout loop: 100 steps per launch (mg_rollout_persistent), every step's outputs written to its own slice ----
    TP = 100
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
    persistent_ms = e0.elapsed_time(e1) / (5 * TP)
    del pout

    # ---- the same K warm steps through the Python surface, env.step(actions) called in a Python loop -------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for t in range(K):
        env.step(actions[t % POOL])
    e1.record()
    host_issue_s = time.perf_counter() - t0  # time the host needed to enqueue K steps
    barrier()
    py_total_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    del fams[1:]

    # ---- e2e: host buffers through the C ABI engine ----------------------------------------------
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
    for t in range(5):
        _lib.check(L.mg_engine_step(h, p_acts[t % 8], p_obs, p_rew, p_done, 1), "mg_engine_step")
    barrier()
    t0 = time.perf_counter()
    KE = args.e2e_steps
    for t in range(KE):
        _lib.check(L.mg_engine_step(h, p_acts[t % 8], p_obs, p_rew, p_done, 1), "mg_engine_step")
    barrier()
    e2e_s = time.perf_counter() - t0
    obs_last = np.ctypeslib.as_array(ctypes.cast(p_obs, ctypes.POINTER(ctypes.c_uint8)), shape=(obs_bytes,))
    e2e_checksum = int(obs_last[:: 4099].astype(np.int64).sum())
    L.mg_engine_destroy(h)

    # ---- max over ranks --------------------------------------------------------------------------
    times = torch.tensor([cold_total_ms, warm_total_ms, e2e_s * 1e3, persistent_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    cold_total_ms, warm_total_ms, e2e_ms, persistent_ms = (float(x) for x in times.tolist())

    if rank == 0:
        peak, peak_src = measured_peak()
        value = world * B * K / (cold_total_ms * 1e-3)
        warm_value = world * B * K / (warm_total_ms * 1e-3)
        e2e_value = world * B * KE / (e2e_ms * 1e-3)
        srt = sorted(cold_ms)
        avg_step_s = (cold_total_ms / K) * 1e-3           # one launch per step, back to back: the kernel's average launch duration (launch gaps included)
        med_step_s = srt[len(srt) // 2] * 1e-3            # flushed single step on which no episode ends (99 of 100); event resolution ~2 us
        achieved = ALGO_BYTES_PER_ENV_STEP * B / avg_step_s / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("fused_kernel_dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            thr = host_threads()
            n_envs = 65536 if thr >= 32 else 8192
            v, dt, n_steps = cpu_port_throughput(n_envs, thr, target_s=22.0)
            cpu = {"value": v, "unit": "env-steps/s", "cores": thr, "kind": "port",
                   "sample": f"{n_envs} envs x {n_steps} steps of the same workload ({dt:.1f} s), C port of the reference step+reset+encode (oracle/mg_oracle.c)"}
        line = {
            "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": cold_total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config(args, world),
            "agent_steps_per_s": value * A,
            "step_ms_flushed": {"min": srt[0], "median": srt[len(srt) // 2], "p99": srt[min(len(srt) - 1, int(0.99 * len(srt)))], "max": srt[-1], "steps": len(srt),
                                "note": "secondary: single steps, 256 MiB L2 flush before each, per-step CUDA events (~2 us resolution)"},
            "warm": {"value": warm_value, "ms_per_step": warm_total_ms / K, "note": "K steps back to back, state L2-resident, launched from mg_rollout_fused"},
            "rollout_persistent": {"value": world * B / (persistent_ms * 1e-3), "ms_per_step": persistent_ms, "steps_per_launch": TP,
                                   "note": "on-device rollout loop (mg_rollout_persistent): 100 steps per launch on a fixed action tape, the tiles' state stays in "
                                           "shared memory between steps, every step's obs / rewards / done go to their own HBM slice (informational: an "
                                           "open-loop rollout; `value` is one launch per step)"},
            "python_api": {"value": world * B * K / (py_total_ms * 1e-3), "ms_per_step": py_total_ms / K, "host_issue_ms_per_step": 1e3 * host_issue_s / K,
                           "note": "env.step(actions) in a Python loop on one family (rank 0's figures; host-bound when host_issue_ms_per_step ~ ms_per_step)"},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": act_bytes, "d2h_bytes_per_step": obs_bytes + rew_bytes + B,
                    "steps": KE, "api": "mg_engine_step (C ABI, pinned host buffers, synchronous; batch cut into 4 env ranges whose D2H copies overlap the next range's kernel)", "checksum": e2e_checksum},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "fused2_kernel<V=7,A=3> (env.step + auto-reset + egocentric encode: the only launch of a step)",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * B, "avg_launch_ms": avg_step_s * 1e3,
                         "peak_source": peak_src,
                         "note": "algorithmic bytes = SURVEY.md 8(d) 1272 B/env-step x 65536; K launches back to back over env families larger than L2, one CUDA event pair",
                         "flushed_median_launch": {"ms": med_step_s * 1e3, "frac": ALGO_BYTES_PER_ENV_STEP * B / med_step_s / 1e9 / peak}},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "wall_s": {"cold_loop": wall_cold},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
