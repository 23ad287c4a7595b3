#!/bin/bash
# Run on the B200 box via:  gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tests] [bench] [ncu] [cfg4]'
# Everything it produces lands in gpurun_out/.
set -u
mkdir -p gpurun_out
what="${*:-tests bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [[ " $what " == *" tests "* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests exit $?" | tee -a gpurun_out/tests.log
  tail -3 gpurun_out/tests.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
fi
if [[ " $what " == *" bench "* ]]; then
  timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
  cat gpurun_out/bench.json
fi
if [[ " $what " == *" refarm "* ]]; then
  timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
  cat gpurun_out/bench_ref.json
fi
if [[ " $what " == *" cfg4 "* ]]; then
  timeout 600 python bench.py --workload cfg4 --steps 50 --warmup 5 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 exit $?"
  cat gpurun_out/bench_cfg4.json
  timeout 600 python bench.py --workload cfg2 --steps 200 --warmup 20 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
  cat gpurun_out/bench_cfg2.json
fi
if [[ " $what " == *" ncu "* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 100 --warmup 10 --profile --e2e-steps 3 > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused2?_kernel -s 40 -c 2 -f -o gpurun_out/prof \
      python bench.py --steps 20 --warmup 10 --profile --e2e-steps 3 > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out/
fi
if [[ " $what " == *" r2first "* ]]; then
  # round 2, first call: the new big-shape parity tests, the bench line + reference arm, launch-shape experiments
  timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/tests.log 2>&1; echo "tests exit $?" | tee -a gpurun_out/tests.log
  tail -14 gpurun_out/tests.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
  cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
  timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
  cat gpurun_out/bench_ref.json
fi
if [[ " $what " == *" r2b "* ]]; then
  # parity, the bench line, then one full ncu capture of the hot kernel and of the all-reset launch
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests exit $?" | tee -a gpurun_out/tests.log
  tail -5 gpurun_out/tests.log
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
  python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench.json") if l.startswith("{")][-1])
print("value %.3e  %.2f us/step  frac %.3f  sustained %.2f us  warm %.2f us  persistent %.2f us  all_reset %.1f us  desync %.2f us (x%.2f)  e2e %.3e" % (
    d["value"], 1e3 * d["ms_per_step"], d["roofline"]["frac"], 1e3 * d["sustained"]["ms_per_step"], 1e3 * d["warm"]["ms_per_step"],
    1e3 * d["rollout_persistent"]["ms_per_step"], d["all_reset_us"], d["desync"]["us_per_step"], d["desync"]["vs_lockstep_warm"], d["e2e"]["value"]), "pregen hit", d["desync"].get("pregen_hit_rate"))
print({k: (round(1e3 * v["ms_per_step"], 2), round(v["roofline"]["frac"], 3)) for k, v in d["other_configs"].items()})
PY
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused2_kernel -s 40 -c 2 -f -o gpurun_out/prof \
      python bench.py --steps 20 --warmup 10 --profile --e2e-steps 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu exit $?"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused2_kernel -s 99 -c 1 -f -o gpurun_out/prof_reset \
      python bench.py --steps 6 --warmup 120 --replicas 1 --profile --e2e-steps 3 > gpurun_out/ncu_reset.log 2>&1; echo "ncureset exit $?"
fi
if [[ " $what " == *" r2c "* ]]; then
  # parity + the bench line (no ncu)
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests exit $?" | tee -a gpurun_out/tests.log
  tail -3 gpurun_out/tests.log
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
  python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench.json") if l.startswith("{")][-1])
print("value %.3e  %.2f us/step  frac %.3f  sustained %.2f us  warm %.2f us  persistent %.2f us  all_reset %.1f us  desync %.2f us (x%.2f)  e2e %.3e" % (
    d["value"], 1e3 * d["ms_per_step"], d["roofline"]["frac"], 1e3 * d["sustained"]["ms_per_step"], 1e3 * d["warm"]["ms_per_step"],
    1e3 * d["rollout_persistent"]["ms_per_step"], d["all_reset_us"], d["desync"]["us_per_step"], d["desync"]["vs_lockstep_warm"], d["e2e"]["value"]), "pregen hit", d["desync"].get("pregen_hit_rate"))
print({k: (round(1e3 * v["ms_per_step"], 2), round(v["roofline"]["frac"], 3)) for k, v in d["other_configs"].items()}, "python_api us", round(1e3 * d["python_api"]["ms_per_step"], 2), round(1e3 * d["python_api"]["host_issue_ms_per_step"], 2))
PY
fi
if [[ " $what " == *" pgevery "* ]]; then
  for ev in 1 2 4 8 16; do echo "== MG_PREGEN_EVERY=$ev"; MG_PREGEN_EVERY=$ev timeout 120 python tools/desync_profile.py 1200; done 2>&1 | tee gpurun_out/pregen_every.log
fi
if [[ " $what " == *" pgprobe2 "* ]]; then
  timeout 300 python tools/pregen_probe2.py 2>&1 | tee gpurun_out/pregen_probe2.log
fi
if [[ " $what " == *" pgprobe "* ]]; then
  timeout 300 python tools/pregen_probe.py 2>&1 | tee gpurun_out/pregen_probe.log
fi
if [[ " $what " == *" ncudesync "* ]]; then
  timeout 120 python tools/desync_profile.py 400 | tee gpurun_out/desync_profile.log
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused2_kernel -s 120 -c 1 -f -o gpurun_out/prof_desync \
      python tools/desync_profile.py 130 > gpurun_out/ncu_desync.log 2>&1; echo "ncudesync exit $?"
fi
if [[ " $what " == *" exp "* ]]; then
  # quick A/B of launch-shape knobs of the specialised fused kernel (no CPU baseline)
  for knobs in ${EXP_KNOBS:-"MG_F2_CTAS_PER_SM=7" "MG_F2_NST=1,MG_F2_ONE_TILE=1" "MG_F2_NST=1,MG_F2_RAGGED=1" "MG_F2_NST=1" "MG_F2_ONE_TILE=1" "MG_F2_PDL=0"}; do
    echo "== $knobs"
    env ${knobs//,/ } MG_F2_VERBOSE=1 timeout 300 python bench.py --steps 200 --warmup 60 --quick --e2e-steps 3 2>gpurun_out/exp.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); t = d['timing']['repeat_ms']; print('value %.3e  median %.2f us  min %.2f  max %.2f us/step  sustained %.2f' % (d['value'], 1e3*d['ms_per_step'], 1e3*t['min']/d['steps'], 1e3*t['max']/d['steps'], 1e3*d['sustained']['ms_per_step']))
"
    sort -u gpurun_out/exp.err | head -3
  done 2>&1 | tee gpurun_out/exp.log
fi
if [[ " $what " == *" ncureset "* ]]; then
  # the step on which every episode ends (step_count hits max_steps = 100): launch 100 of the single-family warm-up
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused2?_kernel -s 99 -c 1 -f -o gpurun_out/prof_reset \
      python bench.py --steps 6 --warmup 120 --replicas 1 --profile --e2e-steps 3 > gpurun_out/ncu_reset.log 2>&1
  ls -la gpurun_out/
fi
if [[ " $what " == *" ncurgb "* ]]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused -s 6 -c 1 -f -o gpurun_out/prof_rgb \
      python bench.py --workload cfg4 --steps 4 --warmup 4 --profile > gpurun_out/ncu_rgb.log 2>&1
  ls -la gpurun_out/
fi
if [[ " $what " == *" sanitize "* ]]; then
  # memory-safety and shared-memory hazard checks of the kernels on a small problem (tools/sanitize_workload.py)
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_workload.py 103 > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/sanitizer_memcheck.log
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_workload.py 103 > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/sanitizer_racecheck.log
  timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python tools/sanitize_workload.py 103 > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck exit $?" | tee -a gpurun_out/sanitizer_synccheck.log
  for f in memcheck racecheck synccheck; do tail -n 3 gpurun_out/sanitizer_$f.log; done
fi
if [[ " $what " == *" rollout "* ]]; then
  # K steps per launch (mg_rollout_persistent) next to K launches (mg_rollout_fused), same family, state L2-resident
  MG_F2_VERBOSE=1 timeout 300 python - > gpurun_out/rollout.log 2>&1 <<'PY'
import torch, time
from marlgrid_b200 import envs
B, T = 65536, 100
env = envs.make("MarlGrid-3AgentCluttered15x15-v0", num_envs=B, obs_mode="encoded", seed=1337)
env.reset()
act = torch.stack([env.random_actions(t) for t in range(T)])
out = (torch.empty((T, B, 3, 7, 7, 3), dtype=torch.uint8, device="cuda"), torch.empty((T, B, 3), dtype=torch.float64, device="cuda"), torch.empty((T, B), dtype=torch.bool, device="cuda"))
for name, fn in (("persistent (1 launch / 100 steps)", lambda: env.rollout_all(act, out=out)), ("fused (100 launches)", lambda: env.rollout(act))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * T)
    print(f"{name}: {us:.2f} us per step, {B / us * 1e6:.3e} env-steps/s")
PY
  grep -v "^fused2<" gpurun_out/rollout.log; grep "^fused2<" gpurun_out/rollout.log | sort -u | head -3
fi
if [[ " $what " == *" soak "* ]]; then
  timeout 1200 python tools/soak.py > gpurun_out/soak.log 2>&1; echo "soak exit $?" | tee -a gpurun_out/soak.log
  cat gpurun_out/soak.log | tail -8
fi
if [[ " $what " == *" desync "* ]]; then
  timeout 600 python tools/desync_probe.py > gpurun_out/desync.log 2>&1; cat gpurun_out/desync.log | tail -10
fi
if [[ " $what " == *" final "* ]]; then
  # last call of a round on a tight budget: parity first, then one full ncu capture of the hot kernel, the bench line, the desync probe
  timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests exit $?" | tee -a gpurun_out/tests.log
  tail -2 gpurun_out/tests.log
  timeout 60 ncu --set full --clock-control none --import-source on -k regex:fused2_kernel -s 40 -c 2 -f -o gpurun_out/prof \
      python bench.py --steps 20 --warmup 10 --profile --e2e-steps 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu exit $?"
  timeout 90 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
  cat gpurun_out/bench.json
  timeout 40 python tools/desync_probe.py > gpurun_out/desync.log 2>&1; tail -8 gpurun_out/desync.log
fi
if [[ " $what " == *" policy "* ]]; then
  # closed-loop rollout with the on-device policy: parity against the CPU statement, then its cost next to the open-loop kernel
  timeout 600 python -m pytest tests -m gpu -x -q -k "policy or persistent_rollout or trainer" > gpurun_out/tests_policy.log 2>&1; echo "policy tests exit $?" | tee -a gpurun_out/tests_policy.log
  tail -15 gpurun_out/tests_policy.log
  timeout 300 python tools/policy_probe.py 2>&1 | tee gpurun_out/policy_probe.log
  timeout 300 python tools/train_linear_q.py 150 16384 2>&1 | tee gpurun_out/train_linear_q.log
fi
if [[ " $what " == *" ncupolicy "* ]]; then
  # one full capture of the K-steps-per-launch kernel with the policy hand-off (launch 7 of the probe = the first timed closed-loop rollout)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused2_kernel -s 7 -c 1 -f -o gpurun_out/prof_policy \
      python tools/policy_probe.py > gpurun_out/ncu_policy.log 2>&1; echo "ncupolicy exit $?"
  ls -la gpurun_out/*.ncu-rep
fi
if [[ " $what " == *" features "* ]]; then
  timeout 300 python -m pytest tests -m gpu -x -q -k "hide or golden or traj" > gpurun_out/tests_features.log 2>&1; echo "feature tests exit $?"; tail -2 gpurun_out/tests_features.log
  timeout 300 python tools/feature_probe.py 2>&1 | tee gpurun_out/feature_probe.log
fi
