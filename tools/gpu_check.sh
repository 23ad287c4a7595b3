#!/bin/bash
# Run on the B200 box via:  gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tests] [bench] [ncu] [cfg4]'
# Everything it produces lands in gpurun_out/.
set -u
mkdir -p gpurun_out
what="${*:-tests bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [[ "$what" == *tests* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests exit $?" | tee -a gpurun_out/tests.log
  tail -3 gpurun_out/tests.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
fi
if [[ "$what" == *bench* ]]; then
  timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
  cat gpurun_out/bench.json
fi
if [[ "$what" == *refarm* ]]; then
  timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
  cat gpurun_out/bench_ref.json
fi
if [[ "$what" == *cfg4* ]]; then
  timeout 600 python bench.py --workload cfg4 --steps 50 --warmup 5 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 exit $?"
  cat gpurun_out/bench_cfg4.json
  timeout 600 python bench.py --workload cfg2 --steps 200 --warmup 20 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
  cat gpurun_out/bench_cfg2.json
fi
if [[ "$what" == *ncu* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 40 -c 2 -f -o gpurun_out/prof \
      python bench.py --steps 20 --warmup 10 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out/
fi
