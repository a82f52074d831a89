#!/bin/bash
# One gpurun call: GPU tests + smoke + bench (+ ncu launch list with NCU=1).  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
if [ -n "$NCU" ]; then
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:actor_sample_tc -s 3 -c 2 -o gpurun_out/prof_sampler_tc \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  tail -5 gpurun_out/ncu_full.log
fi
