#!/bin/bash
# H1 sampler: parity tests, per-phase cycles and bench for the main build and every A/B variant build.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tc_gpu.py -m gpu -q --timeout 120 -k "sampler or selftest or get_actions" > gpurun_out/pytest_h1.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_h1.log; tail -4 gpurun_out/pytest_h1.log
timeout 120 python tools/tc_timing.py 2>&1 | tee gpurun_out/h1_timing.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_h1.json | cut -c1-400
for v in ddiffpg_b200/libvariant_*.so; do
  [ -f "$v" ] || continue
  echo "== $v"
  DDP_LIB_PATH=$PWD/$v timeout 120 python tools/tc_timing.py 2>&1 | tee gpurun_out/h1_timing_$(basename $v .so).txt
  DDP_LIB_PATH=$PWD/$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_h1_$(basename $v .so).json | cut -c1-400
done
for b in 75776 151552; do
  timeout 300 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_h1_b$b.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch $b', round(d['ms_per_step'],4), 'ms', round(d['value']/1e6,1), 'M/s frac', round(d['roofline']['frac'],4))"
done
