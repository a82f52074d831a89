#!/bin/bash
# H1 sampler: parity tests, per-phase cycles and bench lines at the sizes that matter
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tc_gpu.py tests/test_parity_gpu.py -m gpu -q --timeout 120 -k "sampler or selftest or get_actions or actor or sample" > gpurun_out/pytest_h1.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_h1.log; tail -4 gpurun_out/pytest_h1.log
timeout 120 python tools/tc_timing.py 2>&1 | tee gpurun_out/h1_timing.txt
for b in 65536 256 4096 9472 75776; do
  timeout 300 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_h1_b$b.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch $b', round(d['ms_per_step'],4), 'ms', round(d['value']/1e6,1), 'M/s frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e6,2))"
done
DDP_TC_NO_HALF_TILES=1 timeout 300 python bench.py --batch 256 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch 256, 128-row tiles', round(d['ms_per_step'],4), 'ms')"
DDP_TC_NO_HALF_TILES=1 timeout 300 python bench.py --batch 4096 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch 4096, 128-row tiles', round(d['ms_per_step'],4), 'ms')"
