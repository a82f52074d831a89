#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_tc_gemm_gpu.py tests/test_parity_gpu.py -m gpu -q --timeout 120 -k "q_ or train or trainer or gemm" > gpurun_out/pytest_h23.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_h23.log
grep -E "^E  .*(Error|assert)|passed|failed|^FAILED" gpurun_out/pytest_h23.log | head -20
for wl in train ascent; do timeout 300 python bench.py --workload $wl --precision bf16 --batch 65536 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_${wl}_bf16.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl bf16', d['value'], d['ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/bench_${wl}_bf16.err; done
