#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm_gpu.py tests/test_tc_gpu.py -m gpu -q --timeout 120 -k "gemm or q_ or train or trainer" > gpurun_out/pytest_var.log 2>&1; grep -E "^E  .*(Error|assert)|passed|failed|^FAILED" gpurun_out/pytest_var.log | head
for wl in train ascent; do timeout 300 python bench.py --workload $wl --batch 65536 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/err_$wl.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', round(d['ms_per_step'],3), 'ms', d['value'])"; tail -2 gpurun_out/err_$wl.txt; done
