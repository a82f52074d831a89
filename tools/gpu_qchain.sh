#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -q --timeout 120 -x -k "q_" > gpurun_out/pytest_qchain.log 2>&1; echo "pytest exit $?"; grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_qchain.log | head -20
for nc in 0 1; do DDP_Q_NO_CHAIN=$nc timeout 300 python bench.py --workload ascent --batch 65536 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/err_ascent.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ascent nochain=$nc', round(d['ms_per_step'],3), 'ms', d['value'])"; tail -2 gpurun_out/err_ascent.txt; done
