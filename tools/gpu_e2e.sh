#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -q --timeout 120 -k "host" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_e2e.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"; tail -3 gpurun_out/bench_e2e.err
