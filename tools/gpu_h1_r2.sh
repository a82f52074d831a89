#!/bin/bash
# H1 round 2: packed-fp16 Mish variant vs the round-1 plan (legacy .so), micro-benchmark, parity, phase cycles
mkdir -p gpurun_out
./tools/micro/mish_bw 2>&1 | tee gpurun_out/micro_mish_bw_r2.txt
timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -q --timeout 200 -k "sampler or get_actions" > gpurun_out/pytest_h1.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_h1.log; tail -6 gpurun_out/pytest_h1.log
timeout 120 python tools/tc_timing.py 2>&1 | tee gpurun_out/h1_timing_h2.txt
DDP_LIB_PATH=$PWD/ddiffpg_b200/libddp_legacy.so timeout 120 python tools/tc_timing.py 2>&1 | tee gpurun_out/h1_timing_legacy.txt
for b in 65536 75776; do
  timeout 300 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_h1_h2_b$b.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('H2     batch $b', round(d['ms_per_step'],4), 'ms', round(d['value']/1e6,1), 'M/s frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e6,2))"
  DDP_LIB_PATH=$PWD/ddiffpg_b200/libddp_legacy.so timeout 300 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('legacy batch $b', round(d['ms_per_step'],4), 'ms', round(d['value']/1e6,1), 'M/s frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e6,2))"
done
timeout 600 python tools/parity_probe.py h1 h3 2>&1 | grep -E "^H1|^H3" | tee gpurun_out/parity_probe_r2b.txt
