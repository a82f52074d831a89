#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -q --timeout 120 -k "q_" > gpurun_out/pytest_h2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_h2.log
grep -E "^E  .*(Error|assert)|passed|failed|^FAILED" gpurun_out/pytest_h2.log | head -20
for prec in fp32 bf16; do timeout 300 python bench.py --workload ascent --precision $prec --batch 65536 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_ascent_$prec.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$prec', d['value'], d['ms_per_step'], d['roofline']['frac'])"; tail -2 gpurun_out/bench_ascent_$prec.err; done
