#!/bin/bash
# One gpurun call: GPU tests + smoke + bench (+ optional ncu launch list).  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
PREC=${1:-fp32}
timeout 600 python bench.py --steps 10 --warmup 3 --precision $PREC > gpurun_out/bench_$PREC.json 2> gpurun_out/bench_$PREC.err; echo "bench exit $?"; cat gpurun_out/bench_$PREC.json; tail -3 gpurun_out/bench_$PREC.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
