#!/bin/bash
# H3 training step: parity tests, bench (graph + eager) and launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 200 -k "train or trainer or gemm or critic or rnd or q_" > gpurun_out/pytest_h3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_h3.log; tail -3 gpurun_out/pytest_h3.log
for wl in train ascent; do
timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_${wl}_bf16.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl', round(d['ms_per_step'],4), 'ms', round(d['value']/1e6,1), 'M/s frac', round(d['roofline']['frac'],4))"
done
if [ -n "$NCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_train.csv python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_train.log 2>&1
python tools/launch_summary.py gpurun_out/launches_train.csv | head -14
fi
