#!/bin/bash
# H2 critic chain + the round's new tests, then the ascent / train bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 200 -k "q_ or ascent or layerwise or requires_its_workspace or critic" > gpurun_out/pytest_h2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_h2.log; tail -3 gpurun_out/pytest_h2.log
for wl in ascent train; do
timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_${wl}_bf16.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl', round(d['ms_per_step'],4), 'ms', round(d['value']/1e6,1), 'M/s frac', round(d['roofline']['frac'],4))"
done
timeout 100 python tools/qc_timing.py 2>&1 | head -13
