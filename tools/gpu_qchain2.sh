#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/qc_timing.py 65536 2>&1 | tail -14
timeout 300 python bench.py --workload ascent --batch 65536 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_ascent.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_ascent.csv \
  python bench.py --workload ascent --batch 65536 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ascent.log 2>&1
python tools/launch_summary.py gpurun_out/launches_ascent.csv | head -14
