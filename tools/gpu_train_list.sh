#!/bin/bash
mkdir -p gpurun_out
python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_train.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_train.log 2>&1
python tools/launch_summary.py gpurun_out/launches_train.csv | head -30
