#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for wl in train ascent; do
  python bench.py --workload $wl --precision bf16 --batch 65536 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$wl.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$wl.csv \
      python bench.py --workload $wl --precision bf16 --batch 65536 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$wl.log 2>&1
  echo "$wl ncu exit $?"
done
