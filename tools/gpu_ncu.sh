#!/bin/bash
# usage: gpu_ncu.sh <out-name> [lib path]
mkdir -p gpurun_out
[ -n "$2" ] && export DDP_LIB_PATH=$PWD/$2
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:actor_sample_tc -s 3 -c 1 -o gpurun_out/$1 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
