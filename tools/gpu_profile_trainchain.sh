#!/bin/bash
# one ncu --set full capture of the fused training forward after a plain run of the same command, then the GPU tests
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload train --batch 131072"
$B > gpurun_out/plain_train.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:actor_train_chain -s 3 -c 1 -f -o gpurun_out/prof_trainchain_r02 $B > gpurun_out/ncu_trainchain.log 2>&1
echo "train chain rc $?"
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 2
python bench.py --workload train --batch 131072 --steps 20 --warmup 5 > gpurun_out/bench_train_n1_final.json 2> gpurun_out/bench_train_n1_final.err; echo "bench rc $?"
