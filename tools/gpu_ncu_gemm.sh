#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload train --precision bf16 --batch 65536 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_t.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"row_gemm_kernel|dw_gemm_kernel" -s 60 -c 12 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
python __graft_entry__.py smoke 2>&1 | tail -2
