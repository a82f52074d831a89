#!/bin/bash
# A/B of the rearranged Mish / Mish' of the fused training forward (-DDDP_FC_MISH_V2=1 build under tools/ab/)
mkdir -p gpurun_out
B="python bench.py --workload train --batch 131072 --no-cpu-baseline --steps 20 --warmup 5"
for v in 0 1 0 1; do
  if [ "$v" = 0 ]; then unset DDP_LIB_PATH; else export DDP_LIB_PATH=/root/repo/tools/ab/libfc_mishv2.so; fi
  timeout 150 $B > gpurun_out/abfc_v$v.json 2> gpurun_out/abfc_v$v.err
  echo "variant $v rc $?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/abfc_v$v.json | head -1
done
export DDP_LIB_PATH=/root/repo/tools/ab/libfc_mishv2.so
timeout 300 python -m pytest tests/test_tc_gpu.py -q -m gpu -k "train or trainer" 2>&1 | tail -n 2
