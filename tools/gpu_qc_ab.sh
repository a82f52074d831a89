#!/bin/bash
for v in "$@"; do
  lib=variants/lib_$v.so
  [ "$v" = base ] && lib=ddiffpg_b200/libddiffpg_b200.so
  echo "=== $v"
  DDP_LIB_PATH=$lib timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -q --timeout 120 -x -k "q_" 2>&1 | tail -2
  DDP_LIB_PATH=$lib timeout 200 python tools/qc_timing.py 65536 2>&1 | head -13
done
