#!/bin/bash
# Timing ablations of the fused critic chain (csrc/q_chain_tc.cu, DDP_QC_ABLATE): per-tile phase cycles per variant.
for a in "$@"; do
  echo "=== variant $a"
  DDP_LIB_PATH=/root/repo/ddiffpg_b200/libvar_$a.so timeout 120 python tools/qc_timing.py 131072 2>&1 | head -12
done
