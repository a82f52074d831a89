#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -q --timeout 120 > gpurun_out/pytest_tc.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tc.log
tail -5 gpurun_out/pytest_tc.log
timeout 300 python tools/gpu_probe.py bf16 2>&1 | grep -E "H1|B200" | tee gpurun_out/probe_tc.log
for v in ddiffpg_b200/libvariant_*.so; do
  [ -f "$v" ] || continue
  DDP_LIB_PATH=$PWD/$v timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -q --timeout 120 2>&1 | tail -1
  DDP_LIB_PATH=$PWD/$v timeout 300 python tools/gpu_probe.py bf16 2>&1 | grep -E "H1" | sed "s|^|$(basename $v) |" | tee -a gpurun_out/probe_tc.log
done
