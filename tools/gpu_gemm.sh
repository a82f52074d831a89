#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm_gpu.py -m gpu -q --timeout 120 > gpurun_out/pytest_gemm.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gemm.log
grep -E "^E  .*Error|^E  .*assert|passed|failed|^FAILED" gpurun_out/pytest_gemm.log | head -30
