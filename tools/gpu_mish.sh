#!/bin/bash
# A/B of the Mish epilogue variants of the fused sampler: per-phase clocks + bench value + numerics.
mkdir -p gpurun_out
for v in "$@"; do
  lib=variants/lib_$v.so
  [ "$v" = base ] && lib=ddiffpg_b200/libddiffpg_b200.so
  echo "=== $v"
  DDP_LIB_PATH=$lib timeout 200 python tools/tc_timing.py 65536 2>&1 | tail -8
  DDP_LIB_PATH=$lib timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', round(d['ms_per_step'],4), 'ms', round(d['value']/1e6,2), 'M/s', d['roofline']['frac'])"
  DDP_LIB_PATH=$lib timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -q --timeout 120 -k "sampler" 2>&1 | tail -2
done
