#!/bin/bash
# A/B of the ELU' scratch handling of the fused critic chain (DDP_QC_SCRATCH builds under tools/ab/): bench line,
# DRAM bytes of one launch, parity tests of the ascent on each variant.
mkdir -p gpurun_out
B="python bench.py --workload ascent --batch 131072 --no-cpu-baseline"
for v in "$@"; do
  if [ "$v" = 0 ]; then unset DDP_LIB_PATH; else export DDP_LIB_PATH=/root/repo/tools/ab/libqc_s$v.so; fi
  timeout 150 $B --steps 10 --warmup 3 > gpurun_out/ab_s$v.json 2> gpurun_out/ab_s$v.err
  echo "variant $v bench rc $?"
  timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:q_chain_tc -s 30 -c 2 --csv --log-file gpurun_out/ab_ncu_s$v.csv $B --steps 2 --warmup 3 > gpurun_out/ab_ncu_s$v.log 2>&1
  echo "variant $v ncu rc $?"
  if [ "$v" != 0 ]; then
    timeout 200 python -m pytest tests/test_tc_gpu.py -q -m gpu -k "q_" > gpurun_out/ab_pytest_s$v.log 2>&1
    echo "variant $v pytest rc $?"
  fi
done
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/ab_s*.json | head -20
tail -n 3 gpurun_out/ab_ncu_s*.csv
tail -n 2 gpurun_out/ab_pytest_s*.log
