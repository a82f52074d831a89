#!/bin/bash
# A/B of the evict_last policy on the sampler's state-partial scratch (DDP_TC_KEEP_PARTIAL build under tools/ab/).
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-secondary"
for v in 0 1 0 1; do
  if [ "$v" = 0 ]; then unset DDP_LIB_PATH; else export DDP_LIB_PATH=/root/repo/tools/ab/libh1_keep1.so; fi
  timeout 150 $B --steps 30 --warmup 5 > gpurun_out/abh1_k$v.json 2> gpurun_out/abh1_k$v.err
  echo "variant $v bench rc $?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/abh1_k$v.json | head -2
done
for v in 0 1; do
  if [ "$v" = 0 ]; then unset DDP_LIB_PATH; else export DDP_LIB_PATH=/root/repo/tools/ab/libh1_keep1.so; fi
  timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:actor_sample_tc -s 8 -c 1 --csv --log-file gpurun_out/abh1_ncu_k$v.csv $B --steps 2 --warmup 3 > gpurun_out/abh1_ncu_k$v.log 2>&1
  echo "variant $v ncu rc $?"; tail -n 3 gpurun_out/abh1_ncu_k$v.csv | awk -F'","' '{print $(NF-2), $NF}'
done
export DDP_LIB_PATH=/root/repo/tools/ab/libh1_keep1.so
timeout 300 python -m pytest tests/test_tc_gpu.py -q -m gpu -k "sampler or get_actions" > gpurun_out/abh1_pytest_k1.log 2>&1
echo "variant 1 pytest rc $?"; tail -n 2 gpurun_out/abh1_pytest_k1.log
unset DDP_LIB_PATH
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "main pytest rc $?"; tail -n 3 gpurun_out/pytest_gpu.log
