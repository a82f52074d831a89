#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload ascent --batch 65536 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_ascent.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:q_chain_tc -s 8 -c 1 -f -o gpurun_out/prof_qchain \
  python bench.py --workload ascent --batch 65536 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_qchain.log 2>&1
tail -3 gpurun_out/ncu_qchain.log; ls -la gpurun_out/prof_qchain.ncu-rep
