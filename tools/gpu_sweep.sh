#!/bin/bash
# BASELINE configs[4]: T x width sweep of the fused sampler (bf16 tensor path at 65 536 rows, both paths at 256 rows),
# plus the large-batch shapes of configs[2]/[3] on one GPU.
mkdir -p gpurun_out
out=gpurun_out/sweep.txt; : > $out
for T in 5 20 100; do for W in 256 512 1024; do
  for B in 65536 256; do
    steps=10; [ $T = 100 ] && steps=3
    timeout 300 python bench.py --T $T --width $W --batch $B --steps $steps --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('sample bf16 T=$T h=$W B=$B', round(d['ms_per_step'],4),'ms', round(d['value']/1e6,3),'M/s frac', round(d['roofline']['frac'],4))" >> $out
  done
  timeout 300 python bench.py --T $T --width $W --batch 256 --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('sample fp32 T=$T h=$W B=256', round(d['ms_per_step'],4),'ms', round(d['value']/1e6,3),'M/s')" >> $out
done; done
timeout 300 python bench.py --workload ascent --batch 262144 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('ascent bf16 B=262144 (4 modes)', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,3),'M/s frac', round(d['roofline']['frac'],4))" >> $out
timeout 300 python bench.py --workload train --batch 262144 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('train bf16 B=262144', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,3),'M/s frac', round(d['roofline']['frac'],4))" >> $out
cat $out
