#!/bin/bash
# BASELINE configs[4], reduced: T x width sweep of the fused sampler on the bf16 tensor path at 65 536 rows, the
# reference width at full waves (75 776 rows), and the large-batch shapes of configs[2]/[3] on one GPU.
mkdir -p gpurun_out
out=gpurun_out/sweep2.txt; : > $out
run() { # T W B steps
  timeout 300 python bench.py --T $1 --width $2 --batch $3 --steps $4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('sample bf16 T=$1 h=$2 B=$3', round(d['ms_per_step'],4),'ms', round(d['value']/1e6,3),'M/s frac', round(d['roofline']['frac'],4))" >> $out
}
for T in 5 20 100; do for W in 256 512 1024; do
  steps=10; [ $T = 100 ] && steps=3
  run $T $W 65536 $steps
done; done
run 20 1024 75776 5
run 100 1024 75776 3
timeout 300 python bench.py --workload ascent --batch 262144 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('ascent bf16 B=262144 (4 modes)', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,3),'M/s frac', round(d['roofline']['frac'],4))" >> $out
timeout 300 python bench.py --workload train --batch 262144 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('train bf16 B=262144', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,3),'M/s frac', round(d['roofline']['frac'],4))" >> $out
timeout 300 python bench.py --workload train --batch 4096 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('train bf16 B=4096', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,3),'M/s frac', round(d['roofline']['frac'],4))" >> $out
cat $out
