#!/bin/bash
# BASELINE configs[4]: T in {5,20,100} x trunk width in {256,512,1024}, fused bf16 sampler, 65 536 states, one GPU.
for h in 256 512 1024; do for T in 5 20 100; do
  python bench.py --no-secondary --no-cpu-baseline --T $T --width $h --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(f\"h={$h:5d} T={$T:4d}  {d['ms_per_step']:9.3f} ms  {d['value']/1e6:8.2f} M actions/s  {d['roofline']['achieved']:7.1f} TFLOP/s  {100*d['roofline']['frac']:5.1f} % of bf16 peak   e2e {d['e2e']['ms_per_step']:8.3f} ms\")"
done; done
