#!/bin/bash
# one ncu --set full capture of the seven row GEMM launches of a training step (after warm-up)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gemm_gpu.py tests/test_tc_gpu.py -m gpu -q --timeout 200 -k "train or trainer or gemm" 2>&1 | tail -2
timeout 300 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_train_bf16.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train', round(d['ms_per_step'],4), 'ms', round(d['value']/1e6,1), 'M/s frac', round(d['roofline']['frac'],4))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_gemm -s 14 -c 7 -f -o gpurun_out/prof_rowgemm python bench.py --workload train --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rowgemm.log 2>&1
tail -2 gpurun_out/ncu_rowgemm.log
