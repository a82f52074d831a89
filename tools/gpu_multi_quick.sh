#!/bin/bash
# quick torchrun check: sampler and training workloads must print their line AND exit 0 (usage: gpu_multi_quick.sh N)
mkdir -p gpurun_out
N=${1:-2}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "sample exit $?"; cut -c1-330 gpurun_out/bench_n$N.json
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --workload train --batch 131072 --no-cpu-baseline > gpurun_out/bench_train_n$N.json 2> gpurun_out/bench_train_n$N.err
echo "train exit $?"; cut -c1-330 gpurun_out/bench_train_n$N.json
