#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --workload train --batch 131072 > gpurun_out/bench_train_n$N.json 2> gpurun_out/bench_train_n$N.err
echo "exit $?"; cat gpurun_out/bench_train_n$N.json; tail -5 gpurun_out/bench_train_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --workload ascent > gpurun_out/bench_ascent_n$N.json 2> gpurun_out/bench_ascent_n$N.err
echo "exit $?"; cut -c1-300 gpurun_out/bench_ascent_n$N.json; tail -3 gpurun_out/bench_ascent_n$N.err
