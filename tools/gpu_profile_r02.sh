#!/bin/bash
# Round-2 profiles: launch list of the default bench line (all legs) and one ncu --set full capture of each leg's
# dominant kernel.  Each ncu run follows a plain run of the same command in the same call (B200_PROFILING.md).
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_all.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc $?"
$B --no-secondary > gpurun_out/plain_sample.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:actor_sample_tc -s 8 -c 1 -f -o gpurun_out/prof_sampler_r02 $B --no-secondary > gpurun_out/ncu_sampler.log 2>&1
echo "sampler rc $?"
$B --workload ascent --batch 131072 > gpurun_out/plain_ascent.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:q_chain_tc -s 30 -c 1 -f -o gpurun_out/prof_qchain_r02 $B --workload ascent --batch 131072 > gpurun_out/ncu_qchain.log 2>&1
echo "qchain rc $?"
$B --workload train --batch 131072 > gpurun_out/plain_train.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:actor_train_chain -s 3 -c 1 -f -o gpurun_out/prof_trainchain_r02 $B --workload train --batch 131072 > gpurun_out/ncu_trainchain.log 2>&1
echo "train chain rc $?"
ncu --set full --clock-control none --import-source on -k regex:row_gemm -s 20 -c 3 -f -o gpurun_out/prof_rowgemm_r02 $B --workload train --batch 131072 > gpurun_out/ncu_rowgemm.log 2>&1
echo "row gemm rc $?"
$B --workload critic --batch 131072 > gpurun_out/plain_critic.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_critic_r02.csv $B --workload critic --batch 131072 > gpurun_out/ncu_critic.log 2>&1
echo "critic launch list rc $?"
ls -la gpurun_out/*.ncu-rep
