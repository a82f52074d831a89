#!/bin/bash
# Re-capture of what the last round-2 changes touched (critic chain: L2 policy of the ELU' scratch, forward-only image;
# sampler: scratch accessors): launch list of the default bench line, one ncu --set full capture of the sampler and of the
# critic chain, then the default bench record.  Each ncu run follows a plain run of the same command in the same call.
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
python bench.py > gpurun_out/bench_default_n1_final.json 2> gpurun_out/bench_default_n1_final.err
echo "bench rc $?"
ls -la gpurun_out/*.ncu-rep
