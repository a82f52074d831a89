#!/bin/bash
# Round-end regression + measurements in one gpurun call: full GPU suite, smoke, every bench workload, launch lists
# and one ncu --set full capture of the two fused kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cut -c1-300 gpurun_out/bench_ref.json
for wl in ascent train; do for pr in bf16 fp32; do
  timeout 300 python bench.py --workload $wl --precision $pr --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}_${pr}.json 2> gpurun_out/bench_${wl}_${pr}.err
  python -c "import json; d=json.load(open('gpurun_out/bench_${wl}_${pr}.json')); print('$wl $pr', round(d['ms_per_step'],3), 'ms', round(d['value']/1e6,2), 'M/s frac', round(d['roofline']['frac'],4))"
done; done
timeout 300 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sample_fp32.json 2>/dev/null
timeout 300 python bench.py --batch 256 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_sample_b256.json 2>/dev/null
if [ -n "$NCU" ]; then
  for wl in sample ascent; do
    python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$wl.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$wl.csv \
        python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$wl.log 2>&1
  done
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:actor_sample_tc -s 3 -c 1 -f -o gpurun_out/prof_sampler_tc \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_sampler.log 2>&1
  python bench.py --workload ascent --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain3.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:q_chain_tc -s 8 -c 1 -f -o gpurun_out/prof_qchain \
      python bench.py --workload ascent --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_qchain.log 2>&1
  ls -la gpurun_out/*.ncu-rep
fi
