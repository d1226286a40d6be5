#!/bin/bash
# GPU box, round 2 run H: the profiles of record. Per workload: ncu --set full of the step kernel + FP64 work averaged over
# the bench's timed launches (tools/refresh_profiles.sh), then - with those counts in place - the bench line of every
# workload (CPU baseline and parity sample included) and the reference arm.
mkdir -p gpurun_out
tools/refresh_profiles.sh r2 > gpurun_out/h_refresh.log 2>&1
cp gpurun_out/profile_flop_counts.json profiles/flop_counts.json
for w in $(python -c "from gorilla_physics_b200 import WORKLOADS; print(' '.join(WORKLOADS))"); do
  python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/r2_bench_$w.json 2>> gpurun_out/h_bench.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r2_bench_$w.json').readline()); r=d['roofline']; print('$w', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'frac', r['frac'] and round(r['frac'],3), 'cpu %.3g' % d['cpu_baseline']['value'], d['clocks']['sm_mhz'])"
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/h_bench.err; cut -c1-300 gpurun_out/r2_bench_reference.json
