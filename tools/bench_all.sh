#!/bin/bash
# Developer tool: every workload once on the GPU box -> gpurun_out/bench_all.txt (+ .jsonl)
#   tools/bench_all.sh [tag] [workloads...]
TAG=${1:-run}; shift
W=${@:-so101_contact so101 navbot_contact quadruped hopper_1d rimless_wheel double_pendulum cart_pole}
mkdir -p gpurun_out
: > gpurun_out/bench_all_$TAG.txt
for w in $W; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tee -a gpurun_out/bench_all_$TAG.jsonl \
    | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$w', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'], d['clocks']['sm_mhz'])" >> gpurun_out/bench_all_$TAG.txt
done
cat gpurun_out/bench_all_$TAG.txt
