#!/bin/bash
# Developer tool (GPU box): A/B of run-time switches of ONE build (GP_NO_TICKETS=1, GP_PIPE_CHUNKS=n,
# GP_STEP_BLOCK=n ...), two repetitions per workload -> gpurun_out/ab_env.txt
#   tools/ab_env.sh "<env assignment A>|<env assignment B>" workloads...
IFS='|' read -ra CFG <<< "$1"; shift
: > gpurun_out/ab_env.txt
for w in "$@"; do for rep in 1 2; do for cfg in "${CFG[@]}"; do
  env $cfg timeout 200 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --sustain 0 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$w', '$cfg', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'])" >> gpurun_out/ab_env.txt
done; done; done
cat gpurun_out/ab_env.txt
