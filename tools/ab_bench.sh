#!/bin/bash
# Developer tool: compare builds of libgorilla_b200.so on the GPU box (tuning builds: see csrc/Makefile).
#   tools/ab_bench.sh "<libA.so> <libB.so> ..." [workloads...]   -> gpurun_out/ab.txt
#   AB_ARGS="--steps 20 --warmup 3" overrides the bench arguments (default: 64 fused steps x 13 launches,
#   i.e. the first ~800 time steps; contact workloads that start in the air need the longer default run)
LIBS=$1; shift
W=${@:-so101_contact so101 navbot_contact quadruped hopper_1d rimless_wheel double_pendulum cart_pole}
ARGS=${AB_ARGS:---steps 10 --warmup 3 --inner 64}
mkdir -p gpurun_out
: > gpurun_out/ab.txt
for w in $W; do
  for rep in 1 2; do
    for lib in $LIBS; do
      GP_LIB_PATH=$lib python bench.py --workload $w $ARGS --no-cpu-baseline --sustain 0 2>/dev/null \
        | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$w', '$lib', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['clocks']['sm_mhz'])" >> gpurun_out/ab.txt
    done
  done
done
cat gpurun_out/ab.txt
