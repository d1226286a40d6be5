#!/bin/bash
# GPU box, round 2 run E: both mappings compiled, picked per launch. Parity, the refactor's cost on the thread-per-
# environment kernels (against the pre-refactor build), and where warp pairs stop paying (batch-size sweep).
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q --durations=25) > gpurun_out/e_pytest_all.log 2>&1; tail -40 gpurun_out/e_pytest_all.log
L="gorilla_physics_b200/lib/libgorilla_b200.so gorilla_physics_b200/lib/alt/lib_prerefactor.so"
AB_ARGS="--steps 40 --warmup 3" tools/ab_bench.sh "$L" navbot_contact quadruped so101_contact double_pendulum; cp gpurun_out/ab.txt gpurun_out/e_ab_refactor.txt
: > gpurun_out/e_pairs_sweep.txt
for w in navbot_contact quadruped; do for n in 4096 8192 16384 24576 32768 49152; do for pairs in 0 1; do
  GP_STEP_PAIRS=$pairs python bench.py --workload $w --envs $n --steps 20 --warmup 3 --no-cpu-baseline --sustain 0 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$w', $n, 'pairs=$pairs', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'])" >> gpurun_out/e_pairs_sweep.txt
done; done; done
cat gpurun_out/e_pairs_sweep.txt
