#!/bin/bash
# GPU box, round 2 run F: full suite with a warm kernel cache, cost of the per-step torque branch, ncu of the warp-pair kernels
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q --durations=10) > gpurun_out/f_pytest_all.log 2>&1; tail -22 gpurun_out/f_pytest_all.log
L="gorilla_physics_b200/lib/libgorilla_b200.so gorilla_physics_b200/lib/alt/lib_notau.so"
AB_ARGS="--steps 40 --warmup 3" tools/ab_bench.sh "$L" navbot_contact quadruped so101_contact; cp gpurun_out/ab.txt gpurun_out/f_ab_notau.txt
BENCH_ARGS="--envs 8192" tools/ncu_capture.sh navbot_contact 8192 r2_pairs8k
BENCH_ARGS="--envs 8192" GP_STEP_PAIRS=0 tools/ncu_capture.sh quadruped 8192 r2_whole8k
BENCH_ARGS="--envs 8192" tools/ncu_capture.sh quadruped 8192 r2_pairs8k
for w in navbot_contact quadruped; do python tools/ncu_stall_map.py gpurun_out/r2_pairs8k_$w.ncu-rep 300 > gpurun_out/r2_pairs8k_${w}_stall_map.txt 2>&1; done
rm -f gpurun_out/*.ncu-rep
