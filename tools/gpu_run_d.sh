#!/bin/bash
# GPU box, round 2 run D: the warp-pair kernels (navbot, quadruped and the run-time-compiled twins): parity, then A/B
# against the thread-per-environment build (lib/alt/lib_nosides.so, make EXTRA=-DGP_NO_SIDES) at 64 K and 8 K environments.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x -k "navbot or quadruped or ticket or golden or featherstone") > gpurun_out/d_pytest_pairs.log 2>&1; tail -8 gpurun_out/d_pytest_pairs.log
if ! grep -q " passed" gpurun_out/d_pytest_pairs.log || grep -q " failed" gpurun_out/d_pytest_pairs.log; then echo "PARITY FAILED: no bench"; exit 1; fi
L="gorilla_physics_b200/lib/libgorilla_b200.so gorilla_physics_b200/lib/alt/lib_nosides.so"
AB_ARGS="--steps 20 --warmup 3" tools/ab_bench.sh "$L" navbot_contact quadruped so101_contact; cp gpurun_out/ab.txt gpurun_out/d_ab_full.txt
AB_ARGS="--steps 20 --warmup 3 --envs 8192" tools/ab_bench.sh "$L" navbot_contact quadruped; cp gpurun_out/ab.txt gpurun_out/d_ab_8k.txt
(time python -m pytest tests -m gpu -q) > gpurun_out/d_pytest_all.log 2>&1; tail -8 gpurun_out/d_pytest_all.log
