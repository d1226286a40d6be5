#!/bin/bash
# GPU box, round 2 run B: full GPU suite, the hardened bench on the headline and the new workloads, the
# per-step-control paths, the reference arm, ncu of the run-time-compiled twins.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q --durations=8) > gpurun_out/b_pytest.log 2>&1; tail -15 gpurun_out/b_pytest.log
python bench.py --steps 20 --warmup 3 --per-step-control > gpurun_out/b_bench_so101_contact.json 2> gpurun_out/b_bench.err; cat gpurun_out/b_bench_so101_contact.json; tail -5 gpurun_out/b_bench.err
for w in so101_contact_resting so101_pd so101_contact_pd acrobot_swingup; do
  python bench.py --workload $w --steps 20 --warmup 3 --sustain 0 > gpurun_out/b_bench_$w.json 2>> gpurun_out/b_bench.err; cut -c1-900 gpurun_out/b_bench_$w.json
done
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/b_reference.json 2>&1; cat gpurun_out/b_reference.json
BENCH_ARGS="--kernel jit_twin" tools/ncu_capture.sh so101_contact 262144 r2_jit_twin
BENCH_ARGS="--kernel jit_twin" tools/ncu_capture.sh navbot_contact 65536 r2_jit_twin
for w in so101_contact navbot_contact; do
  python tools/ncu_stall_map.py gpurun_out/r2_jit_twin_$w.ncu-rep 300 > gpurun_out/r2_jit_twin_${w}_stall_map.txt 2>&1
done
