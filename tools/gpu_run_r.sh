#!/bin/bash
# GPU box: the per-half named barrier of the warp-pair kernels (step-loop barrier no longer a block-wide __syncthreads
# reached from two instantiations): synccheck over the pair / trot / ticket tests, the GPU suite, 8 K-environment bench
# lines of the two sided workloads -> gpurun_out/r_*
mkdir -p gpurun_out
SEL='test_warp_pair_mapping_through_every_entry_point or test_quadruped_trot_controller_in_kernel or test_torque_sequence_is_the_per_step_control_closure or test_ticket_mode_replication_property'
(time timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "$SEL") > gpurun_out/sanitizer3_synccheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|real" gpurun_out/sanitizer3_synccheck.log | tail -4
(time python -m pytest tests -x -q -m gpu) > gpurun_out/r_pytest.log 2>&1; tail -3 gpurun_out/r_pytest.log
for w in navbot_contact quadruped; do
  python bench.py --workload $w --envs 8192 --steps 20 --warmup 5 --no-cpu-baseline --sustain 0 > gpurun_out/r_bench_${w}_8k.json 2> gpurun_out/r_bench_${w}_8k.err
  cut -c1-120 gpurun_out/r_bench_${w}_8k.json
done
