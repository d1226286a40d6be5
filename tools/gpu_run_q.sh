#!/bin/bash
# GPU box: compute-sanitizer over the kernels added in round 2: warp pairs (shared-memory exchange behind named barriers),
# the trot controller (state read-modify-written at L2), torque sequences, ticket mode -> gpurun_out/sanitizer2_*.log
mkdir -p gpurun_out
SEL='test_warp_pair_mapping_through_every_entry_point or test_quadruped_trot_controller_in_kernel or test_torque_sequence_is_the_per_step_control_closure or test_ticket_mode_replication_property'
for tool in racecheck memcheck synccheck; do
  (time timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "$SEL") > gpurun_out/sanitizer2_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|real" gpurun_out/sanitizer2_$tool.log | tail -4
done
