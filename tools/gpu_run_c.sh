#!/bin/bash
# GPU box, round 2 run C: the new tests, then compute-sanitizer (memcheck, racecheck, synccheck) over the ticket-mode,
# pipelined-simulate and torque-sequence tests -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -k "cpp or c_host or sharded or config1 or diagnostic" -s) > gpurun_out/c_pytest.log 2>&1; tail -12 gpurun_out/c_pytest.log
SEL='test_ticket_mode_replication_property or test_pipelined_simulate_matches_resident_stepping or test_torque_sequence_is_the_per_step_control_closure or test_ragged_batch_sizes_and_simulate_host_path'
for tool in memcheck racecheck synccheck; do
  (time timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "$SEL") > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|real" gpurun_out/sanitizer_$tool.log | tail -4
done
