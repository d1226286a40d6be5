#!/bin/bash
# GPU box: compute-sanitizer (racecheck, memcheck, synccheck) over the round-2 kernels as they are now - warp pairs behind
# the out-of-line named barrier, per-warp ticket items, trot controller, torque sequences -> gpurun_out/sanitizer4_*.log;
# then the rimless wheel's profile of record and the pair kernels' 8 K captures again.
mkdir -p gpurun_out
SEL='test_warp_pair_mapping_through_every_entry_point or test_quadruped_trot_controller_in_kernel or test_torque_sequence_is_the_per_step_control_closure or test_ticket_mode_replication_property'
for tool in synccheck memcheck racecheck; do
  (time timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "$SEL") > gpurun_out/sanitizer4_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|real" gpurun_out/sanitizer4_$tool.log | tail -4
done
W=rimless_wheel tools/refresh_profiles.sh r2 > gpurun_out/t_refresh.log 2>&1; tail -3 gpurun_out/t_refresh.log
for w in navbot_contact quadruped; do
  BENCH_ARGS="--envs 8192" tools/ncu_capture.sh $w 8192 r2_pairs8k > /dev/null 2>&1
  python tools/ncu_stall_map.py gpurun_out/r2_pairs8k_$w.ncu-rep 300 > gpurun_out/r2_pairs8k_${w}_stall_map.txt 2>&1
  rm -f gpurun_out/r2_pairs8k_$w.ncu-rep
  head -c 600 gpurun_out/r2_pairs8k_$w.json
done
python bench.py --workload rimless_wheel --steps 20 --warmup 3 > gpurun_out/t_bench_rimless_wheel.json 2> gpurun_out/t_bench_rimless_wheel.err; cut -c1-160 gpurun_out/t_bench_rimless_wheel.json
