#!/bin/bash
# GPU box: FP64 work per environment time step over the bench's own trajectory, again, for the workloads whose kernels
# lost the sin/cos negations (every workload with revolute joints on the lockstep sin/cos) -> gpurun_out/v_<w>_flops.json
mkdir -p gpurun_out
for w in ${W:-so101_contact so101 so101_pd so101_contact_pd so101_contact_resting navbot_contact double_pendulum cart_pole acrobot_swingup}; do
  N=$(python -c "from gorilla_physics_b200 import WORKLOADS; print(WORKLOADS['$w'].n_envs)")
  SKIP=$(python -c "from gorilla_physics_b200 import WORKLOADS; print(3 + (1 if WORKLOADS['$w'].settle_steps else 0))")
  ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \
      --clock-control none -k regex:step_kernel --csv --log-file gpurun_out/v_${w}_flops.csv \
      python bench.py --workload $w --steps 40 --warmup 3 --no-cpu-baseline --sustain 0 > /dev/null 2>&1
  python tools/ncu_flops_over_bench.py gpurun_out/v_${w}_flops.csv $N 128 40 $SKIP > gpurun_out/v_${w}_flops.json
  rm -f gpurun_out/v_${w}_flops.csv
  echo $w $(cat gpurun_out/v_${w}_flops.json | head -c 300)
done
