#!/bin/bash
# Developer tool (GPU box): one `ncu --set full` capture of the step kernel of a workload, digested.
#   [BENCH_ARGS="--kernel jit_twin"] tools/ncu_capture.sh <workload> <n_envs> [tag]   -> gpurun_out/<tag>_<workload>.{ncu-rep,json,stalls.txt}
W=$1; N=$2; TAG=${3:-r1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:step_kernel -c 1 -s ${NCU_SKIP:-3} \
    --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \
    -o gpurun_out/${TAG}_$W -f python bench.py --workload $W --inner 64 --steps ${NCU_STEPS:-3} --warmup 3 --no-cpu-baseline --sustain 0 ${BENCH_ARGS:-} > gpurun_out/${TAG}_$W.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_$W.ncu-rep $N 64 > gpurun_out/${TAG}_$W.json
ncu -i gpurun_out/${TAG}_$W.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
d=dict(zip(h,v))
for k in sorted(d):
    if 'issue_stalled' in k and k.endswith('_per_warp_active.pct') or 'inst_executed_pipe' in k and k.endswith('.sum') or k in ('smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):
        print(k, d[k])
" > gpurun_out/${TAG}_$W.stalls.txt
cat gpurun_out/${TAG}_$W.json | head -30
