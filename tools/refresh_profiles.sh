#!/bin/bash
# GPU box: one `ncu --set full` capture of the step kernel per bench workload, digested into
# profiles/<round>_<workload>_step_kernel.json, plus the launch list of the default bench command and
# profiles/flop_counts.json (what bench.py's roofline reads).   [W="workload ..."] tools/refresh_profiles.sh r2
R=${1:-r2}
mkdir -p gpurun_out profiles
declare -A N
ALL=$(python -c "from gorilla_physics_b200 import WORKLOADS; print(' '.join(WORKLOADS))")
for w in $ALL; do N[$w]=$(python -c "from gorilla_physics_b200 import WORKLOADS; print(WORKLOADS['$w'].n_envs)"); done
for w in ${W:-$ALL}; do
  tools/ncu_capture.sh $w ${N[$w]} $R > /dev/null 2>&1
  cp gpurun_out/${R}_$w.json gpurun_out/profile_${R}_${w}_step_kernel.json
  python tools/ncu_stall_map.py gpurun_out/${R}_$w.ncu-rep 300 > gpurun_out/profile_${R}_${w}_stall_map.txt 2>&1
  rm -f gpurun_out/${R}_$w.ncu-rep
  # FP64 work averaged over the launches the bench times (tools/ncu_flops_over_bench.py)
  ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \
      --clock-control none -k regex:step_kernel --csv --log-file gpurun_out/${R}_${w}_flops.csv \
      python bench.py --workload $w --steps 40 --warmup 3 --no-cpu-baseline --sustain 0 > /dev/null 2>&1
  # (step-kernel launches before the timed ones: 3 warm-up launches, plus the settling rollout of the "resting" workloads)
  SKIP=$(python -c "from gorilla_physics_b200 import WORKLOADS; print(3 + (1 if WORKLOADS['$w'].settle_steps else 0))")
  python tools/ncu_flops_over_bench.py gpurun_out/${R}_${w}_flops.csv ${N[$w]} 128 40 $SKIP > gpurun_out/profile_${R}_${w}_flops_over_bench.json
  rm -f gpurun_out/${R}_${w}_flops.csv
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/profile_${R}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sustain 0 > gpurun_out/profile_${R}_launches.log 2>&1
python - <<PY
import json, glob
out = {"_how": "ncu --set full + smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on.sum on one step-kernel launch of \`python bench.py --workload W --inner 64 --steps 3 --warmup 3\` (tools/refresh_profiles.sh on the GPU box, B200); flop = 2*dfma + dadd + dmul; per env-step = / (n_envs * 64) -> flop_per_env_step_one_launch. flop_per_env_step (what bench.py uses) = the same three counters summed over the 40 timed launches of \`python bench.py --workload W\` (defaults: 40 x 128 steps after 3 warm-up launches) / (n_envs * 128 * 40), i.e. averaged over the trajectory the bench times (tools/ncu_flops_over_bench.py). Per-workload summaries: profiles/${R}_<workload>_step_kernel.json (tools/ncu_summary.py)",
       "flop_per_env_step": {}, "flop_per_env_step_one_launch": {}, "fp64_inst_per_env_step": {}, "dram_bytes_per_launch": {}, "fp64_pipe_pct_active": {}}
try:  # workloads not captured in this run keep their previous numbers
    prev = json.load(open("profiles/flop_counts.json"))
    for k in out:
        if k != "_how" and isinstance(prev.get(k), dict):
            out[k].update(prev[k])
except Exception:
    pass
for f in sorted(glob.glob("gpurun_out/profile_${R}_*_step_kernel.json")):
    w = f.split("profile_${R}_")[1].replace("_step_kernel.json", "")
    d = json.load(open(f))[0]
    out["flop_per_env_step_one_launch"][w] = round(d["flop_per_env_step"], 1)
    try:
        avg = json.load(open(f.replace("_step_kernel.json", "_flops_over_bench.json")))
        out["flop_per_env_step"][w] = round(avg["flop_per_env_step"], 1)
        out["fp64_inst_per_env_step"][w] = round(avg["fp64_inst_per_env_step"], 1)
    except Exception:
        out["flop_per_env_step"][w] = round(d["flop_per_env_step"], 1)
        out["fp64_inst_per_env_step"][w] = round(d["fp64_inst_per_env_step"], 1)
    out["dram_bytes_per_launch"][w] = int(d["dram_bytes_read"] + d["dram_bytes_write"])
    out["fp64_pipe_pct_active"][w] = round(d["fp64_pipe_pct_of_peak_active"], 1)
json.dump(out, open("gpurun_out/profile_flop_counts.json", "w"), indent=1)
print(json.dumps(out, indent=1))
PY
