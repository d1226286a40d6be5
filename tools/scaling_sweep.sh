#!/bin/bash
# GPU box with 8 GPUs (gpurun --gpus 8): BASELINE config 5 (navbot, 64 K environments at 1/2/4/8 GPUs) as strong
# scaling (64 K environments in total, split over the ranks) and weak scaling (64 K per GPU), the quadruped twin, the
# two-rank NCCL diagnostic test and the plain-C sharded test on distinct devices.
#   tools/scaling_sweep.sh [tag]  -> gpurun_out/<tag>_scaling.jsonl
TAG=${1:-r2}
OUT=gpurun_out/${TAG}_scaling.jsonl
mkdir -p gpurun_out; : > $OUT
run() {  # n_gpus, bench args...
  local n=$1; shift
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 "$@" --no-cpu-baseline --sustain 0 2>> gpurun_out/${TAG}_scaling.err >> $OUT
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $n "$@" --no-cpu-baseline --sustain 0 2>> gpurun_out/${TAG}_scaling.err >> $OUT
  fi
}
nvidia-smi -L | head -8
for w in navbot_contact quadruped; do
  for n in 1 2 4 8; do run $n --workload $w --envs-total 65536 --steps 20 --warmup 3; done   # strong
done
for n in 2 4 8; do run $n --workload navbot_contact --steps 20 --warmup 3; done               # weak (N = 1 is the first strong line)
run 8 --workload so101_contact --steps 20 --warmup 3
python - <<PY
import json
for l in open("$OUT"):
    if not l.startswith("{"):
        continue  # (NCCL prints its version on stdout)
    d = json.loads(l)
    print(d["config"]["workload"], d["scaling"], "N=%d" % d["n_gpus"], "envs/GPU", d["config"]["n_envs_per_gpu"], "%.4g" % d["value"],
          "e2e %.4g" % d["e2e"]["value"], "ms/launch %.3f" % d["ms_per_step"], d["config"]["kernel"])
PY
(time python -m pytest tests -m gpu -q -k "diagnostic or c_host or sharded") > gpurun_out/${TAG}_multigpu_tests.log 2>&1; tail -5 gpurun_out/${TAG}_multigpu_tests.log
