#!/bin/bash
# 8-GPU box: the N = 4 and N = 8 points of the 64 K strong-scaling sweep again (after the block-size rule of the warp pairs changed)
mkdir -p gpurun_out; OUT=gpurun_out/r2_scaling_n8.jsonl; : > $OUT
for w in navbot_contact quadruped; do for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $n --workload $w --envs-total 65536 --steps 20 --warmup 3 --no-cpu-baseline --sustain 0 2>> gpurun_out/r2_scaling_n8.err | grep '^{' >> $OUT
done; done
python - <<PY
import json
for l in open("$OUT"):
    d = json.loads(l)
    print(d["config"]["workload"], d["scaling"], "N=%d" % d["n_gpus"], "envs/GPU", d["config"]["n_envs_per_gpu"], "%.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "ms/launch %.3f" % d["ms_per_step"], d["config"]["mapping"])
PY
