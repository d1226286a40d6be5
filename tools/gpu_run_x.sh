#!/bin/bash
# GPU box: forced step-kernel block sizes at full batch sizes (GP_STEP_BLOCK, run-time switch) -> gpurun_out/x_blocks.txt
mkdir -p gpurun_out; : > gpurun_out/x_blocks.txt
run() { # workload block
  if [ "$2" = default ]; then unset GP_STEP_BLOCK; else export GP_STEP_BLOCK=$2; fi
  python bench.py --workload $1 --steps 20 --warmup 3 --no-cpu-baseline --sustain 0 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1 block=$2', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['clocks']['sm_mhz'])" | tee -a gpurun_out/x_blocks.txt
}
for b in default 128 64; do run so101_contact $b; done
for b in default 128; do run navbot_contact $b; run quadruped $b; done
for b in default 64 256; do run rimless_wheel $b; done
for b in default 64; do run hopper_1d $b; done
