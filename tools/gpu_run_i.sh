#!/bin/bash
# GPU box, round 2 run I: warp-pair kernels at 384 threads per block (168 registers, 12 warps per SM) against 256,
# conditioning of the mass matrices, the componentwise parity test.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -k "componentwise" 2>&1 | tail -8
: > gpurun_out/i_pairs384.txt
for w in navbot_contact quadruped; do for n in 8192 16384 32768 65536; do
 for cfg in "256 gorilla_physics_b200/lib/libgorilla_b200.so 1" "384 gorilla_physics_b200/lib/alt/lib_pairs384.so 1" "thread gorilla_physics_b200/lib/libgorilla_b200.so 0"; do
  set -- $cfg
  GP_LIB_PATH=$2 GP_STEP_PAIRS=$3 python bench.py --workload $w --envs $n --steps 20 --warmup 3 --no-cpu-baseline --sustain 0 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$w', $n, 'mapping=$1', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'])" >> gpurun_out/i_pairs384.txt
 done; done; done
cat gpurun_out/i_pairs384.txt
python tools/cond_mass_matrix.py > gpurun_out/r2_mass_matrix_conditioning.json 2> gpurun_out/i_cond.err; tail -12 gpurun_out/i_cond.err | cut -c1-400
