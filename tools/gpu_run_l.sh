#!/bin/bash
# GPU box, round 2 run L: warp-pair coverage test, and the block barrier cadence (every 4th / 2nd / every step) at small and
# full batches of the 14-dof trees
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "warp_pair_mapping" 2>&1 | tail -5
: > gpurun_out/l_sync_every.txt
for w in navbot_contact quadruped; do for n in 8192 16384 65536; do
 for cfg in "4 gorilla_physics_b200/lib/libgorilla_b200.so" "2 gorilla_physics_b200/lib/alt/lib_sync2.so" "1 gorilla_physics_b200/lib/alt/lib_sync1.so"; do
  set -- $cfg
  GP_LIB_PATH=$2 python bench.py --workload $w --envs $n --steps 20 --warmup 3 --no-cpu-baseline --sustain 0 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$w', $n, 'barrier_every=$1', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'])" >> gpurun_out/l_sync_every.txt
 done; done; done
cat gpurun_out/l_sync_every.txt
