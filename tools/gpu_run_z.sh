#!/bin/bash
# GPU box, NOT YET RUN (written after the round's GPU budget was spent): everything the cuboid-built models
# (biped / leg / leg_from_foot, DESIGN.md section 7 last row) still owe on a B200 -> gpurun_out/z_*
#   gpurun --timeout 900 -- 'bash tools/gpu_run_z.sh'
mkdir -p gpurun_out
# 1. their parity tests alone, then the whole suite
(time timeout 600 python -m pytest tests/test_zz_widening_gpu.py -q) > gpurun_out/z_pytest_cuboid.log 2>&1
grep -E "passed|failed|error" gpurun_out/z_pytest_cuboid.log | tail -2
(time timeout 600 python -m pytest tests -x -q -m gpu) > gpurun_out/z_pytest.log 2>&1; grep -E "passed|failed|error" gpurun_out/z_pytest.log | tail -2
# 2. the bench line of record of the biped workload (64 K environments, thread per environment)
timeout 300 python bench.py --workload biped --steps 20 --warmup 3 > gpurun_out/z_bench_biped.json 2> gpurun_out/z_bench_biped.err
# 3. which mapping should a 13-body tree run at 64 K? its whole-tree kernel carries 1.7 KB of stack, the halves 168 B:
#    the break-even of the 9-body trees (19 K environments, gp_launch.h use_pairs) need not hold for it
: > gpurun_out/z_ab.txt
for n in 8192 16384 32768 65536; do for cfg in GP_STEP_PAIRS=0 GP_STEP_PAIRS=1; do
  env $cfg timeout 200 python bench.py --workload biped --envs $n --steps 20 --warmup 3 --no-cpu-baseline --sustain 0 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('biped', $n, '$cfg', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['config']['mapping'])" | tee -a gpurun_out/z_ab.txt
done; done
# 4. executed flop per env-step over the bench's launches (-> profiles/flop_counts.json "biped", then roofline.frac is no longer null)
W=biped bash tools/gpu_run_v.sh   # -> gpurun_out/v_biped_flops.json
