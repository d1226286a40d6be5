#!/bin/bash
# GPU box: refresh the quadruped's profile of record (its kernels now carry the trot controller) and add bench lines of the
# warp-pair regime (8 192 environments)
mkdir -p gpurun_out
W="quadruped" tools/refresh_profiles.sh r2 > gpurun_out/p_refresh.log 2>&1
cp gpurun_out/profile_flop_counts.json profiles/flop_counts.json
python bench.py --workload quadruped --steps 20 --warmup 3 > gpurun_out/r2_bench_quadruped.json 2>> gpurun_out/p_bench.err
for w in navbot_contact quadruped; do python bench.py --workload $w --envs 8192 --steps 20 --warmup 3 > gpurun_out/r2_bench_${w}_8k.json 2>> gpurun_out/p_bench.err; done
for f in gpurun_out/r2_bench_quadruped.json gpurun_out/r2_bench_navbot_contact_8k.json gpurun_out/r2_bench_quadruped_8k.json; do python -c "import sys,json; d=json.loads(open('$f').readline()); print(d['config']['workload'], d['config']['n_envs_per_gpu'], d['config']['mapping'], '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['parity_sample']['one_step_rel_err_max'])"; done
