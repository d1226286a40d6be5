#!/bin/bash
# GPU box: the GPU suite, then bench lines of the main workloads after the integer sign flips of sincos_finish -> gpurun_out/u_*
mkdir -p gpurun_out
(time python -m pytest tests -x -q -m gpu) > gpurun_out/u_pytest.log 2>&1; tail -3 gpurun_out/u_pytest.log
: > gpurun_out/u_ab.txt
for w in so101_contact so101 navbot_contact so101_pd acrobot_swingup double_pendulum; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --sustain 0 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$w', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['clocks']['sm_mhz'])" | tee -a gpurun_out/u_ab.txt
done
