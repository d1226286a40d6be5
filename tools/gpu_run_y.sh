#!/bin/bash
# GPU box: the GPU suite and three bench lines after the literal-zero terms left the root -> leaf pass -> gpurun_out/y_*
mkdir -p gpurun_out
(time timeout 100 python -m pytest tests -x -q -m gpu) > gpurun_out/y_pytest.log 2>&1; grep -E "passed|failed|error" gpurun_out/y_pytest.log | tail -2
: > gpurun_out/y_ab.txt
for w in so101_contact navbot_contact quadruped so101; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --sustain 0 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$w', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['clocks']['sm_mhz'])" | tee -a gpurun_out/y_ab.txt
done
