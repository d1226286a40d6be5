#!/bin/bash
# GPU box: per-warp ticket items (kernels with blocks below 256 threads): ticket tests, rimless wheel bench with and
# without ticket mode, hopper / generic twins as controls -> gpurun_out/s_*
mkdir -p gpurun_out
(time python -m pytest tests/test_parity_gpu.py -q -m gpu -k "ticket_mode or ragged or pipelined") > gpurun_out/s_pytest.log 2>&1; tail -3 gpurun_out/s_pytest.log
for rep in 1 2; do
  for mode in tickets notickets; do
    if [ $mode = notickets ]; then export GP_NO_TICKETS=1; else unset GP_NO_TICKETS; fi
    python bench.py --workload rimless_wheel --steps 20 --warmup 5 --no-cpu-baseline --sustain 0 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('rimless_wheel $mode', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['clocks']['sm_mhz'])" | tee -a gpurun_out/s_ab.txt
  done
done
unset GP_NO_TICKETS
python bench.py --workload so101_contact --kernel generic --envs 65536 --steps 10 --warmup 3 --no-cpu-baseline --sustain 0 2>/dev/null | cut -c1-100 | tee -a gpurun_out/s_ab.txt
