#!/bin/bash
# GPU box: what the driver runs at round end (tools/gpu_final_check.sh), the headline bench line of record, and the WHOLE
# GPU suite under compute-sanitizer memcheck -> gpurun_out/final_*, w_*
bash tools/gpu_final_check.sh
python bench.py --workload so101_contact --steps 20 --warmup 3 > gpurun_out/w_bench_so101_contact.json 2> gpurun_out/w_bench_so101_contact.err; cut -c1-200 gpurun_out/w_bench_so101_contact.json
(time timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -q -x) > gpurun_out/w_memcheck_suite.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|real" gpurun_out/w_memcheck_suite.log | tail -4
