#!/bin/bash
# GPU box: what the driver runs at round end, in the same order: the GPU suite, smoke(), the reference arm, the bench.
mkdir -p gpurun_out
(time python -m pytest tests -x -q -m gpu) > gpurun_out/final_pytest.log 2>&1; tail -6 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
(time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > gpurun_out/final_reference.json 2> gpurun_out/final_reference.time; cut -c1-260 gpurun_out/final_reference.json; tail -4 gpurun_out/final_reference.time
(time python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.time; cut -c1-400 gpurun_out/final_bench.json; tail -4 gpurun_out/final_bench.time
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sustain 0 > /dev/null 2>&1; grep -c step_kernel gpurun_out/final_launches.csv
