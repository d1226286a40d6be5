#!/usr/bin/env python
"""FP64 work per environment time step AVERAGED OVER THE BENCH'S OWN TRAJECTORY.

A contact workload executes more instructions once its bodies lie on the ground than while they fall,
so the count of one early launch understates (or overstates) what the timed region runs. This reads the
CSV log of
   ncu --metrics smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on.sum -k regex:step_kernel
       --csv --log-file X.csv python bench.py --workload W --steps K --warmup 3 --no-cpu-baseline
(three counters: one replay pass per launch) and divides the totals over the TIMED launches by
n_envs * inner * K. bench.py launches W warm-up steps, then the K timed ones, then the chunks of its
end-to-end leg: the timed launches are step-kernel launches W .. W+K-1 of the process.
usage: python tools/ncu_flops_over_bench.py X.csv n_envs inner [K=40] [W=3]"""
import csv
import json
import sys

path, n_envs, inner = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
K = int(sys.argv[4]) if len(sys.argv) > 4 else 40
W = int(sys.argv[5]) if len(sys.argv) > 5 else 3
rows = []
with open(path, newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
per_launch = {}
for r in rd:
    key = int(r["ID"])
    d = per_launch.setdefault(key, {"grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
    name = r["Metric Name"]
    val = float(r["Metric Value"].replace(",", ""))
    for op in ("dfma", "dadd", "dmul"):
        if f"op_{op}_pred_on" in name:
            d[op] = val
launches = [per_launch[k] for k in sorted(per_launch)]
timed = launches[W:W + K]
assert len(timed) == K, (len(launches), K, W)
tot = {op: sum(d.get(op, 0.0) for d in timed) for op in ("dfma", "dadd", "dmul")}
env_steps = n_envs * inner * len(timed)
per_launch_flop = [(2 * d.get("dfma", 0) + d.get("dadd", 0) + d.get("dmul", 0)) / (n_envs * inner) for d in timed]
print(json.dumps({
    "launches_counted": len(timed), "env_steps": env_steps,
    "flop_per_env_step": (2 * tot["dfma"] + tot["dadd"] + tot["dmul"]) / env_steps,
    "fp64_inst_per_env_step": (tot["dfma"] + tot["dadd"] + tot["dmul"]) / env_steps,
    "flop_per_env_step_first_launch": per_launch_flop[0], "flop_per_env_step_last_launch": per_launch_flop[-1],
}))
