#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) of a step kernel into the numbers the roofline uses.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep n_envs inner_steps > profiles/xyz.json"""
import csv
import io
import json
import subprocess
import sys

rep, n_envs, inner = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for vals in rows[2:]:
    m = {h: (v, u) for h, v, u in zip(hdr, vals, units)}

    def f(name):
        v, u = m[name]
        x = float(v.replace(",", ""))
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3,
                 "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
        return x * scale

    env_steps = n_envs * inner
    dfma = f("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum")
    dadd = f("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum")
    dmul = f("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum")
    dur = f("gpu__time_duration.sum")
    flops = 2 * dfma + dadd + dmul
    out.append({
        "kernel": m["Kernel Name"][0],
        "duration_ms": dur * 1e3,
        "env_steps_per_launch": env_steps,
        "env_steps_per_s_under_ncu": env_steps / dur,
        "registers_per_thread": int(float(m["launch__registers_per_thread"][0])),
        "grid": m["launch__grid_size"][0], "block": m["launch__block_size"][0],
        "sm_clock_ghz": f("sm__cycles_elapsed.avg.per_second") if "Ghz" not in m["sm__cycles_elapsed.avg.per_second"][1] else float(m["sm__cycles_elapsed.avg.per_second"][0]),
        "fp64_pipe_pct_of_peak_active": float(m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0]),
        "fp64_pipe_pct_of_peak_elapsed": float(m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"][0]),
        "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
        "warps_active_pct": float(m["sm__warps_active.avg.pct_of_peak_sustained_active"][0]),
        "thread_inst_dfma": dfma, "thread_inst_dadd": dadd, "thread_inst_dmul": dmul,
        "fp64_inst_per_env_step": (dfma + dadd + dmul) / env_steps,
        "flop_per_env_step": flops / env_steps,
        "achieved_tflops_under_ncu": flops / dur / 1e12,
        "dram_bytes_read": f("dram__bytes_read.sum"), "dram_bytes_write": f("dram__bytes_write.sum"),
        "dram_bytes_per_env_step": (f("dram__bytes_read.sum") + f("dram__bytes_write.sum")) / env_steps,
        "local_ld_sectors": f("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"),
        "local_st_sectors": f("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"),
    })
print(json.dumps(out, indent=1))
