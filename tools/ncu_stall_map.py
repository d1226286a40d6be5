#!/usr/bin/env python
"""Developer tool: where a step kernel's warps wait. Reads `ncu -i X.ncu-rep --page source --csv`
output and prints, per stretch of N SASS instructions, executed FP64 / other instructions and the
stall samples by reason.   python tools/ncu_stall_map.py rep.ncu-rep [chunk=100]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 100
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
reasons = ["stall_wait", "stall_math", "stall_not_selected", "stall_selected", "stall_short_sb", "stall_no_inst",
           "stall_barrier", "stall_branch_resolving", "stall_dispatch", "stall_long_sb"]
tot = sum(int(r["# Samples"]) for r in rows)
print(f"{len(rows)} instructions, {tot} samples")
print("range        exec_fp64 exec_other samples%  " + " ".join(x.replace("stall_", "")[:8].rjust(8) for x in reasons))
for c in range(0, len(rows), chunk):
    part = rows[c:c + chunk]
    f64 = sum(int(r["Instructions Executed"]) for r in part if r["Source"].strip().split()[0].lstrip("@!P0123456789U ").startswith(("DFMA", "DMUL", "DADD", "DSETP")) or any(r["Source"].strip().startswith(p) for p in ()))
    f64 = 0; oth = 0
    for r in part:
        toks = r["Source"].strip().split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        n = int(r["Instructions Executed"])
        if op.startswith(("DFMA", "DMUL", "DADD", "DSETP")): f64 += n
        else: oth += n
    s = sum(int(r["# Samples"]) for r in part)
    print(f"{c:5d}-{c+len(part):5d} {f64/8192:9.1f} {oth/8192:9.1f} {100*s/tot:7.2f}%  " + " ".join(f"{sum(int(r[x]) for r in part):8d}" for x in reasons))
