#!/usr/bin/env python
"""Developer tool: where do the FP64 instructions of a step kernel come from?

Attributes every DFMA / DMUL / DADD / DSETP / MUFU of one kernel of a built variant to the statement of
dynamics_core (gp_dynamics.cuh) - or of gp_kernels.cuh outside it - that it was inlined from, through the inlining
chains of `nvdisasm -gi` (the library is built with -lineinfo). The step body is fully unrolled, so the static count
of a statement is what one time step executes of it (branches of the contact law counted once per copy).

    python tools/sass_attribution.py so101 [--contact 1] [--pairs] [--lo 360 --hi 1000] [--zero]

--zero lists the FP64 instructions that have RZ as an operand: additions of zero and NEGATIONS (`DADD R, -RZ, -R`:
a negation that feeds a select or a compare cannot become an operand modifier and costs an FP64 issue slot; this is
how the twelve negations of the sin/cos quadrant fix-up were found, profiles/r2_tuning.md).
"""
import argparse
import collections
import re
import subprocess
import tempfile
from pathlib import Path

OBJ = Path(__file__).resolve().parent.parent / "gorilla_physics_b200" / "lib" / "obj" / "variants"
FILE = re.compile(r'File ".*/(\w+\.\w+)", line (\d+)')
INS = re.compile(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+\s+)?(([A-Z0-9_]+)[ .].*)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variant")
    ap.add_argument("--contact", type=int, default=1)
    ap.add_argument("--pairs", action="store_true")
    ap.add_argument("--lo", type=int, default=360, help="first line of dynamics_core's body in gp_dynamics.cuh")
    ap.add_argument("--hi", type=int, default=1000)
    ap.add_argument("--zero", action="store_true")
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", str(OBJ / f"variant_{a.variant}.o")], cwd=tmp, check=True, capture_output=True)
        cubin = next(Path(tmp).glob("*.cubin"))
        sass = subprocess.run(["nvdisasm", "-gi", "-c", str(cubin)], capture_output=True, text=True).stdout.splitlines()
    tag = f"ELi{a.contact}ELi0ELb{1 if a.pairs else 0}ELb0E"  # <Topo, CONTACT, IntegSIE, PAIRS, no torque sequence>
    inside, chain, pend = False, [], []
    count, kinds, zero = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
    for line in sass:
        if line.startswith(".text."):
            inside = "step_kernel" in line and tag in line
            continue
        if not inside:
            continue
        m = FILE.search(line)
        if m:
            pend.append((m.group(1), int(m.group(2))))
            continue
        m = INS.match(line)
        if not m:
            continue
        if pend:
            chain, pend = pend, []
        text, op = m.group(1), m.group(2)
        if op not in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"):
            continue
        # innermost frame inside dynamics_core's body, else the innermost frame in gp_kernels.cuh
        key = next((f for f in chain if f[0] == "gp_dynamics.cuh" and a.lo <= f[1] <= a.hi), None)
        if key is None:
            key = next((f for f in chain if f[0] == "gp_kernels.cuh"), chain[0] if chain else ("?", 0))
        count[key] += 1
        kinds[key][op] += 1
        if a.zero and op != "MUFU" and "RZ" in text:
            zero[(key, chain[0] if chain else key, re.sub(r"R\d+", "R", text.split(";")[0]))] += 1
    print(f"variant_{a.variant}.o, step kernel {tag}: {sum(count.values())} FP64-pipe instructions")
    for key, n in sorted(count.items()):
        print(f"  {key[0]}:{key[1]:<5d} {n:5d}  {dict(kinds[key])}")
    if a.zero:
        print("with RZ as an operand:")
        for (key, inner, text), n in sorted(zero.items(), key=lambda kv: str(kv[0])):
            print(f"  {key[0]}:{key[1]} <- {inner[0]}:{inner[1]}  {text}  x{n}")


if __name__ == "__main__":
    main()
