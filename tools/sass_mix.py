#!/usr/bin/env python
"""Developer tool: static FP64 instruction mix (DFMA / DMUL / DADD, other) and code size of every
step kernel in the built objects.   python tools/sass_mix.py [variant ...]"""
import re
import subprocess
import sys
from pathlib import Path

OBJ = Path(__file__).resolve().parent.parent / "gorilla_physics_b200" / "lib" / "obj" / "variants"
pat = re.compile(r"^\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)")

def main():
    names = sys.argv[1:] or sorted(p.stem.replace("variant_", "") for p in OBJ.glob("variant_*.o"))
    for n in names:
        out = subprocess.run(["cuobjdump", "-sass", str(OBJ / f"variant_{n}.o")], capture_output=True, text=True).stdout
        fn, mix, last = None, {}, 0
        def flush():
            if fn and "step_kernel" in fn:
                f = mix.get("DFMA", 0); m = mix.get("DMUL", 0); a = mix.get("DADD", 0)
                tot = sum(mix.values())
                short = re.sub(r".*step_kernelINS_(\w+?)EEE?ELi(\d)ELi(\d)EEE.*", r"\1 C\2 I\3", fn)
                print(f"{n:16s} {short[-40:]:40s} DFMA {f:5d} DMUL {m:5d} DADD {a:5d} fp64 {f+m+a:5d} all {tot:5d} bytes {last+16}")
        for line in out.splitlines():
            if "Function :" in line:
                flush()
                fn, mix, last = line.split(":")[1].strip(), {}, 0
                continue
            mm = pat.match(line)
            if mm:
                last = int(mm.group(1), 16)
                mix[mm.group(2)] = mix.get(mm.group(2), 0) + 1
        flush()

if __name__ == "__main__":
    main()
