#!/usr/bin/env python
"""Generates the Rust `extern "C"` block for EVERY function declared in include/gorilla_b200.h (the block of
INTEGRATION.md section 2, so that it cannot fall behind the header).   python tools/gen_rust_ffi.py
tests/test_host_cpu.py checks that INTEGRATION.md carries exactly this output."""
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

SCALARS = {"int": "c_int", "double": "c_double", "int32_t": "i32", "int64_t": "i64", "uint32_t": "u32", "uint64_t": "u64",
           "unsigned": "c_uint", "size_t": "usize", "char": "c_char", "void": "c_void"}
OPAQUE = {"gp_mechanism": "GpMechanism", "gp_batch": "GpBatch", "gp_sharded": "GpSharded", "gp_comm": "GpComm",
          "gp_mechanism_desc": "GpMechanismDesc", "gp_state_dist": "GpStateDist"}


def rust_type(ctype: str) -> str:
    t = ctype.strip()
    const = "const" in t.split()
    t = " ".join(w for w in t.split() if w != "const")
    stars = t.count("*")
    base = t.replace("*", "").strip()
    r = SCALARS.get(base) or OPAQUE.get(base)
    if r is None:
        raise ValueError(f"unknown C type {ctype!r}")
    for _ in range(stars):
        r = ("*const " if const else "*mut ") + r
        const = False if stars > 1 else const  # `const T**` does not occur in the header
    return r


def declarations(header: str):
    text = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    for m in re.finditer(r"^([A-Za-z_][\w\s\*]*?)\b(gp_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.M | re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        if ret.startswith("typedef"):
            continue
        yield ret, name, args


def rust_block() -> str:
    header = (ROOT / "include" / "gorilla_b200.h").read_text()
    lines = ['extern "C" {']
    for ret, name, args in declarations(header):
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                arr = re.match(r"(.*?)(\w+)\s*\[\s*\w*\s*\]$", a)   # `const double point[3]` is a pointer
                if arr:
                    ctype, pname = arr.group(1) + "*", arr.group(2)
                else:
                    mm = re.match(r"(.*?)(\w+)$", a)
                    ctype, pname = mm.group(1), mm.group(2)
                if pname in ("type", "box", "ref", "in", "fn", "mod", "move"):
                    pname += "_"
                params.append(f"{pname}: {rust_type(ctype)}")
        r = "" if ret == "void" else f" -> {rust_type(ret)}"
        lines.append(f"    pub fn {name}({', '.join(params)}){r};")
    lines.append("}")
    return "\n".join(lines)


if __name__ == "__main__":
    sys.stdout.write(rust_block() + "\n")
