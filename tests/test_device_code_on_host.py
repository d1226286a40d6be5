"""The device dynamics code (gp_dynamics.cuh: the body of every step / dynamics kernel) compiled FOR THE HOST
(tools/host_debug.cu, -DGP_HOST_DEBUG: test infrastructure, never part of libgorilla_b200.so - the product has no CPU
path) and held against the oracle on every model, for the compile-time-topology instantiation the library would pick
AND the run-time-topology one. This is what lets a change to the kernel arithmetic be checked without a GPU (the
literal-zero special cases of the root -> leaf pass were developed against it, profiles/r2_tuning.md); the GPU suite
repeats the comparison on the real kernels."""
import re
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def debug_library(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("gpdbg") / "libgpdbg.so"
    cuda = Path("/usr/local/cuda")
    cmd = ["g++", "-DGP_HOST_DEBUG", "-std=c++17", "-O1", "-shared", "-fPIC", f"-I{cuda / 'include'}", "-x", "c++", "-o", str(out),
           str(ROOT / "tools" / "host_debug.cu"), str(ROOT / "gorilla_physics_b200" / "csrc" / "gp_mechanism.cpp"),
           str(ROOT / "gorilla_physics_b200" / "csrc" / "gp_models.cpp"), f"-L{cuda / 'lib64'}", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return out


def test_device_dynamics_code_matches_the_oracle_on_the_host(debug_library):
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "host_debug.py"), "all", str(debug_library)], capture_output=True,
                       text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    worst = {m.group(1): float(m.group(2)) for m in re.finditer(r"^([\w:]+) worst vdot err (\S+)$", r.stdout, re.M)}
    static = re.findall(r"^(\w+) env 0: static-topology vdot err (\S+), contact force vs generic (\S+)$", r.stdout, re.M)
    # every model of the GPU suite ran, through both instantiations
    # (+ 12 random trees with mixed joint types through the run-time-topology instantiation)
    assert len(worst) >= 24 and len(static) >= 12, r.stdout[-2000:]
    for name, err in worst.items():
        assert err < 1e-10, f"{name}: vdot off by {err} (relative) against the oracle"
    for name, _, cf in static:
        assert float(cf) < 1e-9, f"{name}: contact forces of the static and run-time-topology instantiations differ by {cf}"
