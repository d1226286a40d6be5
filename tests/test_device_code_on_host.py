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


def build_debug_library(out, extra=(), tree=ROOT):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    cuda = Path("/usr/local/cuda")
    cmd = ["g++", "-DGP_HOST_DEBUG", *extra, "-std=c++17", "-O1", "-shared", "-fPIC", f"-I{cuda / 'include'}", "-x", "c++", "-o", str(out),
           str(tree / "tools" / "host_debug.cu"), str(tree / "gorilla_physics_b200" / "csrc" / "gp_mechanism.cpp"),
           str(tree / "gorilla_physics_b200" / "csrc" / "gp_models.cpp"), f"-L{cuda / 'lib64'}", "-lcudart", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return out


@pytest.fixture(scope="module")
def debug_library(tmp_path_factory):
    return build_debug_library(tmp_path_factory.mktemp("gpdbg") / "libgpdbg.so")


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


def run_models(library, names):
    out = ""
    for name in names:
        r = subprocess.run([sys.executable, str(ROOT / "tools" / "host_debug.py"), name, str(library)], capture_output=True, text=True,
                           cwd=ROOT, timeout=600)
        assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
        out += r.stdout
    return out


def test_device_dynamics_code_on_many_random_trees(debug_library):
    """120 more random mechanisms (2 ... 9 bodies, mixed joint types, fixed joints inside chains, floating joints off
    other bodies, springs, 1 ... 8 contact points, one or two halfspaces) through the run-time-topology instantiation
    of the device code, 8 states each, against the oracle"""
    out = run_models(debug_library, ["random_trees:100:120"])
    worst = {m.group(1): float(m.group(2)) for m in re.finditer(r"^(random_tree:\d+) worst vdot err (\S+)$", out, re.M)}
    assert len(worst) == 120, out[-2000:]
    # (1e-9: the conditioning of H of a random tree - light bodies at the end of long chains - is part of the error)
    assert max(worst.values()) < 1e-9, sorted(worst.items(), key=lambda kv: -kv[1])[:5]


def test_cuboid_built_models_on_the_host(debug_library):
    """biped / leg / leg_from_foot (builders/biped_builder.rs, leg_builder.rs; 13 / 6 / 6 bodies, 16 / 24 / 24 contact
    points) on the ground, through the run-time-topology instantiation of the device code, against the oracle"""
    out = run_models(debug_library, ["model:biped", "model:leg", "model:leg_from_foot"])
    worst = {m.group(1): float(m.group(2)) for m in re.finditer(r"^(model:\w+) worst vdot err (\S+)$", out, re.M)}
    assert set(worst) == {"model:biped", "model:leg", "model:leg_from_foot"}, out[-2000:]
    assert max(worst.values()) < 1e-10, worst


def patched_source_tree(dst):
    """A copy of the sources in which the HOST stub of pair_exchange_sum (gp_dynamics.cuh; on the device: shared-memory
    exchange + named barrier between the two warps of a pair) calls the hook of tools/host_debug.cu instead of doing
    nothing. Nothing else differs, and the product's files are not touched."""
    (dst / "tools").mkdir(parents=True)
    shutil.copy(ROOT / "tools" / "host_debug.cu", dst / "tools")
    shutil.copytree(ROOT / "gorilla_physics_b200" / "csrc", dst / "gorilla_physics_b200" / "csrc")
    shutil.copytree(ROOT / "include", dst / "include")
    f = dst / "gorilla_physics_b200" / "csrc" / "gp_dynamics.cuh"
    text, stub = f.read_text(), "  (void)out; (void)slot0; (void)vals;\n"
    assert text.count(stub) == 1
    f.write_text(text.replace(stub, "  gp_host_pair_exchange(out.xch, SIDE, slot0, vals, N);\n"))
    return dst


def test_biped_run_time_specialisation_and_warp_pairs_on_the_host(tmp_path):
    """The biped has no shipped kernel: the library compiles StaticTopo<SpecCustom> for its tree at run time (gp_jit.cpp
    hands NVRTC the SpecCustom macros + the kernel sources), and small batches run it as WARP PAIRS (one leg per warp).
    Both instantiations are built here for the host from the macros the run-time specialisation derives for the tree
    (tools/custom_topo.py + the policy of gp_jit.cpp jit_policy_for: per-lane contact lists on the two feet, the
    right leg = half 1) and held against the oracle. The two halves of a pair run as two threads that meet at the
    exchange points (a pthread barrier for the named barrier, an array for the shared-memory buffer); the shipped pair
    kernels' instantiations (quadruped, navbot) ride along."""
    sys.path.insert(0, str(ROOT / "tools"))
    from custom_topo import custom_topo_vars

    import gorilla_physics_b200 as gp
    kv = dict(x.split("=") for x in custom_topo_vars(gp.Mechanism.from_model("biped").desc(), "biped").split())
    assert kv["CUSTOM_NB"] == "13" and kv["CUSTOM_PARENTS"] == "-1,0,1,2,3,4,5,0,7,8,9,10,11"
    header = tmp_path / "custom_biped.h"
    header.write_text(f'#define GP_CUSTOM_TOPO_NB {kv["CUSTOM_NB"]}\n#define GP_CUSTOM_TOPO_PARENTS {kv["CUSTOM_PARENTS"]}\n'
                      f'#define GP_CUSTOM_TOPO_JOINTS {kv["CUSTOM_JOINTS"]}\n#define GP_CUSTOM_TOPO_AXES {kv["CUSTOM_AXES"]}\n'
                      '#define GP_CUSTOM_TOPO_NAME "biped"\n'
                      '#define GP_CUSTOM_CONTACT_LIST_MASK 0x1040u\n'  # bodies 6, 12: the feet (8 corners each)
                      '#define GP_CUSTOM_SIDE_MASK 0x1f80u\n')         # bodies 7..12: the right leg
    tree = patched_source_tree(tmp_path / "src")
    lib = build_debug_library(tmp_path / "libgpdbg_biped.so", ("-DGP_HOST_PAIRS", "-include", str(header)), tree)
    out = run_models(lib, ["model:biped", "quadruped", "navbot_contact"])
    static = re.findall(r"^model:biped env 0: static-topology vdot err (\S+), contact force vs generic (\S+)$", out, re.M)
    assert len(static) == 1, out[-2000:]
    assert float(static[0][0]) < 1e-10 and float(static[0][1]) < 1e-9
    pairs = dict(re.findall(r"^([\w:]+) env 0: warp-pair vdot err \S+, root copies differ by (\S+)$", out, re.M))
    assert set(pairs) == {"model:biped", "quadruped", "navbot_contact"}, out[-2000:]
    assert all(float(x) == 0.0 for x in pairs.values())  # IEEE sums commute: both halves hold the same root bits
    # (worst covers the run-time-topology, static and warp-pair instantiations of all 8 sampled states)
    worst = {m.group(1): float(m.group(2)) for m in re.finditer(r"^([\w:]+) worst vdot err (\S+)$", out, re.M)}
    assert set(worst) == set(pairs) and max(worst.values()) < 1e-10, worst
