"""CPU-only tests: the C ABI library loads and exports every declared symbol, host-side mechanism
logic (validation, flattening, kernel-variant selection), model literals against the golden
fixture extracted from the reference sources, sharding, and the multi-rank path under gloo."""
import ctypes as C
import json
import math
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import gorilla_physics_b200 as gp
from gorilla_physics_b200 import _abi
from gorilla_physics_b200._abi import GorillaError
from gorilla_physics_b200.sharding import shard_range, shard_sizes
from tests import models

ROOT = Path(__file__).resolve().parent.parent


def test_library_loads_and_exports_every_declared_symbol():
    lib = _abi.lib()
    header = (ROOT / "include" / "gorilla_b200.h").read_text()
    declared = set(re.findall(r"\b(gp_[a-z0-9_]+)\s*\(", header))
    declared -= {"gp_status_code"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in gorilla_b200.h but not exported"
    assert declared == set(_abi.SYMBOLS), declared ^ set(_abi.SYMBOLS)
    assert lib.gp_abi_version() == 1


def test_library_does_not_link_torch_or_the_oracle():
    out = subprocess.run(["ldd", str(_abi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "torch" not in out and "gp_oracle" not in out and "libcudart" not in out  # cudart is static
    nm = subprocess.run(["nm", "-D", "--defined-only", str(_abi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "gpo_" not in nm


def test_compute_entry_points_fail_loudly_without_a_gpu():
    if _abi.lib().gp_device_count() > 0:
        pytest.skip("a GPU is visible")
    mech = gp.Mechanism.from_model("so101")
    with pytest.raises(GorillaError) as e:
        gp.MechanismState(mech, 8)
    assert e.value.code == _abi.GP_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(GorillaError):
        gp.measure_fp64_peak()


def test_mechanism_validation_mirrors_reference_panics():
    d = gp.MechanismDesc()
    with pytest.raises(ValueError):  # reference: "joint 1 has no parent body" (mechanism.rs:110-116)
        d.add_body(3, gp.REVOLUTE)
    d.add_body(0, gp.REVOLUTE, axis=(0, 0, 2.0), moment=np.eye(3), mass=1.0)
    with pytest.raises(GorillaError) as e:  # UnitVector3 in the reference: axis must be unit
        gp.Mechanism.from_desc(d)
    assert e.value.code == _abi.GP_ERR_INVALID
    d = gp.MechanismDesc()
    d.add_body(0, gp.REVOLUTE, moment=[[1, 2, 0], [0, 1, 0], [0, 0, 1]], mass=1.0)
    with pytest.raises(GorillaError):
        gp.Mechanism.from_desc(d)
    with pytest.raises(GorillaError):
        gp.Mechanism.from_model("no_such_model")
    with pytest.raises(GorillaError):
        gp.Mechanism.from_model("cube", [1.0])  # wrong parameter count
    m = gp.Mechanism.from_model("so101")
    with pytest.raises(GorillaError):
        m.add_contact_point(99, (0, 0, 0))
    for _ in range(4):
        m.add_halfspace((0, 0, 1), 0.0)
    with pytest.raises(GorillaError) as e:
        m.add_halfspace((0, 0, 1), 0.0)
    assert e.value.code == _abi.GP_ERR_LIMIT


def test_kernel_variant_selection():
    expect = {"pendulum": "pendulum_R", "double_pendulum": "double_pendulum_RR", "cart_pole": "cart_pole_PR",
              "so101": "so101_X6Rz", "cube": "floating_F", "ball": "floating_F", "rimless_wheel": "floating_F",
              "hopper": "hopper_FPR", "hopper_1d": "hopper1d_FPP", "quadruped": "quadruped_F8R",
              "navbot": "navbot_F8Rz"}
    for name, variant in expect.items():
        assert gp.Mechanism.from_model(name).kernel_variant == variant
    # trees without a shipped specialisation: compiled at run time for their own topology when NVRTC is
    # there (gp_jit.cpp), else the run-time-topology kernel; the mode can be forced either way
    unknown = "jit:P" if gp.jit_available() else "generic"
    assert gp.Mechanism.from_model("cart").kernel_variant == unknown
    # an SO-101-shaped chain whose axes are not all +z does not match the shipped so101 kernels
    d = gp.Mechanism.from_model("so101").desc()
    d._axis[3] = np.array([0.0, 1.0, 0.0])
    assert gp.Mechanism.from_desc(d).kernel_variant == ("jit:XRRRRRR" if gp.jit_available() else "generic")
    assert gp.Mechanism.from_desc(d, kernel=gp.KernelMode.GENERIC).kernel_variant == "generic"
    assert gp.Mechanism.from_desc(d, kernel=gp.KernelMode.SHIPPED).kernel_variant == "generic"
    m = gp.Mechanism.from_model("so101")
    assert m.set_kernel_mode(gp.KernelMode.GENERIC).kernel_variant == "generic"
    assert m.set_kernel_mode(gp.KernelMode.SHIPPED).kernel_variant == "so101_X6Rz"
    if gp.jit_available():
        assert m.set_kernel_mode(gp.KernelMode.JIT).kernel_variant == "jit:XRRRRRR"
        # the mode survives later changes of the mechanism
        m.add_halfspace((0, 0, 1), 0.0)
        assert m.kernel_variant == "jit:XRRRRRR"
    assert m.set_kernel_mode(gp.KernelMode.AUTO).kernel_variant == "so101_X6Rz"
    with pytest.raises(GorillaError):
        m.set_kernel_mode(17)


def test_run_time_specialisation_compiles_without_a_gpu(tmp_path, monkeypatch):
    """gp_mechanism_precompile: NVRTC turns the mechanism's signature + the embedded kernel sources into an
    sm_100a cubin in the on-disk cache; a second request is a cache hit. Needs no GPU (loading does)."""
    if not gp.jit_available():
        pytest.skip("NVRTC not loadable here")
    from gorilla_physics_b200.mechanism import JIT_DYNAMICS, JIT_ENERGY, JIT_STEP_SIE
    monkeypatch.setenv("GP_JIT_CACHE", str(tmp_path))
    assert gp.jit_cache_dir() == str(tmp_path)
    d = gp.MechanismDesc()  # a tree no shipped spec covers: revolute - prismatic(spring) - revolute(+z)
    d.add_body(0, gp.REVOLUTE, axis=(0, 1, 0), moment=np.eye(3) * 0.1, cross_part=(0, 0, -0.2), mass=1.0)
    d.add_body(1, gp.PRISMATIC, axis=(1, 0, 0), moment=np.eye(3) * 0.05, mass=0.5, spring=(30.0, 0.1))
    d.add_body(2, gp.REVOLUTE, axis=(0, 0, 1), init_iso=gp.iso((0.1, 0, 0)), moment=np.eye(3) * 0.02, mass=0.3)
    d.add_contact_point(3, (0.0, 0.0, 0.1))
    d.add_halfspace((0, 0, 1), -1.0)
    m = gp.Mechanism.from_desc(d)
    assert m.kernel_variant == "jit:RPR"
    assert m.precompile(JIT_ENERGY) == 1
    files = sorted(p.name for p in tmp_path.iterdir() if p.suffix == ".cubin")
    assert len(files) == 1
    blob = (tmp_path / files[0]).read_bytes()
    assert blob.startswith(b"GPJIT1 _ZN2gp13energy_kernel") and b"\x7fELF" in blob[:200]
    assert m.precompile(JIT_ENERGY) == 0  # cache hit
    # shipped and generic kernels have nothing to compile
    assert gp.Mechanism.from_model("so101").precompile(JIT_STEP_SIE | JIT_DYNAMICS) == 0
    assert gp.Mechanism.from_desc(d, kernel=gp.KernelMode.GENERIC).precompile(JIT_ENERGY) == 0
    # a different contact-point layout is a different specialisation only when the policy changes
    for k in range(4):
        d.add_contact_point(1, (0.0, 0.1 * k, 0.0))
    m2 = gp.Mechanism.from_desc(d)  # body 1 now runs the per-lane contact list: a new kernel
    assert m2.precompile(JIT_ENERGY) == 1


def test_desc_roundtrip_and_contact_point_order():
    m = gp.Mechanism.from_model("quadruped")
    d = m.desc()
    assert (d.n_bodies, d.n_q, d.n_v, d.n_contact_points) == (9, 15, 14, 12)
    # body-major, insertion order within a body: knee bodies carry (knee origin, foot) in that order
    assert list(d.cp_body) == [2, 3, 3, 4, 5, 5, 6, 7, 7, 8, 9, 9]
    assert d.cp_k[1] == 50e3 and d.cp_k[2] == 10e3
    m2 = gp.Mechanism.from_desc(d)
    d2 = m2.desc()
    for f in ("parent", "joint_type", "axis", "init_iso", "moment", "cross_part", "mass", "cp_body", "cp_location", "cp_k"):
        np.testing.assert_array_equal(getattr(d, f), getattr(d2, f))
    q, v = d.zero_state()
    assert q[3] == 1.0 and q.sum() == 1.0 and not v.any()


def test_supports_table():
    S = gp.Mechanism.from_desc(models.supports_fixture()).supports()
    sets = [set(int(i) + 1 for i in np.nonzero(r)[0]) for r in S]
    assert sets == [{1, 2, 3, 4, 5}, {2, 3}, {3}, {4, 5}, {5}]  # reference mechanism.rs:793-832


def test_simulate_step_count():
    f = _abi.lib().gp_simulate_step_count
    assert f(1.0, 0.1) == 11 and f(0.0, 1e-3) == 0
    from oracle.binding import simulate_step_count
    for ft, dt in ((2.0, 1e-3), (30.0, 1e-3), (20.0, 1.0 / 600.0), (0.01, 1.0 / 6000.0)):
        assert f(ft, dt) == simulate_step_count(ft, dt)
    # where the reference's loop would never end, or count past what a launch sequence can hold: -1, at once
    for ft, dt in ((1.0, 0.0), (1.0, -1e-3), (1.0, float("nan")), (float("nan"), 1e-3), (float("inf"), 1e-3), (1e20, 1.0),
                   (1.0, 1e-300), (1e6, 1e-4)):
        assert f(ft, dt) == -1, (ft, dt)
    assert f(-1.0, 1e-3) == 0 and f(1e3, 1e-3) in (1000000, 1000001)


def test_model_literals_match_reference_sources():
    """tests/golden/model_literals.json was extracted from the reference's Rust builders by
    tools/extract_reference_literals.py; every mass / COM / inertia / joint origin must agree."""
    gold = json.loads((ROOT / "tests" / "golden" / "model_literals.json").read_text())
    for model in ("so101", "navbot"):
        d = gp.Mechanism.from_model(model).desc()
        g = gold[model]
        assert d.n_bodies == len(g["bodies"])
        for i, b in enumerate(g["bodies"]):
            m, com = b["m"], np.array(b["com"])
            mc = np.array([[b["ixx"], b["ixy"], b["ixz"]], [b["ixy"], b["iyy"], b["iyz"]], [b["ixz"], b["iyz"], b["izz"]]])
            moment = mc + m * (com @ com * np.eye(3) - np.outer(com, com))
            assert d.mass[i] == m
            np.testing.assert_allclose(d.cross_part[i], m * com, rtol=1e-15, atol=0)
            np.testing.assert_allclose(d.moment[i].reshape(3, 3), moment, rtol=1e-13, atol=1e-20)
            assert int(d.parent[i]) == b["parent"]
            assert int(d.joint_type[i]) == b["joint_type"]
            if b["xyz"] is not None:
                np.testing.assert_array_equal(d.init_iso[i][4:], b["xyz"])
                np.testing.assert_allclose(d.init_iso[i][:4], gp.quat_from_euler(*b["rpy"]), rtol=0, atol=1e-16)
            if b["joint_type"] != gp.FIXED and b["joint_type"] != gp.FLOATING:
                np.testing.assert_array_equal(d.axis[i], b["axis"])


def test_cuboid_models_match_reference_sources():
    """biped / leg / leg_from_foot (builders/biped_builder.rs, leg_builder.rs) are arithmetic over a few lengths, not
    literals: tests/golden/cuboid_models.json holds the reference's statements EVALUATED by
    tools/extract_reference_cuboid_models.py (per body the cuboid's m, w, d, h and centre, the joint, its origin, the
    contact points in the reference's order); the product's builders must produce exactly that mechanism, with
    RigidBody::new_cuboid / new_cuboid_at's inertia (rigid_body.rs:160-195) applied here in numpy."""
    gold = json.loads((ROOT / "tests" / "golden" / "cuboid_models.json").read_text())
    for model, n_bodies, n_v, n_cp in (("biped", 13, 18, 16), ("leg", 6, 11, 24), ("leg_from_foot", 6, 11, 24)):
        d = gp.Mechanism.from_model(model).desc()
        g = gold[model]["bodies"]
        assert d.n_bodies == len(g) == n_bodies and d.n_v == n_v and d.n_contact_points == n_cp
        cps = []
        for i, b in enumerate(g):
            m, w, dd, h, com = b["m"], b["w"], b["d"], b["h"], np.array(b["com"])
            mc = np.diag([m * (dd * dd + h * h) / 12.0, m * (w * w + h * h) / 12.0, m * (w * w + dd * dd) / 12.0])
            moment = mc + m * (com @ com * np.eye(3) - np.outer(com, com))
            assert d.mass[i] == m
            np.testing.assert_allclose(d.cross_part[i], m * com, rtol=1e-15, atol=0)
            np.testing.assert_allclose(d.moment[i].reshape(3, 3), moment, rtol=1e-14, atol=1e-20)
            assert int(d.parent[i]) == b["parent"] and int(d.joint_type[i]) == b["joint_type"]
            np.testing.assert_array_equal(d.init_iso[i], [0.0, 0.0, 0.0, 1.0] + b["xyz"])
            if b["axis"] is not None:
                np.testing.assert_array_equal(d.axis[i], b["axis"])
            cps += [(i + 1, c) for c in b["contacts"]]
        assert [int(x) for x in d.cp_body] == [body for body, _ in cps]  # the reference's order (body-major already)
        np.testing.assert_allclose(np.asarray(d.cp_location).reshape(-1, 3), [c for _, c in cps], rtol=1e-15, atol=0)
        assert all(k == 50e3 for k in np.asarray(d.cp_k))  # ContactPoint::new's default stiffness, contact.rs:24-30


def test_shard_ranges_cover_everything():
    for n, w in ((65536, 8), (262144, 3), (7, 8), (1, 1), (1000003, 8)):
        ranges = [shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = shard_sizes(n, w)
        assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from gorilla_physics_b200 import Mechanism
from gorilla_physics_b200.sharding import shard_range, reduce_diagnostics
from oracle.binding import OracleMechanism
from tests.test_parity_gpu import random_states
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
mech = Mechanism.from_model("double_pendulum"); desc = mech.desc(); orc = OracleMechanism(desc)
n = 37
q, v = random_states(desc, n, seed=5, q_range=3.0)
lo, hi = shard_range(n, rank, world)
# each rank advances only its own environments (no communication on the step path) ...
q1, v1 = orc.batch_rollout(q[lo:hi], v[lo:hi], 1e-3, 50, n_threads=1)
ke = sum(orc.kinetic_energy(q1[e], v1[e]) for e in range(hi - lo))
pe = sum(orc.gravitational_energy(q1[e]) for e in range(hi - lo))
sums = torch.tensor([ke, pe, 0.0, 0.0], dtype=torch.float64)
reduce_diagnostics(sums)                      # ... and only the diagnostic sums are all-reduced
if rank == 0:
    qa, va = orc.batch_rollout(q, v, 1e-3, 50, n_threads=1)
    ke_all = sum(orc.kinetic_energy(qa[e], va[e]) for e in range(n))
    pe_all = sum(orc.gravitational_energy(qa[e]) for e in range(n))
    assert abs(sums[0].item() - ke_all) < 1e-9 * max(1.0, abs(ke_all)), (sums, ke_all)
    assert abs(sums[1].item() - pe_all) < 1e-9 * max(1.0, abs(pe_all)), (sums, pe_all)
    print("gloo-ok")
dist.destroy_process_group()
"""


def test_two_rank_gloo_sharding_and_diagnostic_reduction(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "gloo-ok" in out.stdout


def test_bench_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_limits_of_the_abi():
    """GP_MAX_* (include/gorilla_b200.h): the maximum-size mechanism is accepted, one more of anything is not."""
    from tests import models
    d = models.maximum_size_mechanism()
    m = gp.Mechanism.from_desc(d)
    assert m.kernel_variant in ("generic", "jit:FFRRRPRRRPRRRPXX") and m.desc().n_bodies == 16 and m.desc().n_v == 24
    with pytest.raises(GorillaError) as e:
        m.add_contact_point(1, (0, 0, 0))  # the 33rd contact point
    assert e.value.code == _abi.GP_ERR_LIMIT
    with pytest.raises(GorillaError) as e:
        m.add_halfspace((0, 0, 1), 0.0)  # the 5th halfspace
    assert e.value.code == _abi.GP_ERR_LIMIT
    d17 = models.maximum_size_mechanism()
    d17.add_body(1, gp.FIXED, moment=np.eye(3) * 0.01, mass=0.1)  # the 17th body
    with pytest.raises(GorillaError) as e:
        gp.Mechanism.from_desc(d17)
    assert e.value.code == _abi.GP_ERR_LIMIT
    d25 = gp.MechanismDesc()
    for _ in range(4):
        d25.add_body(0, gp.FLOATING, moment=np.eye(3), mass=1.0)
    d25.add_body(1, gp.REVOLUTE, moment=np.eye(3), mass=1.0)  # the 25th dof
    with pytest.raises(GorillaError) as e:
        gp.Mechanism.from_desc(d25)
    assert e.value.code == _abi.GP_ERR_LIMIT
    # the oracle and the independent derivation agree on the maximum-size mechanism (CPU)
    from oracle.binding import OracleMechanism
    from tests import featherstone_ref as fs
    from tests.test_oracle_independent import states
    orc = OracleMechanism(m.desc())
    ref = fs.Model(m.desc())
    q, v, tau = states(m.desc(), 4, seed=1, q_range=0.5, t_jitter=0.2)
    for e_ in range(4):
        a = orc.dynamics(q[e_], v[e_], tau[e_], want="all")
        b = fs.dynamics(ref, q[e_], v[e_], tau[e_])
        assert np.abs(a["vdot"] - b["vdot"]).max() <= 1e-10 * max(1.0, np.abs(a["vdot"]).max())
        assert np.abs(a["mass_matrix"] - b["mass_matrix"]).max() <= 1e-11 * np.abs(a["mass_matrix"]).max()


def test_sharded_state_fails_loudly_without_a_gpu_and_validates_its_ranges():
    from gorilla_physics_b200 import ShardedMechanismState
    mech = gp.Mechanism.from_model("double_pendulum")
    with pytest.raises(ValueError):
        ShardedMechanismState(mech, 8, devices=[])
    with pytest.raises(ValueError):
        ShardedMechanismState(mech, 2, devices=[0, 1, 2])  # fewer environments than devices
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(GorillaError) as e:  # no CPU path behind the product API
            ShardedMechanismState(mech, 64, devices=[0, 1])
        assert e.value.code == _abi.GP_ERR_NO_DEVICE


@pytest.mark.skipif(not os.environ.get("GP_TEST_SLOW"), reason="builds a second library (~70 s); GP_TEST_SLOW=1")
def test_custom_topology_build_selects_its_own_variant(tmp_path):
    """csrc/Makefile CUSTOM_* + gp_topology.cuh SpecCustom: a library built with one more specialisation picks it
    for the matching tree and nothing else changes."""
    import subprocess
    import sys
    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root / "tools"))
    from custom_topo import custom_topo_vars
    params = [10, 1, 0.5, 1, 0.1, 0.5, 1, 0.1, 0.3, 1, 1, 0.3]
    d = gp.Mechanism.from_model("hopper_2d", params).desc()
    out = tmp_path / "libcustom.so"
    cmd = ["make", "-C", str(root / "gorilla_physics_b200" / "csrc"), "-j8", f"BUILD={tmp_path}/obj", f"OUT={out}"]
    cmd += custom_topo_vars(d, "hopper2d_FRPP").split()
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    code = ("import gorilla_physics_b200 as gp; "
            f"print(gp.Mechanism.from_model('hopper_2d', {params}).kernel_variant, "
            "gp.Mechanism.from_model('so101').kernel_variant, gp.Mechanism.from_model('cart').kernel_variant)")
    res = subprocess.run([sys.executable, "-c", code], env={**os.environ, "GP_LIB_PATH": str(out)}, cwd=root,
                         capture_output=True, text=True, check=True)
    assert res.stdout.split() == ["hopper2d_FRPP", "so101_X6Rz", "generic"]


def test_frozen_workload_descs_match_the_product_builders():
    """tests/golden/workload_descs.json (tools/make_workload_descs.py) is what bench.py's CPU arm hands to the
    oracle instead of asking the product library: it must be exactly what the library would have built."""
    import bench
    from gorilla_physics_b200 import WORKLOADS
    for name, w in WORKLOADS.items():
        d, f = w.mechanism().desc(), bench.frozen_desc(name)
        assert d.n_bodies == f.n_bodies and d.n_contact_points == f.n_contact_points and d.n_halfspaces == f.n_halfspaces
        for field in ("parent", "joint_type", "axis", "init_iso", "moment", "cross_part", "mass", "has_spring", "spring_k",
                      "spring_l", "cp_body", "cp_location", "cp_k", "hs_point", "hs_normal", "hs_alpha", "hs_mu"):
            np.testing.assert_array_equal(np.asarray(getattr(d, field)), np.asarray(getattr(f, field)), err_msg=f"{name}.{field}")


def test_reference_arm_runs_without_the_product_library():
    """bench.py --impl reference: same config keys as the measured arm, and libgorilla_b200.so never loaded"""
    code = ("import sys, json; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0','--envs','64',"
            "'--inner','4']; import bench; bench.main(); "
            "maps=open('/proc/self/maps').read(); assert 'libgorilla_b200' not in maps, 'product library loaded'")
    out = subprocess.run([sys.executable, "-c", code], cwd=str(ROOT), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    for key in ("workload", "n_envs_per_gpu", "inner_steps_per_launch", "dt", "integrator", "controller"):
        assert key in line["config"]
    assert line["config"]["n_envs_per_gpu"] == 64 and line["e2e"]["h2d_bytes_per_step"] == 0


def test_integration_md_binds_every_declared_symbol():
    """INTEGRATION.md's Rust `extern "C"` block is tools/gen_rust_ffi.py's output for the current header: no
    exported function is missing from the binding a maintainer would add, and no `unimplemented!()` is left."""
    sys.path.insert(0, str(ROOT / "tools"))
    import gen_rust_ffi
    text = (ROOT / "INTEGRATION.md").read_text()
    assert gen_rust_ffi.rust_block() in text
    assert "unimplemented!" not in text
    names = set(re.findall(r"pub fn (gp_[a-z0-9_]+)\(", text))
    assert names == set(_abi.SYMBOLS)


def test_round2_entry_points_fail_loudly_without_a_gpu():
    """the entry points added in round 2 keep the library's contract: argument errors are GP_ERR_INVALID, anything
    that would compute returns GP_ERR_NO_DEVICE on a machine without a GPU - never a CPU substitute"""
    lib = _abi.lib()
    if lib.gp_device_count() > 0:
        pytest.skip("a GPU is visible")
    mech = gp.Mechanism.from_model("navbot")
    h = C.c_void_p()
    ids = (C.c_int * 2)(0, 1)
    assert lib.gp_sharded_create(mech._h, 100, ids, 2, C.byref(h)) == _abi.GP_ERR_NO_DEVICE
    assert lib.gp_sharded_create(mech._h, 1, ids, 2, C.byref(h)) == _abi.GP_ERR_INVALID      # fewer environments than devices
    assert lib.gp_sharded_create(None, 100, ids, 2, C.byref(h)) == _abi.GP_ERR_INVALID
    assert lib.gp_sharded_n_shards(None) == 0 and lib.gp_sharded_n_envs(None) == 0
    assert lib.gp_sharded_step(None, 1e-3, 0, 1, 0, None, 0) == _abi.GP_ERR_INVALID
    ident = C.create_string_buffer(128)
    assert lib.gp_comm_create(0, 2, ident, 0, C.byref(h)) == _abi.GP_ERR_NO_DEVICE
    assert lib.gp_comm_create(3, 2, ident, 0, C.byref(h)) == _abi.GP_ERR_INVALID
    out = (C.c_double * 4)()
    assert lib.gp_batch_reduce_diagnostics(None, None, out) == _abi.GP_ERR_INVALID
    assert lib.gp_batch_step_lanes(None) == 0
    assert lib.gp_batch_step_tau_sequence(None, 1e-3, 0, 4, out) == _abi.GP_ERR_INVALID
    n = C.c_int()
    assert lib.gp_measure_fp64_peak_trace(0, 0.1, None, None, 0, C.byref(n)) == _abi.GP_ERR_NO_DEVICE and n.value == 0
    with pytest.raises(GorillaError) as e:
        gp.ShardedMechanismState(mech, 64, devices=[0, 1])
    assert e.value.code == _abi.GP_ERR_NO_DEVICE


def test_unlisted_trees_are_routed_to_run_time_compiled_kernels():
    """a tree no shipped specialisation matches gets a kernel table of its own (compiled on first launch, nothing
    here), unless the mechanism asks for the run-time-topology kernel"""
    if not gp.jit_available():
        pytest.skip("NVRTC not loadable")
    from tests.test_parity_gpu import generic_twin
    from gorilla_physics_b200 import KernelMode
    nav = generic_twin(gp.WORKLOADS["navbot_contact"].mechanism(), KernelMode.JIT)
    arm = generic_twin(gp.WORKLOADS["so101_contact"].mechanism(), KernelMode.AUTO)
    assert nav.kernel_variant == "jit:FRRRRRRRRX" and arm.kernel_variant == "jit:XRRRRRRX"
    assert generic_twin(gp.WORKLOADS["so101_contact"].mechanism(), KernelMode.GENERIC).kernel_variant == "generic"
    assert gp.WORKLOADS["so101_contact"].mechanism().set_kernel_mode(KernelMode.JIT).kernel_variant == "jit:XRRRRRR"


def test_mechanism_create_rejects_partial_descriptions_and_non_unit_normals():
    """A description that announces halfspaces / contact points / spring contacts but leaves their arrays NULL is
    GP_ERR_INVALID, not a segfault; a halfspace normal must be a unit vector (the reference takes a UnitVector3,
    halfspace.rs:6-22: a longer one would scale penetration and force silently)."""
    m = gp.Mechanism.from_model("so101")
    with pytest.raises(GorillaError) as e:
        m.add_halfspace((0, 0, 2.0), 0.0)
    assert e.value.code == _abi.GP_ERR_INVALID and "unit" in str(e.value)
    d = m.desc()
    d.add_halfspace((0, 0.6, 0.6), 0.0)
    with pytest.raises(GorillaError) as e:
        gp.Mechanism.from_desc(d)
    assert e.value.code == _abi.GP_ERR_INVALID
    # the same mechanism through the raw ABI, per-body arrays only
    lib = _abi.lib()
    good = gp.Mechanism.from_model("so101").desc()
    raw, keep = _abi.GpMechanismDesc(), {}
    raw.n_bodies = good.n_bodies
    for f in ("parent", "joint_type", "has_spring"):
        keep[f] = np.ascontiguousarray(getattr(good, f), dtype=np.int32)
        setattr(raw, f, keep[f].ctypes.data_as(_abi.ip))
    for f in ("axis", "init_iso", "moment", "cross_part", "mass", "spring_k", "spring_l"):
        keep[f] = np.ascontiguousarray(getattr(good, f), dtype=np.float64)
        setattr(raw, f, keep[f].ctypes.data_as(C.POINTER(C.c_double)))
    h = C.c_void_p()
    assert lib.gp_mechanism_create(C.byref(raw), C.byref(h)) == _abi.GP_OK  # complete as it is
    lib.gp_mechanism_destroy(h)
    for field in ("n_halfspaces", "n_contact_points", "n_spring_contacts"):
        setattr(raw, field, 2)
        h = C.c_void_p()
        assert lib.gp_mechanism_create(C.byref(raw), C.byref(h)) == _abi.GP_ERR_INVALID, field
        assert not h.value
        setattr(raw, field, 0)


def test_examples_build_against_the_facade_and_fail_loudly_without_a_gpu(tmp_path):
    """examples/*.cpp restate the reference's examples (examples/rimless_wheel.rs, examples/cube.rs, interface/biped.rs
    createBiped) against include/gorilla_b200.hpp. They must compile and link against the library; on a machine without a
    GPU they stop at the first device call with GP_ERR_NO_DEVICE (exit code 2) - there is no CPU path to fall back to."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    import torch
    lib_dir = ROOT / "gorilla_physics_b200" / "lib"
    for src in sorted((ROOT / "examples").glob("*.cpp")):
        exe = tmp_path / src.stem
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe),
                            "-L", str(lib_dir), "-lgorilla_b200", f"-Wl,-rpath,{lib_dir}", "-ldl", "-lpthread", "-lrt"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        if not torch.cuda.is_available():
            r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
            assert r.returncode == 2 and "no CUDA device" in r.stderr, (src.name, r.returncode, r.stderr[-500:])
    if not torch.cuda.is_available():  # the batched Python example: the same, through the ctypes mirror
        import sys
        r = subprocess.run([sys.executable, str(ROOT / "examples" / "batched_so101.py"), "1024", "2"], capture_output=True, text=True,
                           timeout=300)
        assert r.returncode == 2 and "so101_X6Rz" in r.stdout and "no CUDA device" in r.stderr, (r.stdout[-300:], r.stderr[-500:])
