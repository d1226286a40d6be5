"""The oracle against a SECOND, independently derived CPU restatement (tests/featherstone_ref.py:
textbook body-coordinate RNEA + CRBA with dense 6x6 Pluecker algebra and numpy.linalg.solve) and against
frozen golden vectors (tests/golden/oracle_frozen.json, written by tools/make_oracle_golden.py).

Why: the reference holds no known-answer test for SO-101, for the navbot or for the contact force law on
its own (SURVEY.md §8c "unpinned"); the oracle's other models are pinned by the reference's tests in
test_oracle_golden.py. Two formulations that share no code agreeing to 1e-11 is the strongest CPU-side
pin available here without a Rust toolchain; the frozen vectors keep the oracle from drifting afterwards.
"""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from gorilla_physics_b200 import FLOATING, Mechanism, quat_from_euler
from gorilla_physics_b200.desc import JOINT_NQ
from oracle.binding import OracleMechanism
from tests import featherstone_ref as fs
from tests import models

GOLDEN = Path(__file__).resolve().parent / "golden" / "oracle_frozen.json"
TOL = 1e-11  # relative to the largest entry of the quantity (observed: <= 1e-13)


def states(desc, n, seed, q_range=1.0, v_range=1.0, base_t=(0.0, 0.0, 0.0), t_jitter=0.3, rpy_jitter=0.5):
    rng = np.random.default_rng(seed)
    q = np.zeros((n, desc.n_q))
    v = rng.uniform(-v_range, v_range, size=(n, desc.n_v))
    for jt, qo in zip(desc.joint_type, desc.q_offsets()):
        if jt == FLOATING:
            for e in range(n):
                q[e, qo:qo + 4] = quat_from_euler(*rng.uniform(-rpy_jitter, rpy_jitter, size=3))
            q[:, qo + 4:qo + 7] = np.asarray(base_t) + rng.uniform(-t_jitter, t_jitter, size=(n, 3))
        elif JOINT_NQ[int(jt)] == 1:
            q[:, qo] = rng.uniform(-q_range, q_range, size=n)
    tau = rng.uniform(-1.0, 1.0, size=(n, desc.n_v))
    return q, v, tau


def spring_pair_desc():
    return models.spring_pair()[0]


# name -> (description factory, state kwargs, must the sample contain active contacts?)
CASES = {
    "so101": (lambda: Mechanism.from_model("so101").desc(), dict(), False),
    "so101_contact": (lambda: models.so101_with_contact().desc(), dict(q_range=2.6), True),
    "navbot": (lambda: Mechanism.from_model("navbot").desc(), dict(base_t=(0, 0, 0.075), t_jitter=0.01, rpy_jitter=0.1), False),
    "navbot_contact": (lambda: models.navbot_with_contact().desc(),
                       dict(base_t=(0, 0, 0.02), t_jitter=0.01, rpy_jitter=0.1, q_range=0.2), True),
    "quadruped": (lambda: models.quadruped_on_ground().desc(), dict(base_t=(0, 0, 0.3), t_jitter=0.1, rpy_jitter=0.3), True),
    "hopper_1d": (lambda: models.hopper1d_on_ground().desc(), dict(base_t=(0, 0, -8.0), t_jitter=0.5, rpy_jitter=0.2), True),
    "rimless_wheel": (lambda: models.rimless_wheel_on_slope().desc(), dict(base_t=(0, 0, -10.5), t_jitter=1.0, rpy_jitter=0.4), True),
    "spring_pair": (spring_pair_desc, dict(), False),
    "double_pendulum": (lambda: Mechanism.from_model("double_pendulum").desc(), dict(q_range=math.pi), False),
    "cart_pole": (lambda: Mechanism.from_model("cart_pole").desc(), dict(q_range=math.pi), False),
}


def _grounded(model):
    m = Mechanism.from_model(model)
    m.add_halfspace((0, 0, 1), 0.0)
    return m.desc()


# the reference's cuboid-built trees (builders/biped_builder.rs, leg_builder.rs): 13 / 6 / 6 bodies, 16 / 24 / 24 contact points
CASES["biped"] = (lambda: _grounded("biped"), dict(base_t=(0, 0, 0.6), t_jitter=0.1, rpy_jitter=0.3, q_range=0.5), True)
CASES["leg"] = (lambda: _grounded("leg"), dict(base_t=(0, 0, 0.6), t_jitter=0.1, rpy_jitter=0.3, q_range=0.5), True)
CASES["leg_from_foot"] = (lambda: _grounded("leg_from_foot"), dict(base_t=(0, 0, 0.02), t_jitter=0.05, rpy_jitter=0.3, q_range=0.5), True)
for _s in range(8):
    CASES[f"random_tree_{_s}"] = ((lambda s=_s: models.random_tree(s, 3 + s)), dict(), False)


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-9))


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_independent_featherstone_derivation(name):
    factory, kw, need_contact = CASES[name]
    desc = factory()
    orc = OracleMechanism(desc)
    ref = fs.Model(desc)
    q, v, tau = states(desc, 12, seed=11, **kw)
    active = 0
    for e in range(q.shape[0]):
        a = orc.dynamics(q[e], v[e], tau[e], want="all")
        b = fs.dynamics(ref, q[e], v[e], tau[e])
        cf_b = b["contact_forces"][orc.cp_order] if orc.n_cp else b["contact_forces"]  # oracle lists points body-major
        active += int((np.abs(cf_b).sum(axis=1) > 0).sum())
        assert rel_err(a["mass_matrix"], b["mass_matrix"]) < TOL
        assert rel_err(a["bias"], b["bias"]) < TOL
        assert rel_err(a["contact_forces"], cf_b) < TOL
        assert rel_err(a["vdot"], b["vdot"]) < 1e-10  # north-star per-step tolerance (conditioning of H included)
    if need_contact:
        assert active > 0, "sample never touched the ground: the contact path was not exercised"


def test_contact_force_law_corners_match_independent_statement():
    """C2 (reference contact.rs:260-302) through the oracle: a unit point mass on a floating joint, one
    contact point at its origin, so the reported contact force IS the force law; compared with the plain
    numpy statement of the law at its corners (margin, slip regularisation, separating normal velocity)."""
    from gorilla_physics_b200 import MechanismDesc
    d = MechanismDesc()
    d.add_body(0, FLOATING, moment=np.eye(3), mass=1.0)
    d.add_contact_point(1, (0.0, 0.0, 0.0), k=30e3)
    n = np.array([0.2, -0.1, 1.0])
    n /= np.linalg.norm(n)
    d.add_halfspace(n, 0.0, alpha=0.7, mu=0.4)
    orc = OracleMechanism(d)
    for z in (-1e-9, 0.0, 1e-300, 1e-12, 1e-6, 1e-3, 0.05, 0.3):
        for vel in ((0, 0, 0), (0, 0, -1.0), (0, 0, 5.0), (1e-4, 0, 0), (1e-3, 0, 0), (2e-3, 1e-3, -0.1), (3.0, -2.0, 0.5)):
            q = np.concatenate([[0, 0, 0, 1.0], -z * n])
            v = np.concatenate([[0.3, -0.2, 0.1], vel])  # angular velocity does not move the origin
            got = orc.dynamics(q, v, None, want="all")["contact_forces"][0]
            want = fs.contact_force_law(z, np.asarray(vel, dtype=float), n, 30e3, 0.7, 0.4)
            assert np.abs(got - want).max() <= 1e-12 * max(np.abs(want).max(), 1.0), (z, vel, got, want)


def test_frozen_golden_vectors():
    """tools/make_oracle_golden.py wrote these once from the oracle (after the checks above passed)."""
    gold = json.loads(GOLDEN.read_text())
    for name, rec in gold["cases"].items():
        desc = CASES[name][0]()
        orc = OracleMechanism(desc)
        for s in rec["samples"]:
            a = orc.dynamics(np.array(s["q"]), np.array(s["v"]), np.array(s["tau"]), want="all")
            assert rel_err(s["vdot"], a["vdot"]) < 1e-12, name
            assert rel_err(s["contact_forces"], a["contact_forces"]) < 1e-12, name
        q, v = np.array(rec["rollout"]["q0"]), np.array(rec["rollout"]["v0"])
        q1, v1 = orc.rollout(q, v, rec["rollout"]["dt"], rec["rollout"]["steps"])
        assert rel_err(rec["rollout"]["q1"], q1) < 1e-9, name
        assert rel_err(rec["rollout"]["v1"], v1) < 1e-9, name


def test_counting_build_is_the_same_arithmetic():
    """oracle/gp_oracle_count.cpp (the oracle with a flop-counting scalar type, profiles/roofline.json) must
    compute bit-identical results to the oracle proper, and count a plausible amount of work."""
    import ctypes as C
    import subprocess
    root = Path(__file__).resolve().parent.parent
    subprocess.run(["make", "-s", "-C", str(root / "oracle"), "count"], check=True)
    real = C.CDLL(str(root / "oracle" / "libgp_oracle_count.so"))
    desc = CASES["so101_contact"][0]()
    orc = OracleMechanism(desc)
    q, v, tau = states(desc, 4, seed=3, q_range=2.6)
    want_q, want_v = orc.batch_rollout(q, v, 1.0 / 6000.0, 200, n_threads=1)

    from oracle import binding

    class Proxy:
        def __getattr__(self, name):
            return getattr(real, name.replace("gpo_", "gpc_", 1))

    saved = binding._lib
    try:
        binding._lib = binding.configure(Proxy())
        real.gpc_reset_counters.restype = None
        real.gpc_read_counters.restype = None
        cnt_orc = OracleMechanism(desc)
        real.gpc_reset_counters()
        got_q, got_v = cnt_orc.batch_rollout(q, v, 1.0 / 6000.0, 200, n_threads=1)
        cnt = (C.c_longlong * 6)()
        real.gpc_read_counters(cnt)
        del cnt_orc
    finally:
        binding._lib = saved
    np.testing.assert_array_equal(got_q, want_q)
    np.testing.assert_array_equal(got_v, want_v)
    per_step = sum(cnt) / (4 * 200)
    assert 5e3 < per_step < 2e4, per_step  # ~9e3 flop per SO-101 step in the reference's formulation


@pytest.mark.parametrize("name", ["so101", "navbot", "quadruped_free", "double_pendulum", "hopper_1d_free", "biped_free"])
def test_energy_is_conserved_along_oracle_rollouts(name):
    """A physics pin that needs no second implementation of the dynamics: without contact and torques the
    total energy KE + PE is constant along the true trajectory. The reference's Runge-Kutta 4 advances the
    positions with the OLD velocities (integrators.rs:195-271, euler_step), so its energy error is first order
    in dt: the drift at dt and dt/2 must be in the ratio 2, i.e. the Richardson-extrapolated drift
    2 e(dt/2) - e(dt) vanishes. KE comes from the oracle's own kinetic_energy (inertia.rs:182-202) and,
    independently, from 1/2 v^T H v of the Featherstone derivation; PE is the true potential energy of the
    centres of mass (the reference's gravitational_energy uses frame origins)."""
    if name == "quadruped_free":
        desc = Mechanism.from_model("quadruped").desc()
        kw = dict(base_t=(0, 0, 1.0), t_jitter=0.1, rpy_jitter=0.3)
    elif name == "biped_free":
        desc = Mechanism.from_model("biped").desc()
        kw = dict(base_t=(0, 0, 1.0), t_jitter=0.1, rpy_jitter=0.3, q_range=0.5)
    elif name == "hopper_1d_free":
        desc = Mechanism.from_model("hopper_1d").desc()
        kw = dict(base_t=(0, 0, 1.0), t_jitter=0.1, rpy_jitter=0.3, q_range=0.2)
    else:
        factory, kw, _ = CASES[name]
        desc = factory()
    orc = OracleMechanism(desc)
    ref = fs.Model(desc)
    q, v, _ = states(desc, 3, seed=9, **kw)
    horizon, dt = 0.05, 1e-4
    for e in range(3):
        ke0 = orc.kinetic_energy(q[e], v[e])
        e0 = ke0 + fs.potential_energy(ref, q[e])
        assert abs(ke0 - fs.kinetic_energy(ref, q[e], v[e])) <= 1e-12 * max(1.0, abs(e0))
        drift = []
        for h in (dt, dt / 2):
            q1, v1 = orc.rollout(q[e], v[e], h, int(round(horizon / h)), integrator=2)
            drift.append(orc.kinetic_energy(q1, v1) + fs.potential_energy(ref, q1) - e0)
        scale = max(abs(e0), ke0, 1e-3)
        assert abs(2.0 * drift[1] - drift[0]) <= 1e-7 * scale, (name, e, drift, scale)
        assert abs(drift[1]) <= 1e-3 * scale


@pytest.mark.parametrize("name", ["so101_contact", "navbot_contact", "quadruped", "hopper_1d", "rimless_wheel", "spring_pair", "biped",
                                  "leg_from_foot"])
def test_semi_implicit_euler_step_matches_its_definition(name):
    """I1 (integrators.rs:25-39, :276-319) through the oracle against the update written out from its
    definition on top of the independent dynamics: scalar joints and the quaternion / translation update of
    floating joints (new body-frame twist, old rotation, renormalised quaternion)."""
    factory, kw, _ = CASES[name]
    desc = factory()
    orc = OracleMechanism(desc)
    ref = fs.Model(desc)
    q, v, tau = states(desc, 6, seed=4, **kw)
    dt = 1.0 / 600.0
    for e in range(6):
        q1, v1 = orc.step(q[e], v[e], tau[e], dt=dt, integrator=0)
        q2, v2 = fs.semi_implicit_euler_step(ref, q[e], v[e], tau[e], dt)
        assert rel_err(q1, q2) < 1e-11, (name, e)
        assert rel_err(v1, v2) < 1e-10, (name, e)
