"""Property tests (hypothesis, CPU only) over RANDOM joint trees - mixed joint types, fixed joints inside chains,
floating joints hanging off other bodies, random axes / origins / inertias, contact points, one or two halfspaces
(tests/models.py random_tree) - of the two things everything else leans on without a GPU:

* the oracle (oracle/gp_oracle.cpp, the reference's own formulation) against the independent textbook derivation
  (tests/featherstone_ref.py) and against invariants of the physics that need no second implementation;
* the product's host code (gp_mechanism.cpp through the C ABI): what a mechanism description becomes on its way to the
  kernels.

Seeds are drawn by hypothesis (derandomised: the same examples every run, so a failure reproduces)."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from gorilla_physics_b200 import FIXED, FLOATING, Mechanism
from gorilla_physics_b200.desc import JOINT_NQ
from oracle.binding import OracleMechanism
from tests import featherstone_ref as fs
from tests import models
from tests.test_oracle_independent import rel_err, states

COMMON = dict(deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
trees = st.tuples(st.integers(0, 10 ** 6), st.integers(1, 12))


@settings(max_examples=100, **COMMON)
@given(trees)
def test_oracle_agrees_with_the_independent_derivation_on_random_trees(tree):
    seed, n_bodies = tree
    desc = models.random_tree(seed, n_bodies)
    if desc.n_v == 0:
        return
    orc, ref = OracleMechanism(desc), fs.Model(desc)
    q, v, tau = states(desc, 2, seed=seed + 1)
    for e in range(2):
        a = orc.dynamics(q[e], v[e], tau[e], want="all")
        b = fs.dynamics(ref, q[e], v[e], tau[e])
        cf_b = b["contact_forces"][orc.cp_order] if orc.n_cp else b["contact_forces"]
        assert rel_err(a["mass_matrix"], b["mass_matrix"]) < 1e-11
        assert rel_err(a["bias"], b["bias"]) < 1e-11
        assert rel_err(a["contact_forces"], cf_b) < 1e-11
        # (conditioning of H included: light bodies on long chains reach cond(H) ~ 1e5)
        assert rel_err(a["vdot"], b["vdot"]) < 1e-9


@settings(max_examples=100, **COMMON)
@given(trees)
def test_mass_matrix_is_symmetric_positive_definite_and_consistent_with_the_kinetic_energy(tree):
    seed, n_bodies = tree
    desc = models.random_tree(seed, n_bodies, contact=False)
    if desc.n_v == 0:
        return
    orc = OracleMechanism(desc)
    q, v, _ = states(desc, 1, seed=seed + 2)
    H = orc.dynamics(q[0], v[0], None, want="all")["mass_matrix"]
    np.testing.assert_array_equal(H, H.T)
    assert np.linalg.eigvalsh(H).min() > 0.0
    ke = orc.kinetic_energy(q[0], v[0])  # inertia.rs:182-202, a different code path from the CRBA
    assert abs(0.5 * v[0] @ H @ v[0] - ke) <= 1e-12 * max(1.0, abs(ke))


@settings(max_examples=100, **COMMON)
@given(trees, st.sampled_from([0, 1, 2]))
def test_a_step_keeps_quaternions_unit_and_the_state_packing(tree, integrator):
    seed, n_bodies = tree
    desc = models.random_tree(seed, n_bodies)
    orc = OracleMechanism(desc)
    q, v, tau = states(desc, 1, seed=seed + 3)
    q1, v1 = orc.step(q[0], v[0], tau[0] if desc.n_v else None, dt=1e-3, integrator=integrator)
    assert np.isfinite(q1).all() and np.isfinite(v1).all() and q1.shape == q[0].shape and v1.shape == v[0].shape
    for jt, qo in zip(desc.joint_type, desc.q_offsets()):
        if int(jt) == FLOATING:
            assert abs(np.linalg.norm(q1[qo:qo + 4]) - 1.0) < 1e-14  # renormalised every step (integrators.rs:300-312)
    # the packing of q follows the joint types (joint/mod.rs:208-303); a mechanism of fixed joints only has no state at all
    assert sum(JOINT_NQ[int(jt)] for jt in desc.joint_type) == desc.n_q
    assert (desc.n_q == 0) == all(int(jt) == FIXED for jt in desc.joint_type)


@settings(max_examples=100, **COMMON)
@given(trees)
def test_host_code_round_trips_descriptions_and_derives_the_reference_topology(tree):
    """gp_mechanism_create -> gp_mechanism_get_desc gives the description back (contact points body-major, the order
    the kernels and the oracle list them in), and the supports table (mechanism.rs:118-125) the host code derives is
    the one the oracle derives with the reference's own loop."""
    seed, n_bodies = tree
    desc = models.random_tree(seed, n_bodies)
    mech = Mechanism.from_desc(desc)
    back = mech.desc()
    assert back.n_bodies == desc.n_bodies and back.n_q == desc.n_q and back.n_v == desc.n_v
    for field in ("parent", "joint_type", "axis", "init_iso", "moment", "cross_part", "mass", "has_spring", "spring_k", "spring_l",
                  "hs_point", "hs_normal", "hs_alpha", "hs_mu"):
        np.testing.assert_array_equal(np.asarray(getattr(back, field)), np.asarray(getattr(desc, field)), err_msg=field)
    order = np.argsort(np.asarray(desc.cp_body), kind="stable")
    np.testing.assert_array_equal(np.asarray(back.cp_body), np.asarray(desc.cp_body)[order])
    np.testing.assert_array_equal(np.asarray(back.cp_location).reshape(-1, 3), np.asarray(desc.cp_location).reshape(-1, 3)[order])
    np.testing.assert_array_equal(np.asarray(back.cp_k), np.asarray(desc.cp_k)[order])
    np.testing.assert_array_equal(mech.supports(), OracleMechanism(desc).supports())
    # an unlisted tree never lands on a shipped specialisation by accident: shipped names carry no "jit:" / "generic"
    variant = mech.kernel_variant
    assert variant == "generic" or variant.startswith("jit:") or any(
        variant == k for k in ("pendulum_R", "double_pendulum_RR", "cart_pole_PR", "floating_F", "hopper1d_FPP", "hopper_FPR"))


_ALWAYS_USED = ("init_iso", "moment", "cross_part", "mass", "cp_location", "cp_k", "hs_point", "hs_normal", "hs_alpha", "hs_mu")


@settings(max_examples=200, **COMMON)
@given(trees, st.sampled_from(_ALWAYS_USED), st.sampled_from([float("nan"), float("inf"), float("-inf")]), st.integers(0, 10 ** 6))
def test_non_finite_constants_are_rejected_not_stepped(tree, field, bad, where):
    """Every constant of a description becomes an operand of the step kernels; a NaN or an infinity among them is
    GP_ERR_INVALID at gp_mechanism_create (raw C ABI, as a C caller with an uninitialised array would hit it), never a
    mechanism that turns every environment into NaNs at its first step - and never a crash."""
    import ctypes as C

    from gorilla_physics_b200 import _abi
    seed, n_bodies = tree
    desc = models.random_tree(seed, n_bodies)
    lib = _abi.lib()
    raw, keep = _abi.GpMechanismDesc(), {}
    raw.n_bodies, raw.n_contact_points, raw.n_halfspaces = desc.n_bodies, desc.n_contact_points, desc.n_halfspaces
    for f in ("parent", "joint_type", "has_spring", "cp_body"):
        keep[f] = np.ascontiguousarray(getattr(desc, f), dtype=np.int32).copy()
        setattr(raw, f, keep[f].ctypes.data_as(_abi.ip))
    for f in ("axis", "init_iso", "moment", "cross_part", "mass", "spring_k", "spring_l", "cp_location", "cp_k", "hs_point", "hs_normal",
              "hs_alpha", "hs_mu"):
        keep[f] = np.ascontiguousarray(getattr(desc, f), dtype=np.float64).copy()
        setattr(raw, f, keep[f].ctypes.data_as(C.POINTER(C.c_double)))
    h = C.c_void_p()
    assert lib.gp_mechanism_create(C.byref(raw), C.byref(h)) == _abi.GP_OK  # the description itself is fine
    lib.gp_mechanism_destroy(h)
    keep[field].flat[where % keep[field].size] = bad
    h = C.c_void_p()
    assert lib.gp_mechanism_create(C.byref(raw), C.byref(h)) == _abi.GP_ERR_INVALID, (field, bad)
    assert not h.value


def test_negative_mass_and_non_finite_additions_are_rejected():
    import pytest

    from gorilla_physics_b200._abi import GP_ERR_INVALID, GorillaError
    d = models.random_tree(5, 4)
    d._mass[2] = -0.5
    with pytest.raises(GorillaError) as e:
        Mechanism.from_desc(d)
    assert e.value.code == GP_ERR_INVALID
    m = Mechanism.from_model("cube")
    for call in (lambda: m.add_halfspace((0, 0, 1), float("nan")), lambda: m.add_halfspace((0, 0, 1), 0.0, alpha=float("inf")),
                 lambda: m.add_contact_point(1, (0.0, float("nan"), 0.0)), lambda: m.add_contact_point(1, (0, 0, 0), k=float("inf"))):
        with pytest.raises(GorillaError) as e:
            call()
        assert e.value.code == GP_ERR_INVALID
    assert m.n_halfspaces == 0 and m.n_contact_points == 8  # nothing was added
