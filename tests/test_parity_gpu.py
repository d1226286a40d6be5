"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star): per-step vdot and contact force 1e-10 relative; one
integrator step 1e-10 relative; short rollouts 1e-8 (chaotic systems only over a bounded horizon).
"Relative" is per environment against the largest magnitude of the reference vector:
|a - b|_inf <= tol * max(|b|_inf, floor).
"""
import math

import numpy as np
import pytest

from gorilla_physics_b200 import (FIXED, FLOATING, Controller, Integrator, KernelMode, Mechanism, MechanismState,
                                  jit_available)
from gorilla_physics_b200.desc import JOINT_NQ, quat_from_euler
from tests import models
from tests.models import oracle_of

pytestmark = pytest.mark.gpu

TOL_DYN = 1e-10
TOL_STEP = 1e-10
TOL_ROLLOUT = 1e-8


def rel_err(a, b, floor=1e-9):
    """per-environment max-norm relative error, reduced to the worst environment"""
    a = np.asarray(a, dtype=float).reshape(len(a), -1)
    b = np.asarray(b, dtype=float).reshape(len(b), -1)
    if a.shape[1] == 0:
        return 0.0
    scale = np.maximum(np.abs(b).max(axis=1), floor)
    return float((np.abs(a - b).max(axis=1) / scale).max())


def rollout_errors(a, ref, floor=1e-6):
    a = np.asarray(a).reshape(len(a), -1)
    ref = np.asarray(ref).reshape(len(ref), -1)
    return np.abs(a - ref).max(axis=1) / np.maximum(np.abs(ref).max(axis=1), floor)


def assert_rollout_parity(orc, q0, v0, q1, v1, dt, steps, tol=TOL_ROLLOUT, amplification=200.0, **roll_kw):
    """Trajectory parity with the problem's own conditioning taken into account.

    Stiff contact (k = 50e3 on 0.1 kg links), clamped PD torques and chaotic pendula amplify ANY
    rounding difference exponentially, so a fixed tolerance over a long horizon tests the system,
    not the kernel. The oracle is therefore also run from the same states perturbed by 1e-15
    relative (about 5 ulp); the kernel must stay within `tol`, or within `amplification` x the
    oracle's own sensitivity to that perturbation, at the median, the 90th percentile and the worst
    environment."""
    q_ref, v_ref = orc.batch_rollout(q0, v0, dt, steps, **roll_kw)
    q_pert, v_pert = orc.batch_rollout(q0 * (1.0 + 1e-15), v0 * (1.0 - 1e-15), dt, steps, **roll_kw)
    # environments in which the REFERENCE algorithm itself blows up (explicit integration of the
    # Hunt-Crossley damping term can diverge from deep initial penetration) carry no parity
    # information; they must be rare, and are left out of the comparison
    sane = np.ones(len(q_ref), dtype=bool)
    for arr in (q_ref, v_ref, q_pert, v_pert):
        a = np.asarray(arr).reshape(len(arr), -1)
        sane &= np.isfinite(a).all(axis=1) & (np.abs(np.nan_to_num(a)).max(axis=1, initial=0.0) < 1e8)
    assert sane.mean() > 0.97, f"reference diverges in {100 * (1 - sane.mean()):.1f}% of the environments"
    for got, ref, pert in ((q1, q_ref, q_pert), (v1, v_ref, v_pert)):
        err = rollout_errors(np.asarray(got)[sane], np.asarray(ref)[sane])
        sens = rollout_errors(np.asarray(pert)[sane], np.asarray(ref)[sane])
        for quant in (0.5, 0.9, 1.0):
            bound = max(tol, amplification * float(np.quantile(sens, quant)))
            assert float(np.quantile(err, quant)) <= bound, (quant, float(np.quantile(err, quant)), bound)
    return q_ref, v_ref


def random_states(desc, n, seed, q_range=1.0, v_range=1.0, base_t=(0.0, 0.0, 0.0), t_jitter=0.3, rpy_jitter=0.5):
    rng = np.random.default_rng(seed)
    q = np.zeros((n, desc.n_q))
    v = rng.uniform(-v_range, v_range, size=(n, desc.n_v))
    for jt, qo in zip(desc.joint_type, desc.q_offsets()):
        if jt == FLOATING:
            rpy = rng.uniform(-rpy_jitter, rpy_jitter, size=(n, 3))
            for e in range(n):
                q[e, qo:qo + 4] = quat_from_euler(*rpy[e])
            q[:, qo + 4:qo + 7] = np.asarray(base_t) + rng.uniform(-t_jitter, t_jitter, size=(n, 3))
        elif JOINT_NQ[int(jt)] == 1:
            q[:, qo] = rng.uniform(-q_range, q_range, size=n)
    return q, v


def generic_twin(mech: Mechanism, kernel: KernelMode = KernelMode.GENERIC) -> Mechanism:
    """Same physics, but a massless fixed leaf is appended so no shipped specialisation matches; run on the
    run-time-topology kernel (GENERIC) or on a kernel compiled at run time for the twin's own tree (JIT)."""
    d = mech.desc()
    d.add_body(d.n_bodies, FIXED, moment=np.zeros((3, 3)), mass=0.0)
    g = Mechanism.from_desc(d, kernel=kernel)
    assert g.kernel_variant == ("generic" if kernel == KernelMode.GENERIC else "jit:" + "".join("XRPF"[int(t)] for t in d.joint_type))
    return g


def kernel_flavour(mech: Mechanism, flavour: str) -> Mechanism:
    """static: the shipped specialisation; generic: the twin on the run-time-topology kernel; jit: the twin on
    its run-time-compiled kernel; jit_same: the mechanism itself, forced onto a run-time-compiled kernel"""
    if flavour == "static":
        return mech
    if flavour == "generic":
        return generic_twin(mech)
    if not jit_available():
        pytest.skip("NVRTC not loadable: no run-time specialisation on this machine")
    if flavour == "jit":
        return generic_twin(mech, KernelMode.JIT)
    return mech.set_kernel_mode(KernelMode.JIT)


# name -> (factory, state kwargs, expected static variant)
WORKLOADS = {
    "pendulum": (lambda: Mechanism.from_model("pendulum"), dict(q_range=math.pi), "pendulum_R"),
    "double_pendulum": (lambda: Mechanism.from_model("double_pendulum"), dict(q_range=math.pi), "double_pendulum_RR"),
    "cart_pole": (lambda: Mechanism.from_model("cart_pole"), dict(q_range=math.pi), "cart_pole_PR"),
    "so101": (lambda: Mechanism.from_model("so101"), dict(), "so101_X6Rz"),
    "so101_contact": (models.so101_with_contact, dict(), "so101_X6Rz"),
    "rimless_wheel": (models.rimless_wheel_on_slope, dict(base_t=(0, 0, -10.5), t_jitter=1.0, rpy_jitter=0.4),
                      "floating_F"),
    "cube": (lambda: _with_ground(Mechanism.from_model("cube"), -0.4), dict(t_jitter=0.2, rpy_jitter=0.6),
             "floating_F"),
    "hopper_1d": (models.hopper1d_on_ground, dict(base_t=(0, 0, -8.0), t_jitter=0.5, rpy_jitter=0.2), "hopper1d_FPP"),
    "hopper": (lambda: _with_ground(Mechanism.from_model("hopper"), 0.0, 1.0, 1.0),
               dict(t_jitter=0.1, rpy_jitter=0.3, q_range=0.3), "hopper_FPR"),
    "quadruped": (models.quadruped_on_ground, dict(base_t=(0, 0, 0.8), t_jitter=0.2, rpy_jitter=0.3),
                  "quadruped_F8R"),
    "navbot": (lambda: Mechanism.from_model("navbot"), dict(base_t=(0, 0, 0.075), t_jitter=0.01, rpy_jitter=0.1,
                                                            q_range=0.2), "navbot_F8Rz"),
    "navbot_contact": (models.navbot_with_contact, dict(base_t=(0, 0, 0.03), t_jitter=0.01, rpy_jitter=0.1,
                                                        q_range=0.2), "navbot_F8Rz"),
}


def _with_ground(m, h, alpha=0.9, mu=0.5):
    m.add_halfspace((0, 0, 1), h, alpha=alpha, mu=mu)
    return m


@pytest.mark.parametrize("name", list(WORKLOADS))
@pytest.mark.parametrize("flavour", ["static", "generic", "jit"])
def test_dynamics_parity(name, flavour):
    factory, kw, variant = WORKLOADS[name]
    mech = factory()
    assert mech.kernel_variant == variant
    mech = kernel_flavour(mech, flavour)
    desc = factory().desc()  # the oracle always sees the original mechanism
    orc = oracle_of(desc)
    n = 1024
    q, v = random_states(desc, n, seed=1234, **kw)
    rng = np.random.default_rng(7)
    tau = rng.uniform(-1.0, 1.0, size=(n, desc.n_v))
    st = MechanismState(mech, n)
    st.update(q, v)
    vdot, cf = st.dynamics(tau=tau, contact_forces=True)
    vdot_ref, cf_ref = orc.batch_dynamics(q, v, tau)
    assert rel_err(vdot, vdot_ref) < TOL_DYN
    if desc.n_contact_points and desc.n_halfspaces:
        assert np.abs(cf_ref).max() > 0.0, "workload never touches the ground: contact path untested"
        assert rel_err(cf, cf_ref) < TOL_DYN
    # zero-torque rule (reference simulate.rs:27-48)
    vdot0 = st.dynamics(tau=None)
    vdot0_ref, _ = orc.batch_dynamics(q, v, None)
    assert rel_err(vdot0, vdot0_ref) < TOL_DYN
    assert not st.status().any()


@pytest.mark.parametrize("name", ["double_pendulum", "so101_contact", "quadruped", "hopper", "navbot_contact"])
def test_mass_matrix_and_bias_parity(name):
    factory, kw, _ = WORKLOADS[name]
    mech = factory()
    desc = mech.desc()
    orc = oracle_of(desc)
    n = 64
    q, v = random_states(desc, n, seed=99, **kw)
    st = MechanismState(mech, n)
    st.update(q, v)
    M, c = st.mass_matrix()
    for e in range(n):
        ref = orc.dynamics(q[e], v[e], None, want="all")
        assert rel_err(M[e][None], ref["mass_matrix"][None]) < 1e-12
        assert rel_err(c[e][None], ref["bias"][None]) < TOL_DYN
        np.testing.assert_array_equal(M[e], M[e].T)


@pytest.mark.parametrize("name", list(WORKLOADS))
@pytest.mark.parametrize("integrator", [Integrator.SemiImplicitEuler, Integrator.RungeKutta2, Integrator.RungeKutta4])
@pytest.mark.parametrize("flavour", ["static", "jit"])
def test_single_step_parity(name, integrator, flavour):
    factory, kw, _ = WORKLOADS[name]
    desc = factory().desc()
    mech = kernel_flavour(factory(), flavour)
    orc = oracle_of(desc)
    n = 256
    q, v = random_states(desc, n, seed=4321, **kw)
    dt = 1.0 / 6000.0
    st = MechanismState(mech, n)
    st.update(q, v)
    st.step(dt, tau=None, integrator=integrator)
    q1, v1 = st.state()
    q_ref, v_ref = orc.batch_rollout(q, v, dt, 1, integrator=int(integrator))
    assert rel_err(q1, q_ref) < TOL_STEP
    assert rel_err(v1, v_ref) < TOL_STEP


@pytest.mark.parametrize("name,dt,steps", [
    ("double_pendulum", 1e-3, 1000), ("cart_pole", 1e-3, 1000), ("so101", 1.0 / 6000.0, 1000),
    ("so101_contact", 1.0 / 6000.0, 1000), ("navbot_contact", 1.0 / 6000.0, 600), ("hopper", 5e-4, 500),
])
@pytest.mark.parametrize("flavour", ["static", "jit"])
def test_rollout_parity_fused_steps(name, dt, steps, flavour):
    """n_steps fused in one launch == the oracle stepping one by one (short horizon, 1e-8)."""
    factory, kw, _ = WORKLOADS[name]
    desc = factory().desc()
    mech = kernel_flavour(factory(), flavour)
    orc = oracle_of(desc)
    n = 128
    q, v = random_states(desc, n, seed=2024, v_range=0.5, **kw)
    st = MechanismState(mech, n)
    st.update(q, v)
    st.step(dt, tau=None, integrator=Integrator.SemiImplicitEuler, n_steps=steps)
    q1, v1 = st.state()
    assert_rollout_parity(orc, q, v, q1, v1, dt, steps, integrator=0)
    # fused == unfused: n single-step launches give bitwise the same state
    st2 = MechanismState(mech, n)
    st2.update(q, v)
    for _ in range(20):
        st2.step(dt, n_steps=1)
    st3 = MechanismState(mech, n)
    st3.update(q, v)
    st3.step(dt, n_steps=20)
    np.testing.assert_array_equal(st2.q, st3.q)
    np.testing.assert_array_equal(st2.v, st3.v)


def test_energy_and_poses_parity():
    for name in ("so101", "quadruped", "hopper", "double_pendulum"):
        factory, kw, _ = WORKLOADS[name]
        mech = factory()
        desc = mech.desc()
        orc = oracle_of(desc)
        n = 64
        q, v = random_states(desc, n, seed=5, **kw)
        st = MechanismState(mech, n)
        st.update(q, v)
        ke, pe, se = st.energies()
        poses = st.poses()
        for e in range(n):
            assert abs(ke[e] - orc.kinetic_energy(q[e], v[e])) <= 1e-11 * max(1.0, abs(ke[e]))
            assert abs(pe[e] - orc.gravitational_energy(q[e])) <= 1e-11 * max(1.0, abs(pe[e]))
            assert abs(se[e] - orc.spring_energy(q[e])) <= 1e-11 * max(1.0, abs(se[e]))
            np.testing.assert_allclose(poses[e], orc.poses(q[e]), rtol=0, atol=1e-12)


def test_in_kernel_controllers_match_oracle():
    # SO101 PD (reference control/so101_control.rs:12-34)
    mech = Mechanism.from_model("so101")
    desc = mech.desc()
    orc = oracle_of(desc)
    n = 64
    q, v = random_states(desc, n, seed=11)
    st = MechanismState(mech, n)
    st.update(q, v)
    params = (1000.0, 0.1, 10.0)
    st.step(1.0 / 6000.0, n_steps=1, controller=Controller.SO101_PD, ctrl_params=params)
    q_ref, v_ref = orc.batch_rollout(q, v, 1.0 / 6000.0, 1, controller=1, params=params)
    assert rel_err(st.q, q_ref) < TOL_STEP and rel_err(st.v, v_ref) < TOL_STEP
    st.update(q, v)
    st.step(1.0 / 6000.0, n_steps=50, controller=Controller.SO101_PD, ctrl_params=params)
    assert_rollout_parity(orc, q, v, st.q, st.v, 1.0 / 6000.0, 50, controller=1, params=params)
    # acrobot swing-up (reference control/swingup.rs:9-69, examples/acrobot.rs), bounded horizon
    mech = Mechanism.from_model("double_pendulum")
    orc = oracle_of(mech)
    q, v = random_states(mech.desc(), n, seed=12, q_range=0.5, v_range=0.2)
    st = MechanismState(mech, n)
    st.update(q, v)
    st.step(1e-3, n_steps=500, controller=Controller.ACROBOT_SWINGUP, ctrl_params=(1.0, 7.0))
    assert_rollout_parity(orc, q, v, st.q, st.v, 1e-3, 500, controller=2, params=(1.0, 7.0))
    # cart-pole swing-up (reference control/swingup.rs:76-110)
    mech = Mechanism.from_model("cart_pole")
    orc = oracle_of(mech)
    q, v = random_states(mech.desc(), n, seed=13, q_range=1.0, v_range=0.2)
    st = MechanismState(mech, n)
    st.update(q, v)
    st.step(1e-2, n_steps=200, controller=Controller.CARTPOLE_SWINGUP, ctrl_params=(1.0, 2.0, 1.0))
    assert_rollout_parity(orc, q, v, st.q, st.v, 1e-2, 200, controller=3, params=(1.0, 2.0, 1.0))
    # single pendulum laws (reference control/mod.rs:57-105): gravity inversion, energy shaping, swing-up + balance
    mech = Mechanism.from_desc(models.hanging_rod_pendulum())
    assert mech.kernel_variant == "pendulum_R"
    orc = oracle_of(mech)
    q, v = random_states(mech.desc(), n, seed=14, q_range=math.pi, v_range=1.0)
    q[:8, 0] = math.pi + np.linspace(-0.2, 0.2, 8)  # both sides of the 0.15 rad switch of the combined law
    for ctrl, integ, oi in ((Controller.PENDULUM_GRAVITY_INVERSION, Integrator.RungeKutta4, 2),
                            (Controller.PENDULUM_ENERGY_SHAPING, Integrator.SemiImplicitEuler, 0),
                            (Controller.PENDULUM_SWINGUP_BALANCE, Integrator.RungeKutta4, 2)):
        st = MechanismState(mech, n)
        st.update(q, v)
        st.step(1e-2, n_steps=1, integrator=integ, controller=ctrl)
        q_ref, v_ref = orc.batch_rollout(q, v, 1e-2, 1, integrator=oi, controller=int(ctrl))
        assert rel_err(st.q, q_ref) < TOL_STEP and rel_err(st.v, v_ref) < TOL_STEP
        st.update(q, v)
        st.step(1e-2, n_steps=300, integrator=integ, controller=ctrl)
        assert_rollout_parity(orc, q, v, st.q, st.v, 1e-2, 300, integrator=oi, controller=int(ctrl))
    # the reference's own two outcome tests (control/mod.rs:146-240) as ONE fused launch each
    st = MechanismState(mech, 1)
    st.update(np.array([[0.1]]), np.array([[0.0]]))
    st.step(1e-2, n_steps=20000, integrator=Integrator.RungeKutta4, controller=Controller.PENDULUM_GRAVITY_INVERSION)
    assert abs(st.q[0, 0] - math.pi) < 1e-3 and abs(st.v[0, 0]) < 1e-4
    # other topologies are refused
    with pytest.raises(Exception):
        MechanismState(Mechanism.from_model("so101"), 4).step(1e-3, controller=Controller.PENDULUM_ENERGY_SHAPING)


def test_full_size_replication_property():
    """BASELINE size (256K SO-101 envs with contact): an environment's result must not depend on
    where it sits in the batch. 64 copies of 4096 distinct states -> all copies bitwise equal, and
    the first copy matches the oracle."""
    mech = models.so101_with_contact()
    desc = mech.desc()
    base_n, copies = 4096, 64
    q0, v0 = random_states(desc, base_n, seed=77)
    q = np.tile(q0, (copies, 1))
    v = np.tile(v0, (copies, 1))
    st = MechanismState(mech, base_n * copies)
    st.update(q, v)
    st.step(1.0 / 6000.0, n_steps=10)
    q1, v1 = st.state()
    q1 = q1.reshape(copies, base_n, -1)
    v1 = v1.reshape(copies, base_n, -1)
    assert (q1 == q1[0]).all() and (v1 == v1[0]).all()
    assert_rollout_parity(oracle_of(desc), q0, v0, q1[0], v1[0], 1.0 / 6000.0, 10)
    assert not st.status().any()


def test_ragged_batch_sizes_and_simulate_host_path():
    mech = models.so101_with_contact()
    desc = mech.desc()
    orc = oracle_of(desc)
    for n in (1, 31, 33, 127, 129, 1000):
        q, v = random_states(desc, n, seed=n)
        st = MechanismState(mech, n)
        q_in, v_in = q.copy(), v.copy()
        nsteps, q_out, v_out = st.simulate(0.01, 1.0 / 6000.0, q_in, v_in)
        from oracle.binding import simulate_step_count
        assert nsteps == simulate_step_count(0.01, 1.0 / 6000.0)
        assert_rollout_parity(orc, q, v, q_out, v_out, 1.0 / 6000.0, nsteps)


def test_status_flags_nan_and_singular():
    mech = Mechanism.from_model("double_pendulum")
    st = MechanismState(mech, 4)
    q = np.zeros((4, 2))
    v = np.zeros((4, 2))
    v[2, 0] = np.nan
    st.update(q, v)
    st.step(1e-3)
    flags = st.status()
    assert flags[2] & 1 and not flags[0] and not flags[1] and not flags[3]


@pytest.mark.parametrize("seed", range(12))
def test_random_trees_on_the_runtime_topology_kernel(seed):
    """Random trees (floating joints hanging off bodies, fixed joints mid-chain, springs, one or two
    halfspaces, up to 16 bodies / 24 dof) through the generic kernel: dynamics, forces, one step."""
    _random_tree_case(seed, int(np.random.default_rng(1000 + seed).integers(1, 17)), KernelMode.GENERIC)


@pytest.mark.parametrize("seed", range(6))
def test_random_trees_on_run_time_compiled_kernels(seed):
    """The same kind of random tree (1-8 bodies here: each one costs three NVRTC compilations unless
    tools/warm_jit_cache.py already put them in the cache) on a kernel compiled for that very tree."""
    if not jit_available():
        pytest.skip("NVRTC not loadable: no run-time specialisation on this machine")
    _random_tree_case(seed, 1 + (3 * seed + 2) % 8, KernelMode.AUTO)


def _random_tree_case(seed, nb, kernel):
    desc = models.random_tree(1000 + seed, nb)
    mech = Mechanism.from_desc(desc, kernel=kernel)
    assert mech.kernel_variant.startswith("generic" if kernel == KernelMode.GENERIC else ("jit:", "pendulum", "floating",
                                                                                         "double_pendulum", "cart_pole", "hopper"))
    orc = oracle_of(mech.desc())
    if desc.n_v == 0:
        pytest.skip("no degrees of freedom drawn")
    n = 96
    q, v = random_states(desc, n, seed=seed, t_jitter=0.3, rpy_jitter=0.8)
    tau = np.random.default_rng(seed).uniform(-1, 1, size=(n, desc.n_v))
    st = MechanismState(mech, n)
    st.update(q, v)
    vdot, cf = st.dynamics(tau=tau, contact_forces=True)
    vdot_ref, cf_ref = orc.batch_dynamics(q, v, tau)
    assert rel_err(vdot, vdot_ref) < TOL_DYN
    assert rel_err(cf, cf_ref, floor=1e-6) < TOL_DYN
    for integ in (Integrator.SemiImplicitEuler, Integrator.RungeKutta4):
        st.update(q, v)
        st.step(1e-4, tau=tau, integrator=integ)
        q_ref, v_ref = orc.batch_rollout(q, v, 1e-4, 1, integrator=int(integ), tau=tau)
        assert rel_err(st.q, q_ref) < TOL_STEP and rel_err(st.v, v_ref) < TOL_STEP


def test_two_halfspaces_contact_mode():
    """cube wedged between ground and a tilted wall: the multi-halfspace step kernels (static, generic and,
    where NVRTC is there, run-time-compiled)"""
    flavours = [models.cube_in_corner(), generic_twin(models.cube_in_corner())]
    if jit_available():
        flavours.append(generic_twin(models.cube_in_corner(), KernelMode.JIT))
    for mech in flavours:
        desc = models.cube_in_corner().desc()
        orc = oracle_of(desc)
        n = 512
        q, v = random_states(desc, n, seed=3, t_jitter=0.15, rpy_jitter=0.6)
        st = MechanismState(mech, n)
        st.update(q, v)
        vdot, cf = st.dynamics(tau=None, contact_forces=True)
        vdot_ref, cf_ref = orc.batch_dynamics(q, v)
        touching_both = ((cf_ref != 0).any(axis=2).sum(axis=1) > 0).mean()
        assert touching_both > 0.2
        assert rel_err(vdot, vdot_ref) < TOL_DYN and rel_err(cf, cf_ref, floor=1e-6) < TOL_DYN
        st.step(1e-3, n_steps=1)
        q_ref, v_ref = orc.batch_rollout(q, v, 1e-3, 1)
        assert rel_err(st.q, q_ref) < TOL_STEP and rel_err(st.v, v_ref) < TOL_STEP
        st.update(q, v)
        st.step(1e-3, n_steps=100)
        assert_rollout_parity(orc, q, v, st.q, st.v, 1e-3, 100)


def test_mechanism_without_degrees_of_freedom_and_per_env_tau():
    from gorilla_physics_b200 import MechanismDesc, REVOLUTE
    d = MechanismDesc()
    d.add_body(0, FIXED, moment=np.eye(3), mass=1.0)
    d.add_body(1, FIXED, moment=np.eye(3), mass=1.0)
    st = MechanismState(Mechanism.from_desc(d), 5)
    st.step(1e-3, n_steps=3)  # nothing to integrate, must not fault
    q, v = st.state()
    assert q.shape == (5, 0) and v.shape == (5, 0)
    assert not st.status().any()
    # distinct torques per environment reach the right environment
    mech = Mechanism.from_model("pendulum")
    orc = oracle_of(mech)
    n = 77
    tau = np.linspace(-50, 50, n).reshape(n, 1)
    st = MechanismState(mech, n)
    vdot = st.dynamics(tau=tau)
    ref, _ = orc.batch_dynamics(np.zeros((n, 1)), np.zeros((n, 1)), tau)
    assert rel_err(vdot, ref) < TOL_DYN and np.unique(np.round(vdot, 9)).size == n


def test_simulate_history_matches_reference_layout():
    """simulate() returns every state including the initial one (reference simulate.rs:99-108)"""
    mech = Mechanism.from_model("double_pendulum")
    orc = oracle_of(mech)
    n = 9
    q, v = random_states(mech.desc(), n, seed=8, q_range=1.0)
    st = MechanismState(mech, n)
    nsteps, hq, hv = st.simulate(0.02, 1e-3, q.copy(), v.copy(), history=True)
    assert hq.shape == (nsteps + 1, n, 2) and hv.shape == (nsteps + 1, n, 2)
    np.testing.assert_array_equal(hq[0], q)
    for e in range(n):
        _, _, rq, rv = orc.rollout(q[e], v[e], 1e-3, nsteps, history=True)
        assert np.abs(hq[:, e] - rq).max() < 1e-10 and np.abs(hv[:, e] - rv).max() < 1e-9
    # the python-level simulate() with a host closure (the reference's control_fn)
    from gorilla_physics_b200 import simulate as py_simulate
    st.update(q, v)
    qs, vs = py_simulate(st, 0.005, 1e-3, control_fn=lambda s: np.zeros((n, 2)))
    np.testing.assert_allclose(qs, hq[:qs.shape[0]], rtol=0, atol=1e-13)


def test_stateful_hopper_controller_matches_oracle():
    """Hopper1DController in-kernel (reference control/energy_control.rs:24-101): per-environment
    controller state persists across launches; parity over several hops."""
    mech = models.hopper1d_on_ground()
    desc = mech.desc()
    orc = oracle_of(desc)
    n = 200
    q, v = random_states(desc, n, seed=21, base_t=(0, 0, -2.0), t_jitter=2.0, rpy_jitter=0.0, q_range=0.0, v_range=0.0)
    q[:, 0:3] = 0.0
    q[:, 3] = 1.0
    params = (200.0, 0.0, 2.0, 10.0)
    dt = 1.0 / 500.0
    st = MechanismState(mech, n)
    st.update(q, v)
    # 3000 steps split over several launches: the controller state must carry over
    for chunk in (1, 499, 1000, 1500):
        st.step(dt, n_steps=chunk, controller=Controller.HOPPER_1D, ctrl_params=params)
    assert_rollout_parity(orc, q, v, st.q, st.v, dt, 3000, controller=4, params=params)
    cs = st.controller_state()
    assert cs.shape == (n, 2) and np.isfinite(cs).all() and np.abs(cs[:, 1]).max() > 0
    st.set_controller_state(None)
    assert not st.controller_state().any()
    # a mechanism that is not floating + prismatic + prismatic is refused
    other = MechanismState(Mechanism.from_model("so101"), 4)
    with pytest.raises(Exception):
        other.step(dt, controller=Controller.HOPPER_1D, ctrl_params=params)


def test_pipelined_simulate_matches_resident_stepping():
    """gp_batch_simulate cuts big batches into chunks and overlaps copies with rollouts on two streams;
    the result must be bitwise what stepping the resident state gives (ragged last chunk included)."""
    mech = models.so101_with_contact()
    desc = mech.desc()
    n = 70001
    q, v = random_states(desc, n, seed=5)
    tau = np.random.default_rng(5).uniform(-0.05, 0.05, size=(n, desc.n_v))
    dt = 1.0 / 6000.0
    ref = MechanismState(mech, n)
    ref.update(q, v)
    ref.step(dt, tau=tau, n_steps=16)
    q_ref, v_ref = ref.state()
    st = MechanismState(mech, n)
    n_steps, q_out, v_out = st.simulate(15.5 * dt, dt, q.copy(), v.copy(), tau=tau)
    assert n_steps == 16
    np.testing.assert_array_equal(q_out, q_ref)
    np.testing.assert_array_equal(v_out, v_ref)
    # device state after the call is the final state too
    q_dev, v_dev = st.state()
    np.testing.assert_array_equal(q_dev, q_ref)
    # and without torques
    ref.update(q, v)
    ref.step(dt, tau=None, n_steps=8)
    n_steps, q_out, v_out = st.simulate(7.5 * dt, dt, q.copy(), v.copy())
    np.testing.assert_array_equal(q_out, ref.q)


def test_free_velocity_and_armature():
    """Articulated::free_velocity (reference hybrid/articulated/mod.rs:124-197): v + M^-1 (tau - c) dt with
    armature on the diagonal, no contact, optional gravity — against the oracle's restatement."""
    desc = Mechanism.from_model("so101").desc()
    for i in range(1, 7):
        desc._armature[i] = 1e-4 * (i + 1)
    desc.add_contact_point(7, (0, 0, 0))
    desc.add_halfspace((0, 0, 1), 0.5)  # everything is "in contact": free_velocity must ignore it
    mech = Mechanism.from_desc(desc)
    assert mech.kernel_variant == "so101_X6Rz"
    orc = oracle_of(mech.desc())
    n = 300
    q, v = random_states(desc, n, seed=31)
    tau = np.random.default_rng(31).uniform(-0.2, 0.2, size=(n, 6))
    st = MechanismState(mech, n)
    st.update(q, v)
    for gravity in (True, False):
        vf = st.free_velocity(1e-3, gravity_enabled=gravity, tau=tau)
        ref = np.stack([orc.free_velocity(q[e], v[e], 1e-3, tau=tau[e], gravity_enabled=gravity) for e in range(n)])
        assert rel_err(vf, ref) < TOL_DYN
    np.testing.assert_array_equal(st.free_velocity(0.0), v)
    # the armature does NOT enter MechanismState's own path (dynamics / step / mass_matrix), as in the reference:
    # same accelerations as the mechanism without armature
    vdot = st.dynamics(tau=tau)
    vdot_ref, _ = orc.batch_dynamics(q, v, tau)
    assert rel_err(vdot, vdot_ref) < TOL_DYN
    bare = mech.desc()
    for i in range(bare.n_bodies):
        bare._armature[i] = 0.0
    plain = MechanismState(Mechanism.from_desc(bare), n)
    plain.update(q, v)
    np.testing.assert_array_equal(plain.dynamics(tau=tau), vdot)
    # quadruped (floating base, generic axes)
    mech = Mechanism.from_model("quadruped")
    orc = oracle_of(mech)
    q, v = random_states(mech.desc(), 64, seed=32, base_t=(0, 0, 0.8), rpy_jitter=0.3)
    st = MechanismState(mech, 64)
    st.update(q, v)
    vf = st.free_velocity(1.0 / 3000.0)
    ref = np.stack([orc.free_velocity(q[e], v[e], 1.0 / 3000.0) for e in range(64)])
    assert rel_err(vf, ref) < TOL_DYN


@pytest.mark.parametrize("flavour", ["static", "generic", "jit"])
def test_spring_contact_slip_matches_oracle(flavour):
    """SpringContact (reference contact.rs:74-94, :133-186): the stateful SLIP leg. Many hoppers with
    different launch speeds, stepped on the GPU and in the oracle, the apex logic of SLIP_hopping
    (contact.rs:878-889) applied on the host to both; state AND spring-contact state must agree.
    Single-floating-body specialisation and the run-time-topology kernel."""
    mech = Mechanism.from_model("slip")
    mech.add_halfspace((0, 0, 1), -0.3)
    assert mech.kernel_variant == "floating_F" and mech.n_spring_contacts == 1
    orc = oracle_of(mech)
    if flavour != "static":
        mech = kernel_flavour(mech, flavour)
        assert mech.n_spring_contacts == 1 and mech.n_halfspaces == 1
    a = math.radians(45.0)
    direction = np.array([math.sin(a), 0.0, -math.cos(a)])
    direction /= np.linalg.norm(direction)
    n = 48
    rng = np.random.default_rng(3)
    q = np.tile(np.array([0, 0, 0, 1.0, 0, 0, 0]), (n, 1))
    v = np.zeros((n, 6))
    v[:, 3] = rng.uniform(3.0, 6.0, size=n)
    v[:, 5] = rng.uniform(-0.5, 0.5, size=n)
    dt = 1.0 / 2000.0
    st = MechanismState(mech, n)
    st.update(q, v)
    sc = st.spring_contact_state()
    np.testing.assert_array_equal(sc[:, 0], np.tile(orc.spring_state_init()[0], (n, 1)))
    qo, vo = q.copy(), v.copy()
    sco = np.stack([orc.spring_state_init() for _ in range(n)])
    vz_prev = v[:, 5].copy()
    registered_seen = detached_seen = False
    for step in range(1200):
        st.step(dt)
        qg, vg = st.state()
        scg = st.spring_contact_state()
        for e in range(n):
            qo[e], vo[e], sco[e], flags = orc.step_sc(qo[e], vo[e], sco[e], dt)
            assert flags == 0
        registered_seen |= bool((sco[:, 0, 0] != 0).any())
        apex = (vz_prev > 0.0) & (vo[:, 5] <= 0.0)
        if apex.any():
            detached_seen = True
            sco[apex, 0, 4:7] = direction
            sco[apex, 0, 7] = 0.2
            scg[apex, 0, 4:7] = direction
            scg[apex, 0, 7] = 0.2
            st.set_spring_contact_state(scg)
        vz_prev = vo[:, 5].copy()
        if step % 100 == 99 or step < 3:
            assert rel_err(qg, qo) < 1e-9 and rel_err(vg, vo) < 1e-8, step
            np.testing.assert_array_equal(scg[:, 0, 0], sco[:, 0, 0])  # same registration events
            np.testing.assert_allclose(scg, sco, rtol=0, atol=1e-9)
    assert registered_seen and detached_seen
    assert not st.status().any()
    # Runge-Kutta is refused with spring contacts (reference simulate.rs:57-69)
    with pytest.raises(Exception):
        st.step(dt, integrator=Integrator.RungeKutta4)
    st.set_spring_contact_state(None)
    np.testing.assert_array_equal(st.spring_contact_state()[:, 0], np.tile(orc.spring_state_init()[0], (n, 1)))


def test_quadruped_trot_to_position_on_the_gpu():
    """The reference's own end-to-end test of the 14-dof + 12-contact-point model
    (control/quadruped_control.rs:415-478), with the host closure path: one controller per
    environment computes torques on the host, the GPU does step(). Same acceptance as the reference
    (|x - target| < 0.1, |v_x| < 0.3 after 3 s), for several targets at once, and the first part of the
    trajectory against the oracle driven by the same controller."""
    from tests.controllers_ref import QuadrupedTrottingController, quadruped_initial_state
    mech = models.quadruped_on_ground()
    assert mech.kernel_variant.startswith("quadruped")
    orc = oracle_of(mech)
    dt = 1.0 / (60.0 * 50.0)
    targets = [-0.2, -0.6, -1.5, -1.8]
    n = len(targets)
    q0, v0 = quadruped_initial_state()
    st = MechanismState(mech, n)
    st.update(np.tile(q0, (n, 1)), np.tile(v0, (n, 1)))
    ctrls = [QuadrupedTrottingController(dt, t, -0.8) for t in targets]
    ctrl_o = QuadrupedTrottingController(dt, targets[0], -0.8)
    qo, vo = q0.copy(), v0.copy()
    q, v = st.state()
    check_until = 600
    for step in range(int(3.0 / dt)):
        tau = np.stack([c.control(q[e], v[e]) for e, c in enumerate(ctrls)])
        st.step(dt, tau=tau)
        q, v = st.state()
        if step < check_until:
            qo, vo = orc.step(qo, vo, ctrl_o.control(qo, vo), dt, 0)
            if step in (0, 9, 99, check_until - 1):
                assert np.abs(q[0] - qo).max() < 1e-9 * (step + 1) and np.abs(v[0] - vo).max() < 1e-8 * (step + 1), step
    assert not st.status().any()
    for e, t in enumerate(targets):
        assert abs(q[e, 4] - t) < 1e-1, (e, q[e, 4])
        x, y, z, w = q[e, 0:4]
        Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                       [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                       [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        assert abs((Rq.T @ v[e, 3:6])[0]) < 3e-1


@pytest.mark.gpu
def test_huge_joint_angles_take_the_library_reduction():
    """The step kernels evaluate the joint sines/cosines of an environment together with a three-part
    pi/2 reduction that is only valid below 1e5 rad; an environment holding a larger angle must fall back
    to the library's exact (Payne-Hanek) reduction for all its joints (reference: libm sin_cos through
    nalgebra, revolute.rs:97-102). Static and run-time-topology kernels, dynamics and one fused step."""
    for factory in (lambda: Mechanism.from_model("so101"), lambda: Mechanism.from_model("double_pendulum")):
        mech = factory()
        desc = mech.desc()
        orc = oracle_of(desc)
        n = 64
        q, v = random_states(desc, n, seed=77)
        q[1, 0] = 123456.789            # one huge angle next to ordinary ones
        q[2, -1] = -9.87654321e7
        q[3, :] = 1.0e5                 # exactly at the switch-over
        q[4, :] = np.nextafter(1.0e5, 0.0)
        q[5, 0] = 3.0e15                # spacing of doubles here is 0.5 rad: still a defined angle
        st = MechanismState(mech, n)
        st.update(q, v)
        vdot = st.dynamics(tau=None)
        vdot_ref, _ = orc.batch_dynamics(q, v, None)
        assert rel_err(vdot, vdot_ref) < TOL_DYN
        dt = 1.0e-4
        st.step(dt, n_steps=1)
        q1, v1 = st.state()
        q_ref, v_ref = orc.batch_rollout(q, v, dt, 1)
        assert rel_err(v1, v_ref) < TOL_DYN
        assert np.abs(q1 - q_ref).max() <= 1e-10 * np.maximum(np.abs(q_ref), 1.0).max()
        assert not st.status().any()


@pytest.mark.gpu
def test_contact_force_law_corner_cases():
    """Hunt-Crossley + regularised Coulomb (reference contact.rs:260-302) at the places where its branches
    meet: penetrations from 1e-300 to 0.3 (the kernel takes z^(3/2) from a reciprocal square root seed),
    the (0, 1e-8] margin where powf(negative, 1.5) is NaN and the force is zero, sliding speeds on both
    sides of the 1e-3 regularisation, no sliding at all. One ball per environment with a contact point at
    its origin (placed exactly) and one off-centre."""
    mech = Mechanism.from_desc(models.ball())
    mech.add_contact_point(1, (0.0, 0.0, 0.0))      # at the body origin: the penetration is exactly -t_z
    mech.add_contact_point(1, (0.25, 0.0, -1.0e-3))  # and one off-centre point (moment arm), 1 mm lower
    mech = _with_ground(mech, 0.0)
    desc = mech.desc()
    orc = oracle_of(desc)
    depths = [1e-300, 1e-200, 1e-30, 1e-17, 1e-12, 1e-9, 1e-6, 1e-3, 0.3, 0.0, -1e-9, -9.9e-9, -1.0e-8, -1.1e-8]
    speeds = [0.0, 1e-30, 1e-9, 9.99e-4, 1.0e-3, 1.001e-3, 0.5]
    states = []
    for d in depths:
        for s in speeds:
            for vz in (-0.3, 0.0, 0.4):
                qv = np.zeros(desc.n_q)
                qv[3] = 1.0                    # identity quaternion (x, y, z, w)
                qv[6] = -d                     # the origin sits d below the ground plane z = 0
                vv = np.zeros(desc.n_v)
                vv[3], vv[4], vv[5] = s, 0.5 * s, vz
                states.append((qv, vv))
    q = np.array([s[0] for s in states])
    v = np.array([s[1] for s in states])
    st = MechanismState(mech, len(states))
    st.update(q, v)
    vdot, cf = st.dynamics(tau=None, contact_forces=True)
    vdot_ref, cf_ref = orc.batch_dynamics(q, v, None)
    assert np.isfinite(cf).all() and np.isfinite(vdot).all()
    assert (np.abs(cf_ref).reshape(len(states), -1).max(axis=1) > 0).sum() > len(states) // 3
    # per contact point, relative to that point's own force (they span 300 orders of magnitude)
    scale = np.maximum(np.abs(cf_ref).max(axis=2, keepdims=True), 1e-290)
    assert (np.abs(cf - cf_ref) / scale).max() < TOL_DYN
    # where the reference gives exactly zero force (margin, separating faster than the spring pushes) so do we
    assert not cf[np.abs(cf_ref).max(axis=2) == 0.0].any()
    assert rel_err(vdot, vdot_ref) < TOL_DYN
    # and the step kernel's copy of the law (CONTACT == 1 instantiation) agrees with the oracle after one step
    dt = 1.0e-4
    st.step(dt, n_steps=1)
    q1, v1 = st.state()
    q_ref, v_ref = orc.batch_rollout(q, v, dt, 1)
    assert rel_err(v1, v_ref) < TOL_DYN
    assert not st.status().any()


@pytest.mark.parametrize("name", ["so101", "so101_contact", "navbot", "navbot_contact", "quadruped", "hopper_1d",
                                  "rimless_wheel"])
def test_committed_golden_vectors(name):
    """The CUDA path against tests/golden/oracle_frozen.json (tools/make_oracle_golden.py): the models the
    reference holds no known-answer test for, frozen after the oracle was cross-checked against an
    independent derivation (tests/test_oracle_independent.py). No oracle code runs in this test."""
    import json
    from pathlib import Path
    from tests.test_oracle_independent import CASES
    rec = json.loads((Path(__file__).resolve().parent / "golden" / "oracle_frozen.json").read_text())["cases"][name]
    mech = Mechanism.from_desc(CASES[name][0]())
    q = np.array([s["q"] for s in rec["samples"]])
    v = np.array([s["v"] for s in rec["samples"]])
    tau = np.array([s["tau"] for s in rec["samples"]])
    st = MechanismState(mech, len(q))
    st.update(q, v)
    vdot, cf = st.dynamics(tau=tau, contact_forces=True)
    assert rel_err(vdot, np.array([s["vdot"] for s in rec["samples"]])) < TOL_DYN
    if mech.desc().n_contact_points:
        assert rel_err(cf, np.array([s["contact_forces"] for s in rec["samples"]])) < TOL_DYN
    ro = rec["rollout"]
    st = MechanismState(mech, 1)
    st.update(np.array([ro["q0"]]), np.array([ro["v0"]]))
    st.step(ro["dt"], integrator=Integrator.SemiImplicitEuler, n_steps=ro["steps"])
    q1, v1 = st.state()
    assert rel_err(q1, np.array([ro["q1"]])) < TOL_ROLLOUT
    assert rel_err(v1, np.array([ro["v1"]]), floor=1e-3) < 1e-6  # 50 steps through stiff contact onset


@pytest.mark.parametrize("name", ["so101_contact", "navbot_contact", "quadruped"])
def test_dynamics_against_independent_featherstone_derivation(name):
    """The CUDA path directly against the textbook body-coordinate RNEA + CRBA in numpy
    (tests/featherstone_ref.py), which shares no code with the oracle or the kernels."""
    from tests import featherstone_ref as fs
    from tests.test_oracle_independent import CASES, states
    factory, kw, _ = CASES[name]
    desc = factory()
    mech = Mechanism.from_desc(desc)
    q, v, tau = states(desc, 32, seed=5, **kw)
    st = MechanismState(mech, len(q))
    st.update(q, v)
    vdot = st.dynamics(tau=tau)
    ref = fs.Model(desc)
    want = np.array([fs.dynamics(ref, q[e], v[e], tau[e])["vdot"] for e in range(len(q))])
    assert rel_err(vdot, want) < TOL_DYN


@pytest.mark.parametrize("name,n_copies,base_n,steps,dt", [
    ("navbot_contact", 32, 2048, 37, 1.0 / 6000.0),     # 65536 envs: 256 blocks on 148 one-block SMs -> ticket mode
    ("rimless_wheel", 64, 4096, 40, 1.0 / 600.0),       # 262144 envs, 2048 blocks on 592 slots -> ticket mode
    ("quadruped", 33, 2000, 16, 1.0 / 3000.0),          # ragged: 66000 envs, last block partly filled
    ("so101_contact:generic", 33, 2000, 16, 1.0 / 6000.0),  # run-time-topology kernel: per-warp work items, ragged (66000 envs)
])
def test_ticket_mode_replication_property(name, n_copies, base_n, steps, dt):
    """Batches whose blocks do not fill whole waves run in ticket mode (gp_kernels.cuh: the fused steps are
    cut into chunks that a persistent grid draws from a counter, the state of an environment block travelling
    through the q / v planes between chunks, possibly across SMs). The result must be what the plain mode
    gives: copies of the same states are bitwise equal wherever they sit in the batch, equal to a small
    batch (one wave, plain mode) of the same states bit for bit (to rounding where the small batch runs the
    warp-pair mapping), and within parity of the oracle."""
    name, _, flavour = name.partition(":")
    factory, kw, _ = WORKLOADS[name]
    mech = kernel_flavour(factory(), flavour or "static")
    desc = mech.desc()
    q0, v0 = random_states(desc, base_n, seed=5, **kw)
    small = MechanismState(mech, 256)
    small.update(q0[:256], v0[:256])
    small.step(dt, n_steps=steps)
    qs, vs = small.state()
    st = MechanismState(mech, base_n * n_copies)
    st.update(np.tile(q0, (n_copies, 1)), np.tile(v0, (n_copies, 1)))
    st.step(dt, n_steps=steps)
    q1, v1 = st.state()
    q1 = q1.reshape(n_copies, base_n, -1)
    v1 = v1.reshape(n_copies, base_n, -1)
    same = lambda a, b: np.array_equal(a, b, equal_nan=True)  # (a few deep-penetration states blow up, as in the reference)
    assert all(same(q1[c], q1[0]) and same(v1[c], v1[0]) for c in range(n_copies))
    if mech.kernel_variant in ("navbot_F8Rz", "quadruped_F8R"):
        # topologies with halves: the small batch ran the warp-pair mapping (two warps per 32 environments, the
        # halves' sums meeting at the root), the big one a thread per environment: same numbers to rounding
        # (per environment; the few deep-penetration states in which the reference algorithm itself blows up amplify
        # the last bit without bound and carry no information)
        ok = np.isfinite(qs).all(axis=1) & np.isfinite(q1[0][:256]).all(axis=1)
        eq, ev = rollout_errors(q1[0][:256][ok], qs[ok]), rollout_errors(v1[0][:256][ok], vs[ok], floor=1e-3)
        assert ok.mean() > 0.97 and np.median(eq) < 1e-12 and np.quantile(eq, 0.9) < 1e-9 and np.quantile(ev, 0.9) < 1e-7
    else:
        assert same(q1[0][:256], qs) and same(v1[0][:256], vs)
    assert_rollout_parity(oracle_of(desc), q0[:512], v0[:512], q1[0][:512], v1[0][:512], dt, steps)
    # and through host buffers (simulate(): chunks on two streams, each with its own ticket scratch)
    q_in = np.tile(q0, (n_copies, 1))
    v_in = np.tile(v0, (n_copies, 1))
    nsteps, q_out, v_out = st.simulate((steps - 0.5) * dt, dt, q_in, v_in)
    assert nsteps == steps
    q_out = q_out.reshape(n_copies, base_n, -1)
    v_out = v_out.reshape(n_copies, base_n, -1)
    assert all(same(q_out[c], q1[0]) and same(v_out[c], v1[0]) for c in range(n_copies))


@pytest.mark.parametrize("kernel", [KernelMode.GENERIC, KernelMode.JIT], ids=["generic", "jit"])
def test_maximum_size_mechanism(kernel):
    """16 bodies / 24 dofs / 32 contact points / 4 halfspaces, all limits of the ABI at once, against the
    oracle and against the independent derivation: dynamics, contact forces, one step of every integrator,
    a fused rollout, and a ragged batch. On the run-time-topology kernel and on a kernel compiled for this tree."""
    from tests import featherstone_ref as fs
    desc = models.maximum_size_mechanism()
    if kernel == KernelMode.JIT and not jit_available():
        pytest.skip("NVRTC not loadable: no run-time specialisation on this machine")
    mech = Mechanism.from_desc(desc, kernel=kernel)
    assert mech.kernel_variant == ("generic" if kernel == KernelMode.GENERIC else "jit:FFRRRPRRRPRRRPXX")
    orc = oracle_of(mech.desc())
    n = 333
    q, v = random_states(desc, n, seed=8, t_jitter=0.2, rpy_jitter=0.6, q_range=0.5)
    tau = np.random.default_rng(8).uniform(-1, 1, size=(n, desc.n_v))
    st = MechanismState(mech, n)
    st.update(q, v)
    vdot, cf = st.dynamics(tau=tau, contact_forces=True)
    vdot_ref, cf_ref = orc.batch_dynamics(q, v, tau)
    assert np.abs(cf_ref).max() > 0.0
    assert rel_err(vdot, vdot_ref) < TOL_DYN
    assert rel_err(cf, cf_ref, floor=1e-6) < TOL_DYN
    ref = fs.Model(mech.desc())
    want = np.array([fs.dynamics(ref, q[e], v[e], tau[e])["vdot"] for e in range(8)])
    assert rel_err(vdot[:8], want) < TOL_DYN
    for integ in (Integrator.SemiImplicitEuler, Integrator.RungeKutta2, Integrator.RungeKutta4):
        st.update(q, v)
        st.step(1e-4, tau=tau, integrator=integ)
        q_ref, v_ref = orc.batch_rollout(q, v, 1e-4, 1, integrator=int(integ), tau=tau)
        assert rel_err(st.q, q_ref) < TOL_STEP and rel_err(st.v, v_ref) < TOL_STEP
    st.update(q, v)
    st.step(1e-4, tau=tau, n_steps=50)
    q1, v1 = st.state()
    assert_rollout_parity(orc, q, v, q1, v1, 1e-4, 50, tau=tau)


def test_sharded_state_from_one_process():
    """ShardedMechanismState: contiguous environment ranges on several batches driven from ONE process
    (here three shards on the same device: the logic, not the hardware). Same result as one batch."""
    from gorilla_physics_b200 import ShardedMechanismState
    mech = models.so101_with_contact()
    desc = mech.desc()
    n = 1000
    q, v = random_states(desc, n, seed=21)
    one = MechanismState(mech, n)
    one.update(q, v)
    one.step(1.0 / 6000.0, n_steps=40)
    q1, v1 = one.state()
    sh = ShardedMechanismState(mech, n, devices=[0, 0, 0])
    assert [hi - lo for lo, hi in sh.ranges] == [333, 333, 334]
    sh.update(q, v)
    sh.step(1.0 / 6000.0, n_steps=40)
    q2, v2 = sh.state()
    assert np.array_equal(q1, q2) and np.array_equal(v1, v2)
    qs, vs = q.copy(), v.copy()
    assert sh.simulate(39.5 / 6000.0, 1.0 / 6000.0, qs, vs) == 40
    assert np.array_equal(qs, q1) and np.array_equal(vs, v1)
    assert not sh.status().any()
    ke, pe, se = sh.energy_sums()
    k1, p1, s1 = one.energies()
    assert abs(ke - k1.sum()) <= 1e-9 * abs(k1.sum()) and abs(pe - p1.sum()) <= 1e-9 * abs(p1.sum())
    # the single-batch form of the diagnostic (no communicator: this batch's own sums)
    d = one.reduce_diagnostics()
    assert abs(d[0] - k1.sum()) <= 1e-9 * abs(k1.sum()) and abs(d[1] - p1.sum()) <= 1e-9 * abs(p1.sum()) and d[3] == 0
    # per-environment torques reach the right shard and row
    tau = np.random.default_rng(3).uniform(-0.3, 0.3, size=(n, desc.n_v))
    one.update(q, v)
    one.step(1.0 / 6000.0, tau=tau, n_steps=25)
    sh.update(q, v)
    sh.step(1.0 / 6000.0, tau=tau, n_steps=25)
    assert np.array_equal(one.state()[0], sh.state()[0]) and np.array_equal(one.state()[1], sh.state()[1])
    qs, vs = q.copy(), v.copy()
    sh.simulate(24.5 / 6000.0, 1.0 / 6000.0, qs, vs, tau=tau)
    assert np.array_equal(qs, one.state()[0]) and np.array_equal(vs, one.state()[1])
    # the shards remain plain batches
    assert sum(s.n_envs for s in sh.shards) == n and sh.shards[1].state()[0].shape == (333, desc.n_q)


_NCCL_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from gorilla_physics_b200 import Communicator, MechanismState, shard_range
from gorilla_physics_b200.workloads import so101_with_contact
from tests.test_parity_gpu import random_states
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo", rank=rank, world_size=world)   # only carries the 128-byte id
ident = [Communicator.unique_id() if rank == 0 else None]
dist.broadcast_object_list(ident, src=0)
comm = Communicator(rank, world, ident[0], rank)
mech = so101_with_contact(); desc = mech.desc()
n = 4001
q, v = random_states(desc, n, seed=9)
lo, hi = shard_range(n, rank, world)
st = MechanismState(mech, hi - lo, device=rank)
st.update(q[lo:hi], v[lo:hi])
st.step(1.0 / 6000.0, n_steps=64)        # no communication on the step path
total = st.reduce_diagnostics(comm)      # the only exchange: NCCL all-reduce of 4 doubles inside the library
mine = st.reduce_diagnostics()
parts = [None] * world
dist.all_gather_object(parts, mine.tolist())
want = np.sum(np.asarray(parts), axis=0)
assert np.allclose(total, want, rtol=1e-12, atol=0), (total, want)
if rank == 0:
    one = MechanismState(mech, n, device=0)
    one.update(q, v); one.step(1.0 / 6000.0, n_steps=64)
    ke, pe, se = one.energies()
    assert abs(total[0] - ke.sum()) <= 1e-9 * abs(ke.sum()) and abs(total[1] - pe.sum()) <= 1e-9 * abs(pe.sum())
    print("nccl-ok", total.tolist())
comm.close()
dist.destroy_process_group()
"""


def test_diagnostic_reduction_over_nccl_two_ranks(tmp_path):
    """One rank per GPU: states sharded, no exchange on the step path, and the end-of-rollout diagnostic
    all-reduced by the library itself (gp_comm_* + gp_batch_reduce_diagnostics, NCCL through dlopen).
    Needs two GPUs."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    import torch
    from gorilla_physics_b200 import nccl_available
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    if not nccl_available():
        pytest.skip("NCCL not loadable")
    root = str(Path(__file__).resolve().parent.parent)
    script = tmp_path / "worker.py"
    script.write_text(_NCCL_WORKER.format(root=root))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and "nccl-ok" in out.stdout, out.stdout[-1500:] + out.stderr[-3000:]


@pytest.mark.parametrize("name,dt,steps", [("so101_contact", 1.0 / 6000.0, 48), ("navbot_contact", 1.0 / 6000.0, 40),
                                           ("cart_pole", 1e-3, 64)])
def test_torque_sequence_is_the_per_step_control_closure(name, dt, steps):
    """gp_batch_step_tau_sequence: one torque vector per fused step (reference simulate.rs:87-112, control_fn
    before every step). Against (1) set_tau + step(1) repeated on the GPU, bitwise, (2) the oracle stepped one
    step at a time with the same torques, and (3) the device-resident [n_steps][n_v][ld] plane form, bitwise."""
    import torch
    factory, kw, _ = WORKLOADS[name]
    mech = factory()
    desc = mech.desc()
    orc = oracle_of(desc)
    n = 301  # ragged: ld = 320
    q, v = random_states(desc, n, seed=77, v_range=0.3, **kw)
    rng = np.random.default_rng(5)
    tau_seq = rng.uniform(-0.5, 0.5, size=(steps, n, desc.n_v))
    # (1) launch per step
    a = MechanismState(mech, n)
    a.update(q, v)
    for s in range(steps):
        a.step(dt, tau=tau_seq[s], n_steps=1)
    qa, va = a.state()
    # the sequence in one call (host rows, streamed)
    b = MechanismState(mech, n)
    b.update(q, v)
    b.set_tau(np.full((n, desc.n_v), 123.0))  # must be ignored by the sequence and left alone
    b.step_tau_sequence(dt, tau_seq)
    qb, vb = b.state()
    if a.step_lanes == 2:
        # (a small batch of a tree with halves steps as warp pairs, the torque-sequence kernels keep a thread per
        # environment: the same numbers to rounding)
        assert rel_err(qb, qa) < 1e-11 and rel_err(vb, va, floor=1e-3) < 1e-9
    else:
        assert np.array_equal(qa, qb) and np.array_equal(va, vb)
    # (2) oracle
    qo, vo = q.copy(), v.copy()
    for s in range(steps):
        qo, vo = orc.batch_rollout(qo, vo, dt, 1, tau=tau_seq[s])
    assert rel_err(qb, qo) < 1e-9 and rel_err(vb, vo, floor=1e-3) < 1e-7
    # (3) planes on the device
    c = MechanismState(mech, n)
    c.update(q, v)
    planes = np.zeros((steps, desc.n_v, c.ld))
    planes[:, :, :n] = tau_seq.transpose(0, 2, 1)
    dev = torch.from_numpy(planes).cuda()
    torch.cuda.synchronize()
    c.step_tau_sequence_device(dt, dev.data_ptr(), steps)
    qc, vc = c.state()
    assert np.array_equal(qb, qc) and np.array_equal(vb, vc)
    # the batch's own torques were not touched: one more plain step uses them
    a.update(qb, vb)
    b.step(dt, n_steps=1)
    a.step(dt, tau=np.full((n, desc.n_v), 123.0), n_steps=1)
    assert np.array_equal(a.q, b.q)


def test_config1_acrobot_swingup_30s():
    """BASELINE.json config 1 (reference examples/acrobot.rs:14-65): the acrobot from rest, swingup_acrobot
    evaluated before every step (in-kernel here), SemiImplicitEuler, dt = 1e-3, 30 s, every state recorded.
    The reference's check is the energy trace KE + double_pendulum_potential_energy2 (energy.rs:19-26) climbing
    to m g (l + 2 l). Against the oracle's one-step-at-a-time rollout:
      * q, v over the first 2 s (2000 steps): 1e-6 absolute (the swing-up is chaotic: bounded horizon);
      * the energy trace over the whole 30 s: 1e-6 relative to the target energy, or within 100x of the oracle's
        own sensitivity to a 1e-15 perturbation of the initial state where that is larger;
      * the outcome of the example: the energy ends nearer to the target than it started, as in the oracle."""
    m, l = 1.0, 7.0
    mech = models.acrobot()
    assert mech.kernel_variant == "double_pendulum_RR"
    orc = oracle_of(mech)
    steps = 30000
    st = MechanismState(mech, 1)
    q0, v0 = np.zeros((1, 2)), np.zeros((1, 2))
    n, hq, hv = st.simulate(30.0, 1e-3, q0.copy(), v0.copy(), controller=Controller.ACROBOT_SWINGUP, ctrl_params=(m, l),
                            history=True)
    assert n == steps and hq.shape == (steps + 1, 1, 2)
    hq, hv = hq[:, 0], hv[:, 0]
    _, _, rq, rv = orc.rollout(q0[0], v0[0], 1e-3, steps, controller=int(Controller.ACROBOT_SWINGUP), params=(m, l), history=True)
    # (the example starts from q = v = 0, which a relative perturbation leaves alone: perturb the first state
    # reached, by 1e-15 absolute, the size of one rounding error of the O(1) quantities of a step)
    _, _, pq, pv = orc.rollout(rq[1] + 1e-15, rv[1] - 1e-15, 1e-3, steps - 1,
                               controller=int(Controller.ACROBOT_SWINGUP), params=(m, l), history=True)
    assert np.abs(hq[:2001] - rq[:2001]).max() < 1e-6 and np.abs(hv[:2001] - rv[:2001]).max() < 1e-6

    def energy(qq, vv):
        ke = np.array([orc.kinetic_energy(a, b) for a, b in zip(qq[::50], vv[::50])])
        pe = m * 9.81 * (l * np.sin(qq[::50, 0]) + (l * np.sin(qq[::50, 0]) + l * np.sin(qq[::50, 0] + qq[::50, 1])))
        return ke + pe
    target = m * 9.81 * 3.0 * l
    e_gpu, e_ref = energy(hq, hv), energy(rq, rv)
    e_pert = energy(np.vstack([rq[:1], pq]), np.vstack([rv[:1], pv]))
    err = np.abs(e_gpu - e_ref) / target
    sens = np.abs(e_pert - e_ref) / target
    bound = np.maximum(1e-6, 100.0 * np.maximum.accumulate(sens))
    print(f"acrobot 30 s: energy error max {err.max():.2e} (first 2 s {err[:41].max():.2e}), oracle sensitivity max {sens.max():.2e}; "
          f"|q err| at 2 s {np.abs(hq[2000] - rq[2000]).max():.2e}, 10 s {np.abs(hq[10000] - rq[10000]).max():.2e}, "
          f"30 s {np.abs(hq[-1] - rq[-1]).max():.2e}")
    assert (err <= bound).all(), (float(err.max()), float(bound.min()))
    assert err[:41].max() < 1e-6
    assert abs(e_gpu[0]) < 1e-12 and abs(e_gpu[-1] - target) < abs(e_gpu[0] - target)


@pytest.mark.parametrize("name", ["so101_contact", "navbot_contact", "quadruped"])
def test_dynamics_parity_componentwise(name):
    """north_star's 1e-10 is checked per environment against the largest component of vdot (rel_err above), under
    which the small components ride on the largest one. Here every component is held against its OWN magnitude,
    with a floor of 1 % of the environment's largest component (a component that happens to cross zero has no
    relative error to speak of): 1e-8, i.e. each component keeps at least 6 more digits than the floor asks for.
    Contact forces likewise. States: free fall and deep in contact (base dropped into the ground)."""
    factory, kw, _ = WORKLOADS[name]
    mech = factory()
    desc = mech.desc()
    orc = oracle_of(desc)
    n = 2048
    worst = 0.0
    for seed, extra in ((11, {}), (12, dict(q_range=2.5, v_range=3.0))):
        q, v = random_states(desc, n, seed=seed, **{**kw, **extra})
        st = MechanismState(mech, n)
        st.update(q, v)
        vdot, cf = st.dynamics(tau=None, contact_forces=True)
        ref, cf_ref = orc.batch_dynamics(q, v)
        for got, want in ((vdot, ref), (cf.reshape(n, -1), cf_ref.reshape(n, -1))):
            if want.shape[1] == 0:
                continue
            scale = np.maximum(np.abs(want).max(axis=1, keepdims=True), 1e-9)
            comp = np.abs(got - want) / np.maximum(np.abs(want), 1e-2 * scale)
            worst = max(worst, float(comp.max()))
            assert float((np.abs(got - want) / scale).max()) < TOL_DYN
    print(f"{name}: worst componentwise relative error {worst:.2e}")
    assert worst < 1e-8


def test_quadruped_trot_controller_in_kernel():
    """QuadrupedTrottingController (reference control/quadruped_control.rs:10-266) evaluated inside the step kernel
    (Controller.QUADRUPED_TROT, nine values of per-environment state in the batch): the reference's only test of the
    14-dof + 12-contact-point model, quadruped_trot_to_position (:417-478), as fused launches instead of a launch and
    two copies per time step. Against the host-closure path (tests/controllers_ref.py computing the torques on the
    host, one step per launch) over the first 600 steps, in one environment per target (warp-pair mapping) and in a
    batch that fills the GPU (thread per environment), state carried across launches; then the reference's
    acceptance after 3 s: |x - target| < 0.1, |v_x| < 0.3."""
    from tests.controllers_ref import QuadrupedTrottingController, quadruped_initial_state
    mech = models.quadruped_on_ground()
    dt = 1.0 / (60.0 * 50.0)
    targets = [-0.2, -0.6, -1.5, -1.8]
    q0, v0 = quadruped_initial_state()
    # host closure, 600 steps, target -0.2
    host = MechanismState(mech, 1)
    host.update(q0[None], v0[None])
    ctrl = QuadrupedTrottingController(dt, targets[0], -0.8)
    q, v = host.state()
    trace = {}
    for step in range(600):
        host.step(dt, tau=ctrl.control(q[0], v[0])[None])
        q, v = host.state()
        if step + 1 in (1, 10, 100, 250, 600):
            trace[step + 1] = (q[0].copy(), v[0].copy(), [f.copy() for f in ctrl.foot_locations])
    for n_copies in (1, 40000):   # 1: warp pairs (small batch); 40 000: thread per environment, ticket mode
        st = MechanismState(mech, n_copies)
        assert st.step_lanes == (2 if n_copies == 1 else 1)
        st.update(np.tile(q0, (n_copies, 1)), np.tile(v0, (n_copies, 1)))
        done = 0
        for upto in (1, 10, 100, 250, 600):   # uneven launches: the controller state carries over
            st.step(dt, n_steps=upto - done, controller=Controller.QUADRUPED_TROT, ctrl_params=(dt, targets[0], -0.8))
            done = upto
            qg, vg = st.state()
            qh, vh, feet = trace[upto]
            assert np.abs(qg[0] - qh).max() < 1e-9 * upto and np.abs(vg[0] - vh).max() < 1e-8 * upto, (n_copies, upto)
            assert np.array_equal(qg[0], qg[-1])
            cs = st.controller_state(9)
            assert cs[0, 0] == upto + 1 and np.array_equal(cs[0], cs[-1])
            got = cs[0, 1:].reshape(4, 2)
            np.testing.assert_allclose(got, np.array([[f[0], f[2]] for f in feet]), rtol=0, atol=1e-12)
        assert not st.status().any()
    # the reference's acceptance for several targets (the target is a launch parameter), 9000 steps in 9 launches
    for target in targets:
        one = MechanismState(mech, 1)
        one.update(q0[None], v0[None])
        for _ in range(9):
            one.step(dt, n_steps=1000, controller=Controller.QUADRUPED_TROT, ctrl_params=(dt, target, -0.8))
        qf, vf = one.state()
        assert not one.status().any()
        assert abs(qf[0, 4] - target) < 1e-1, (target, qf[0, 4])
        x, y, z, w = qf[0, 0:4]
        Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                       [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                       [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        assert abs((Rq.T @ vf[0, 3:6])[0]) < 3e-1


@pytest.mark.parametrize("name", ["navbot_contact", "quadruped"])
def test_warp_pair_mapping_through_every_entry_point(name):
    """Small batches of the trees with halves run as warp pairs. Everything the step kernel does besides stepping
    must work from both halves: host-buffer simulate() (environment-major staging copies in and out), recorded
    history, per-environment torques, a ragged last block - bitwise equal to resident
    stepping in the same mapping, and within parity of the oracle (thread-per-environment results: to rounding)."""
    factory, kw, _ = WORKLOADS[name]
    mech = factory()
    desc = mech.desc()
    orc = oracle_of(desc)
    n, steps, dt = 333, 24, 1.0 / 6000.0
    q, v = random_states(desc, n, seed=41, **kw)
    tau = np.random.default_rng(41).uniform(-0.02, 0.02, size=(n, desc.n_v))
    a = MechanismState(mech, n)
    assert a.step_lanes == 2
    a.update(q, v)
    a.step(dt, tau=tau, n_steps=steps)
    qa, va = a.state()
    assert_rollout_parity(orc, q, v, qa, va, dt, steps, tau=tau)
    # one step at a time == fused
    b = MechanismState(mech, n)
    b.update(q, v)
    for _ in range(steps):
        b.step(dt, tau=tau, n_steps=1)
    assert np.array_equal(b.q, qa) and np.array_equal(b.v, va)
    # simulate() through host buffers, with history
    c = MechanismState(mech, n)
    nst, hq, hv = c.simulate((steps - 0.5) * dt, dt, q.copy(), v.copy(), tau=tau, history=True)
    assert nst == steps and np.array_equal(hq[0], q) and np.array_equal(hq[-1], qa) and np.array_equal(hv[-1], va)
    b.update(q, v)
    b.step(dt, tau=tau, n_steps=7)
    assert np.array_equal(hq[7], b.q) and np.array_equal(hv[7], b.v)
    nst, q_out, v_out = c.simulate((steps - 0.5) * dt, dt, q.copy(), v.copy(), tau=tau)
    assert np.array_equal(q_out, qa) and np.array_equal(v_out, va)
    # against the thread-per-environment mapping (a batch too large for pairs holding the same states)
    big_n = 20000
    reps = -(-big_n // n)
    big = MechanismState(mech, big_n)
    assert big.step_lanes == 1
    big.update(np.tile(q, (reps, 1))[:big_n], np.tile(v, (reps, 1))[:big_n])
    big.step(dt, tau=np.tile(tau, (reps, 1))[:big_n], n_steps=steps)
    qb, vb = big.state()
    ok = np.isfinite(qa).all(axis=1) & np.isfinite(qb[:n]).all(axis=1)
    eq = rollout_errors(qb[:n][ok], qa[ok])
    assert ok.mean() > 0.97 and np.median(eq) < 1e-12 and np.quantile(eq, 0.9) < 1e-9
