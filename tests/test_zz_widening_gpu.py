"""GPU parity of the reference's cuboid-built trees - build_biped (builders/biped_builder.rs:12-187: floating base +
two 6-joint legs, 13 bodies, 18 dof, 16 foot-corner contact points: the largest tree the reference builds),
build_leg and build_leg_from_foot (builders/leg_builder.rs:8-211, 24 contact points) - on the ground z = 0.

None of them has a shipped kernel specialisation: they run on kernels compiled at run time for their own topology
(gp_jit.cpp), the biped in the warp-pair mapping at these batch sizes (two legs = two halves), and on the
run-time-topology kernel when forced to. Same bars as tests/test_parity_gpu.py: vdot / contact force / one step 1e-10
relative against the oracle; rollouts within the oracle's own sensitivity. Also here: the examples/ programs against
the oracle's numbers, and the reference's LQR tests through the host-closure path. (This file sorts last on purpose:
it was added after the round's GPU budget was spent, developed against the oracle on the host build of the device
code, tests/test_device_code_on_host.py, and has not run on a B200 yet.)
"""
import math

import numpy as np
import pytest

from gorilla_physics_b200 import Integrator, KernelMode, Mechanism, MechanismState, jit_available
from gorilla_physics_b200.workloads import biped_on_ground, biped_standing_pose
from tests.models import oracle_of
from tests.test_parity_gpu import TOL_DYN, TOL_STEP, assert_rollout_parity, random_states, rel_err

# (a test of this file that has not returned after ten minutes ends the pytest process - method "thread" works even
# while the main thread sits in a CUDA call - instead of holding the GPU box until the driver's own limit)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]

STATES = {"biped": dict(base_t=(0, 0, 0.6), t_jitter=0.1, rpy_jitter=0.3, q_range=0.5),
          "leg": dict(base_t=(0, 0, 0.6), t_jitter=0.1, rpy_jitter=0.3, q_range=0.5),
          "leg_from_foot": dict(base_t=(0, 0, 0.02), t_jitter=0.05, rpy_jitter=0.3, q_range=0.5)}
VARIANT = {"biped": "jit:FRRRRRRRRRRRR", "leg": "jit:FRRRRR", "leg_from_foot": "jit:FRRRRR"}


def on_ground(name, kernel=KernelMode.AUTO):
    m = Mechanism.from_model(name)
    m.add_halfspace((0, 0, 1), 0.0)
    if kernel != KernelMode.AUTO:
        m.set_kernel_mode(kernel)
    return m


def flavoured(name, flavour):
    if flavour == "generic":
        m = on_ground(name, KernelMode.GENERIC)
        assert m.kernel_variant == "generic"
        return m
    if not jit_available():
        pytest.skip("NVRTC not loadable: no run-time specialisation on this machine")
    m = on_ground(name)
    assert m.kernel_variant == VARIANT[name]  # what AUTO picks for an unlisted tree
    return m


@pytest.mark.parametrize("name", list(STATES))
@pytest.mark.parametrize("flavour", ["jit", "generic"])
def test_cuboid_model_dynamics_parity(name, flavour):
    mech = flavoured(name, flavour)
    desc = on_ground(name).desc()
    orc = oracle_of(desc)
    n = 512
    q, v = random_states(desc, n, seed=1234, **STATES[name])
    tau = np.random.default_rng(7).uniform(-1.0, 1.0, size=(n, desc.n_v))
    st = MechanismState(mech, n)
    st.update(q, v)
    vdot, cf = st.dynamics(tau=tau, contact_forces=True)
    vdot_ref, cf_ref = orc.batch_dynamics(q, v, tau)
    assert np.abs(cf_ref).max() > 0.0, "never touches the ground: contact path untested"
    assert rel_err(vdot, vdot_ref) < TOL_DYN
    assert rel_err(cf, cf_ref) < TOL_DYN
    vdot0 = st.dynamics(tau=None)
    assert rel_err(vdot0, orc.batch_dynamics(q, v, None)[0]) < TOL_DYN
    assert not st.status().any()


def test_biped_mass_matrix_energy_and_poses():
    mech = flavoured("biped", "jit")
    desc = mech.desc()
    orc = oracle_of(desc)
    n = 32
    q, v = random_states(desc, n, seed=99, **STATES["biped"])
    st = MechanismState(mech, n)
    st.update(q, v)
    M, c = st.mass_matrix()
    ke, pe, se = st.energies()
    poses = st.poses()
    for e in range(n):
        ref = orc.dynamics(q[e], v[e], None, want="all")
        assert M[e].shape == (18, 18)
        assert rel_err(M[e][None], ref["mass_matrix"][None]) < 1e-12
        assert rel_err(c[e][None], ref["bias"][None]) < TOL_DYN
        assert abs(ke[e] - orc.kinetic_energy(q[e], v[e])) <= 1e-11 * max(1.0, abs(ke[e]))
        assert abs(pe[e] - orc.gravitational_energy(q[e])) <= 1e-11 * max(1.0, abs(pe[e]))
        np.testing.assert_allclose(poses[e], orc.poses(q[e]), rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", list(STATES))
@pytest.mark.parametrize("integrator", [Integrator.SemiImplicitEuler, Integrator.RungeKutta2, Integrator.RungeKutta4])
def test_cuboid_model_single_step_parity(name, integrator):
    mech = flavoured(name, "jit")
    desc = mech.desc()
    orc = oracle_of(desc)
    n = 256
    q, v = random_states(desc, n, seed=4321, **STATES[name])
    dt = 1.0 / 6000.0
    st = MechanismState(mech, n)
    st.update(q, v)
    st.step(dt, tau=None, integrator=integrator)
    q1, v1 = st.state()
    q_ref, v_ref = orc.batch_rollout(q, v, dt, 1, integrator=int(integrator))
    assert rel_err(q1, q_ref) < TOL_STEP
    assert rel_err(v1, v_ref) < TOL_STEP


def test_biped_settles_from_the_standing_pose_of_the_reference():
    """interface/biped.rs:16-57 createBiped: knees bent, feet on the ground; 400 fused steps at dt = 1/6000 without
    torques (the legs start to fold), against the oracle stepping one at a time; fused == unfused bitwise."""
    mech = flavoured("biped", "jit")
    desc = mech.desc()
    orc = oracle_of(desc)
    n = 96
    rng = np.random.default_rng(5)
    q = np.tile(biped_standing_pose(), (n, 1))
    q[:, 7:] += rng.uniform(-0.02, 0.02, size=(n, 12))
    q[:, 4:6] += rng.uniform(-0.01, 0.01, size=(n, 2))
    v = np.zeros((n, desc.n_v))
    # the pose of the reference: both feet level, their centres at z = 0 (soles 0.025 inside the ground)
    foot = np.asarray(orc.poses(biped_standing_pose()))[-1]
    assert abs(foot[6]) < 1e-12 and abs(foot[3] - 1.0) < 1e-12
    _, cf0 = orc.batch_dynamics(q, v, None)
    assert (np.abs(cf0).reshape(n, -1).max(axis=1) > 0).all()
    dt, steps = 1.0 / 6000.0, 400
    st = MechanismState(mech, n)
    st.update(q, v)
    st.step(dt, n_steps=steps)
    q1, v1 = st.state()
    assert_rollout_parity(orc, q, v, q1, v1, dt, steps, integrator=0)
    st2 = MechanismState(mech, n)
    st2.update(q, v)
    for _ in range(10):
        st2.step(dt, n_steps=1)
    st3 = MechanismState(mech, n)
    st3.update(q, v)
    st3.step(dt, n_steps=10)
    np.testing.assert_array_equal(st2.q, st3.q)
    np.testing.assert_array_equal(st2.v, st3.v)


def test_biped_workload_of_the_package():
    """gorilla_physics_b200.WORKLOADS['biped'] (bench.py --workload biped): device-randomised states, one launch of 64
    fused steps, against the oracle on the first environments."""
    from gorilla_physics_b200 import WORKLOADS
    w = WORKLOADS["biped"]
    mech = w.mechanism()
    if not jit_available():
        pytest.skip("NVRTC not loadable")
    assert mech.kernel_variant == VARIANT["biped"] and biped_on_ground().desc().n_v == 18
    orc = oracle_of(mech.desc())
    n = 4096
    st = MechanismState(mech, n)
    st.randomize(seed=3, **w.randomize)
    q0, v0 = st.state()
    st.step(w.dt, n_steps=64)
    q1, v1 = st.state()
    k = 128
    assert_rollout_parity(orc, q0[:k], v0[:k], q1[:k], v1[:k], w.dt, 64, integrator=0)
    assert math.isfinite(float(np.abs(q1).max())) and not st.status().any()


def test_examples_run_on_the_gpu(tmp_path):
    """examples/*.cpp (the reference's examples/rimless_wheel.rs, examples/cube.rs and interface/biped.rs createBiped
    against the C++ facade) run to the end on the GPU and print what the oracle gets for the same runs: the rimless
    wheel's limit cycle (oracle: pitch rate 0.302 ... 0.726 over the last 2 s of 20), the cube at rest on the ground
    (oracle: x = 0.322, z = -0.504; a tumbling cube amplifies a 1e-15 perturbation to 2e-4 in 5 s, hence loose bounds), the
    biped 0.1 s after it was set down (oracle: base z 0.47817, feet z 0.047587, kinetic energy 0.35401; sensitivity 4e-15)."""
    import re
    import shutil
    import subprocess
    from pathlib import Path
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    root = Path(__file__).resolve().parent.parent
    lib_dir = root / "gorilla_physics_b200" / "lib"
    out = {}
    for name in ("rimless_wheel", "cube", "biped"):
        exe = tmp_path / name
        subprocess.run(["g++", "-std=c++17", "-O1", "-I", str(root / "include"), str(root / "examples" / f"{name}.cpp"), "-o", str(exe),
                        "-L", str(lib_dir), "-lgorilla_b200", f"-Wl,-rpath,{lib_dir}", "-ldl", "-lpthread", "-lrt"], check=True)
        r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (name, r.stdout[-500:], r.stderr[-500:])
        out[name] = r.stdout
    import sys
    r = subprocess.run([sys.executable, str(root / "examples" / "batched_so101.py"), "8192", "3"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "env-steps/s" in r.stdout and " 0 flagged environments" in r.stdout, (r.stdout[-500:], r.stderr[-500:])
    num = r"(-?[\d.]+(?:e[-+]?\d+)?)"
    lo, hi = map(float, re.search(rf"last 2 s: {num} \.\.\. {num}", out["rimless_wheel"]).groups())
    assert "kernel: floating_F" in out["rimless_wheel"] and 0.25 < lo < 0.36 and 0.65 < hi < 0.80, out["rimless_wheel"]
    x, z = map(float, re.search(rf"x = {num}, z = {num}", out["cube"]).groups())
    speed = float(re.search(rf"final speed {num}", out["cube"]).group(1))
    assert "25001 states" in out["cube"] and 0.2 < x < 0.45 and abs(z + 0.5) < 0.02 and speed < 0.05, out["cube"]
    assert "kernel: jit:FRRRRRRRRRRRR" in out["biped"] or not jit_available(), out["biped"]
    assert abs(float(re.search(rf"base lifted by {num}", out["biped"]).group(1)) - 0.532843) < 1e-5
    bz, lz, rz, ke = map(float, re.search(rf"base z = {num}, left foot z = {num}, right foot z = {num}, kinetic energy {num}",
                                          out["biped"]).groups())
    assert abs(bz - 0.478170) < 1e-5 and abs(lz - 0.0475871) < 1e-6 and abs(rz - 0.0475871) < 1e-6 and abs(ke - 0.354009) < 1e-5


def test_reference_lqr_tests_through_the_host_closure_path():
    """control/lqr.rs:78-186 acrobot_lqr / cart_pole_lqr on the GPU: the controller is a host closure (set_tau + one
    Runge-Kutta-4 step per call, like the reference's simulate()), several perturbed starts at once; the reference's
    own start must end within the reference's tolerances, and the first 200 steps agree with the oracle."""
    from tests import models
    PI = math.pi
    cases = [
        # (description, K, operating point (q), actuated dof, start q, start v, final time, tolerance on q, perturbation of the
        #  other starts: the acrobot's LQR has a narrow basin - 0.01 rad more already diverges, in the oracle as well)
        (models.double_pendulum_hanging(5.0, 7.0), np.array([-14067.26123453, -4689.08739542, -15265.74479887, -5803.13757768]),
         np.array([-PI, 0.0]), 1, np.array([-PI - 0.03, 0.03]), np.array([0.03, 0.03]), 100.0, 1e-3, 0.002),
        (models.cart_pole(3.0, 1.0, 5.0, 7.0, (0, -1, 0)), np.array([-1.0, 210.06025784, -5.27501096, 129.54543534]),
         np.array([0.0, PI]), 0, np.array([-1.0, PI + 0.5]), np.array([1.0, 0.5]), 50.0, 2e-3, 0.01),
    ]
    for desc, K, q_op, actuated, q0, v0, final_time, tol_q, amp in cases:
        mech = Mechanism.from_desc(desc)
        orc = oracle_of(desc)
        n = 8
        rng = np.random.default_rng(1)
        q = np.tile(q0, (n, 1))
        v = np.tile(v0, (n, 1))
        q[1:] += rng.uniform(-amp, amp, size=(n - 1, 2))  # environment 0 is the reference's own start
        st = MechanismState(mech, n)
        st.update(q, v)
        qo, vo = q0.copy(), v0.copy()
        dt = 0.01
        from oracle.binding import simulate_step_count
        for step in range(simulate_step_count(final_time, dt)):
            x = np.concatenate([q - q_op, v], axis=1)
            tau = np.zeros((n, 2))
            tau[:, actuated] = -(x @ K)
            st.step(dt, tau=tau, integrator=Integrator.RungeKutta4)
            q, v = st.state()
            if step < 200:
                to = np.zeros(2)
                to[actuated] = -float(K @ np.concatenate([qo - q_op, vo]))
                qo, vo = orc.step(qo, vo, to, dt, 2)
                assert np.abs(q[0] - qo).max() < 1e-9 and np.abs(v[0] - vo).max() < 1e-9, step
        assert not st.status().any()
        np.testing.assert_allclose(q[0], q_op, atol=tol_q)
        np.testing.assert_allclose(v[0], [0.0, 0.0], atol=1e-3)
        np.testing.assert_allclose(q, np.tile(q_op, (n, 1)), atol=5 * tol_q)  # the perturbed starts settle as well


def test_reference_slip_position_control_on_the_gpu():
    """control/SLIP_control.rs:78-131 position_control on the GPU: the SpringContact leg, semi-implicit Euler at
    dt = 1/600 for 10 s, the touch-down angle set by the controller on the host at every apex (through the batch's
    spring-contact state, as the reference edits state.bodies[0].spring_contacts[0]); the reference's target x = 2
    and two more at once, each to the reference's 1e-3 (oracle: 2.4e-4, 2.2e-5, 3.7e-5)."""
    from tests.test_oracle_golden import _SLIPController, _world_linear_velocity
    l_rest = 0.2
    mech = Mechanism.from_model("slip", [0.54, 1.0, l_rest, 0.0, 2000.0])
    mech.add_halfspace((0, 0, 1), -0.5)
    assert mech.kernel_variant == "floating_F" and mech.n_spring_contacts == 1
    targets = [2.0, 1.0, -1.5]
    n = len(targets)
    ctrls = [_SLIPController(1.0, 0.5, 20.0, 0.0, 0.0) for _ in targets]
    st = MechanismState(mech, n)
    st.update(np.tile(np.array([0, 0, 0, 1.0, 0, 0, 0]), (n, 1)), np.zeros((n, 6)))
    dt = 1.0 / 600.0
    vz_prev = np.zeros(n)
    apexes = 0
    for _ in range(int(10.0 / dt)):
        st.step(dt)
        q, v = st.state()
        v_lin = np.stack([_world_linear_velocity(q[e], v[e]) for e in range(n)])
        apex = (vz_prev >= 0.0) & (v_lin[:, 2] < 0.0)
        if apex.any():
            sc = st.spring_contact_state()
            for e in np.nonzero(apex)[0]:
                angle = ctrls[e].control_to_pos(q[e], v[e], targets[e])
                d = np.array([math.sin(angle), 0.0, -math.cos(angle)])
                sc[e, 0, 4:7] = d / np.linalg.norm(d)
                sc[e, 0, 7] = l_rest
                apexes += 1
            st.set_spring_contact_state(sc)
        vz_prev = v_lin[:, 2]
    assert apexes > 3 * 10 and not st.status().any()
    for e, t in enumerate(targets):
        assert abs(q[e, 4] - t) < 1e-3, (t, q[e, 4])
