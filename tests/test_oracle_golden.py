"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for the
step path. Each test names the reference test it restates. CPU only.

The reference is Rust and cannot be built here, so these golden vectors (closed forms, numbers from
RigidBodyDynamics.jl quoted in the reference's tests, physical outcomes of its rollouts) are what
anchors the oracle; the CUDA path is then compared with the oracle in test_parity_gpu.py.
"""
import math

import numpy as np
import pytest

from gorilla_physics_b200 import (FLOATING, PRISMATIC, REVOLUTE, Mechanism, MechanismDesc, iso, quat_from_euler,
                                  quat_from_scaled_axis)
from oracle.binding import (OracleMechanism, quat_from_euler as o_quat_from_euler, simple_double_pendulum,
                            simulate_step_count, twist_transform)
from tests import models
from tests.models import GRAVITY, PI, oracle_of, pose_q

SIE, RK2, RK4 = 0, 1, 2


# ---------------------------------------------------------------- dynamics.rs known answers
def test_dynamics_rod_pendulum_horizontal():  # dynamics.rs:883 (assert_eq!, exact)
    o = oracle_of(models.rod_pendulum())
    assert o.dynamics([0.0], [0.0], [0.0])[0] == 3.0 * GRAVITY / (2.0 * 7.0)


def test_dynamics_rod_pendulum_horizontal_rotated_frame():  # dynamics.rs:916 (exact)
    rot = iso((0, 0, 0), quat_from_scaled_axis((PI / 2.0, 0, 0)))
    o = oracle_of(models.rod_pendulum(rod_to_world=rot, axis=(0, 0, 1)))
    assert o.dynamics([0.0], [0.0], [0.0])[0] == -3.0 * GRAVITY / (2.0 * 7.0)


def test_dynamics_rod_pendulum_horizontal_moved_frame():  # dynamics.rs:949 (1e-6)
    rot = iso((11.0, 0, 0), quat_from_scaled_axis((PI / 2.0, 0, 0)))
    o = oracle_of(models.rod_pendulum(rod_to_world=rot, axis=(0, 0, 1)))
    assert abs(o.dynamics([0.0], [0.0], [0.0])[0] - (-3.0 * GRAVITY / (2.0 * 7.0))) < 1e-6


def test_dynamics_hold_horizontal_rod_pendulum():  # dynamics.rs:979 (1e-6)
    m, l = 5.0, 7.0
    o = oracle_of(models.rod_pendulum(m, l))
    assert abs(o.dynamics([0.0], [0.0], [-m * GRAVITY * l / 2.0])[0]) < 1e-6


def test_dynamics_simple_pendulum_horizontal():  # dynamics.rs:1011 (exact)
    o = oracle_of(models.rod_pendulum(point_mass=True))
    assert o.dynamics([0.0], [0.0], [0.0])[0] == GRAVITY / 7.0


def test_dynamics_double_pendulum_horizontal():  # dynamics.rs:1038 (1e-6)
    o = oracle_of(models.double_pendulum_horizontal())
    np.testing.assert_allclose(o.dynamics([0, 0], [0, 0], [0, 0]), [GRAVITY / 7.0, -GRAVITY / 7.0], atol=1e-6)


def test_double_pendulum_dynamics_vs_closed_form():  # dynamics.rs:1091 (1e-5)
    m, l = 3.0, 5.0
    o = oracle_of(models.double_pendulum_hanging(m, l))
    vdot = o.dynamics([3.0, 5.0], [3.0, 5.0], [0.0, 0.0])
    np.testing.assert_allclose(vdot, simple_double_pendulum(m, m, l, l, 3.0, 5.0, 3.0, 5.0), atol=1e-5)


def test_simple_double_pendulum_rbdjl_numbers():  # double_pendulum.rs:64-83 (1e-4)
    np.testing.assert_allclose(simple_double_pendulum(1.0, 3.0, 2.0, 4.0, 1.0, 2.0, 3.0, 4.0), [68.8824, -58.9877],
                               atol=1e-4)


def test_ball_dynamics_rbdjl_numbers():  # joint/floating.rs:82-129 (1e-5): body-frame velocity convention
    o = oracle_of(models.ball())
    vdot = o.dynamics(pose_q((0.1, 0.2, 0.3), (1.0, 2.0, 3.0)), [1, 2, 3, 4, 5, 6])
    np.testing.assert_allclose(vdot, [0.0, 0.0, 0.0, 4.948946, -6.959844, -6.566419], atol=1e-5)


def test_motor_turning_mass():  # dynamics.rs:1194-1251 (1e-5)
    o = oracle_of(models.motor_turning_mass())
    q, v = models.motor_turning_mass().zero_state()
    acc = o.dynamics(q, v, [0, 0, 0, 0, 0, 0, 1.0])
    base_angular = -1.0 / (2.0 / 5.0)
    base_linear_x = -1.0
    np.testing.assert_allclose(acc[0:3], [0, 0, base_angular], atol=1e-5)
    np.testing.assert_allclose(acc[3:6], [base_linear_x, 0, -GRAVITY], atol=1e-5)
    assert abs(acc[6] - (-base_angular + 1.0 + -base_linear_x)) < 1e-5


# ---------------------------------------------------------------- mechanism.rs / spatial tests
def test_mass_matrix_structure():  # mechanism.rs:711-786
    d = models.mass_matrix_fixture()
    o = oracle_of(d)
    q, v = d.zero_state()
    M = o.dynamics(q, v, want="all")["mass_matrix"]
    assert M.shape == (8, 8)
    assert np.any(M[6, 0:6] != 0) and M[6, 6] != 0
    assert np.any(M[7, 0:6] != 0) and M[7, 6] != 0 and M[7, 7] != 0
    np.testing.assert_array_equal(M, M.T)


def test_supports():  # mechanism.rs:793-832
    S = oracle_of(models.supports_fixture()).supports()
    sets = [set(int(i) + 1 for i in np.nonzero(row)[0]) for row in S]
    assert sets == [{1, 2, 3, 4, 5}, {2, 3}, {3}, {4, 5}, {5}]
    # the product's host code derives the same table
    S2 = Mechanism.from_desc(models.supports_fixture()).supports()
    np.testing.assert_array_equal(S, S2)


def test_compute_bodies_to_root():  # spatial/transform.rs tests: same 5-body tree, translations compose
    d = models.supports_fixture()
    o = oracle_of(d)
    q, _ = d.zero_state()
    poses = o.poses(q)
    np.testing.assert_array_equal(poses[:, 4:], [[0, 0, 0], [-1, 0, 0], [-1, 0, 1], [1, 0, 0], [1, 0, 1]])
    np.testing.assert_array_equal(poses[:, :4], np.tile([0, 0, 0, 1.0], (5, 1)))


def test_twist_transform():  # spatial/twist.rs:216-284 (1e-6): w' = R w, v' = R v + t x w'
    t = twist_transform(iso((0, 0, 0), quat_from_scaled_axis((0, 0, PI / 2.0))), [1, 0, 0, 0, 0, 0])
    np.testing.assert_allclose(t, [0, 1, 0, 0, 0, 0], atol=1e-6)
    t = twist_transform(iso((1.0, 0, 0), (0, 0, 0, 1)), [0, 0, 1.0, 0, 0, 0])
    np.testing.assert_allclose(t, [0, 0, 1, 0, -1, 0], atol=1e-6)  # (1,0,0) x (0,0,1) = (0,-1,0)


def test_nalgebra_euler_convention():  # R = Rz(yaw) Ry(pitch) Rx(roll)
    q = o_quat_from_euler(0.1, 0.2, 0.3)
    np.testing.assert_allclose(q, quat_from_euler(0.1, 0.2, 0.3), atol=0)
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])

    def rx(a): return np.array([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]])
    def ry(a): return np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
    def rz(a): return np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]])
    np.testing.assert_allclose(R, rz(0.3) @ ry(0.2) @ rx(0.1), atol=1e-15)


def test_simulate_step_count_follows_f64_loop():  # simulate.rs:97-109
    assert simulate_step_count(1.0, 0.1) == 11  # 0.1 * 10 accumulates to 0.9999999999999999 < 1.0
    assert simulate_step_count(2.0, 1e-3) in (2000, 2001)
    assert simulate_step_count(0.0, 1e-3) == 0


# ---------------------------------------------------------------- rollouts: simulate.rs / energy.rs
def test_simulate_horizontal_right_rod():  # simulate.rs:129-174 (SemiImplicitEuler)
    m, l = 5.0, 7.0
    o = oracle_of(models.rod_pendulum(m, l))
    q, v, hq, hv = o.simulate([0.0], [0.0], 10.0, 0.001, SIE, tau=[0.0])
    assert abs(hq[:, 0].max() - PI) < 1e-2
    pe = m * GRAVITY * l / 2.0 * (-math.sin(q[0]))
    ke = 0.5 * (m * l * l / 3.0) * v[0] * v[0]
    assert abs(pe + ke) < 1e-1


def test_simulate_cart():  # simulate.rs:177-214 (RK4)
    o = oracle_of(Mechanism.from_model("cart"))
    q, v, hq, hv = o.simulate([2.0], [-1.0], 10.0, 0.02, RK4, tau=[0.0])
    assert abs(q[0] - (2.0 - 1.0 * 10.0)) < 2e-2 and abs(v[0] + 1.0) < 1e-6


def test_simulate_cart_pole_steady_state():  # simulate.rs:218-272 (RK4, 1e-5)
    o = oracle_of(models.cart_pole(3.0, 1.0, 5.0, 7.0, (0, 1, 0)))
    F = 1.0
    acc = F / (3.0 + 5.0)
    theta = math.atan2(acc, GRAVITY)
    q, v = o.simulate([0.0, theta], [0.0, 0.0], 20.0, 1e-2, RK4, tau=[F, 0.0], history=False)
    assert abs(v[0] - acc * 20.0) < 1e-4 + 1e-2 * acc  # step count follows the f64 loop
    assert abs(q[1] - theta) < 1e-5


def test_double_pendulum_energy():  # energy.rs:82-127 (SemiImplicitEuler, 1e-1)
    m, l = 1.0, 1.0
    o = oracle_of(models.double_pendulum_hanging(m, l))

    def energy(q, v):
        h1 = -l * math.cos(q[0])
        h2 = -l * math.cos(q[0]) - l * math.cos(q[0] + q[1])
        return o.kinetic_energy(q, v) + m * GRAVITY * (h1 + h2)
    q0, v0 = [1.0, 1.0], [0.1, 0.1]
    q, v = o.simulate(q0, v0, 2.0, 1e-3, SIE, tau=[0.0, 0.0], history=False)
    assert abs(energy(q, v) - energy(q0, v0)) < 1e-1


def test_cart_pole_energy():  # energy.rs:130-178 (RK4, 1e-1)
    o = oracle_of(models.cart_pole(1.0, 1.0, 1.0, 1.0, (0, -1, 0)))

    def energy(q, v):
        return o.kinetic_energy(q, v) - 1.0 * GRAVITY * 1.0 * math.cos(q[1])
    q0, v0 = [0.0, PI + 0.1], [0.0, 0.0]
    q, v = o.simulate(q0, v0, 2.0, 0.001, RK4, tau=[0.0, 0.0], history=False)
    assert abs(energy(q, v) - energy(q0, v0)) < 1e-1


def test_spring_on_frictionless_ground():  # dynamics.rs:1133-1190 (RK4, midpoint conserved 1e-5)
    d, l_init = models.spring_pair()
    o = oracle_of(d)
    q, v = d.zero_state()
    q[7] = l_init
    q, v = o.simulate(q, v, 1.1, 1e-3, RK4, history=False)
    poses = o.poses(q)
    x_a, x_b = poses[0, 4], poses[1, 4]
    assert abs((x_a + x_b) / 2.0 - l_init / 2.0) < 1e-5
    assert x_a > 0.0 and x_b < l_init


# ---------------------------------------------------------------- contact.rs rollouts
def test_pendulum_hit_ground():  # contact.rs:371-405 (RK4, rests at 30 degrees, 1e-3)
    m, l = 1.5, 10.0
    d = models.rod_pendulum(m, l, point_mass=True)
    d.add_contact_point(1, (l, 0, 0))
    d.add_halfspace((0, 0, 1), -5.0)
    q, v = oracle_of(d).simulate([0.0], [0.0], 5.0, 1e-2, RK4, tau=[0.0], history=False)
    assert abs(q[0] - 30.0 * PI / 180.0) < 1e-3


def _cube_on(h_ground, alpha=0.9, mu=0.5):
    m = Mechanism.from_model("cube")
    m.add_halfspace((0, 0, 1), h_ground, alpha=alpha, mu=mu)
    return oracle_of(m)


def test_cube_fall_ground():  # contact.rs:408-441
    o = _cube_on(-10.0)
    q, v = o.simulate(pose_q(), [1, 1, 1, 1, 1, 1], 5.0, 1e-3, RK4, history=False)
    assert abs(q[6] - (-10.0 + 0.5)) < 1e-2


def test_cube_slide_ground():  # contact.rs:444-490: friction stopping distance
    mu = 0.5
    o = _cube_on(-0.5, alpha=1.0, mu=mu)
    q, v = o.simulate(pose_q(), [0, 0, 0, 1.0, 0, 0], 2.0, 1e-3, RK4, history=False)
    assert abs(q[6]) < 1e-2
    assert np.linalg.norm(v[3:6]) < 5e-3 and np.linalg.norm(v[0:3]) < 1e-2
    acc = -GRAVITY * mu
    t_slide = 1.0 / -acc
    assert abs(q[4] - (t_slide + acc * t_slide ** 2 / 2.0)) < 1e-2


def test_cube_hit_ground():  # contact.rs:493-530
    o = _cube_on(-10.0)
    q, v = o.simulate(pose_q(), [0, 5.0, 0, 1.0, 0, 0], 5.0, 1e-3, RK4, history=False)
    assert abs(q[6] - (-10.0 + 0.5)) < 1e-2
    assert np.linalg.norm(v[3:6]) < 5e-3 and np.linalg.norm(v[0:3]) < 1e-2


def test_two_cubes_hit_ground():  # contact.rs:533-604: two floating bodies on the world
    m, l = 3.0, 1.0
    d = MechanismDesc()
    for _ in range(2):
        d.add_body(0, FLOATING, moment=np.eye(3) * m * l * l / 6.0, mass=m)
    h = l / 2.0
    for body in (1, 2):
        for sz in (-h, h):
            for sx, sy in ((h, h), (h, -h), (-h, h), (-h, -h)):
                d.add_contact_point(body, (sx, sy, sz))
    d.add_halfspace((0, 0, 1), -10.0)
    o = oracle_of(d)
    q0 = np.concatenate([pose_q(t=(2.0 * l, 0, 0)), pose_q(t=(-2.0 * l, 0, 0))])
    v0 = [0, 5.0, 0, 1.0, 0, 0, 0, -5.0, 0, -1.0, 0, 0]
    q, v = o.simulate(q0, v0, 5.0, 1e-3, RK4, history=False)
    for b in range(2):
        assert abs(q[7 * b + 6] - (-10.0 + l / 2.0)) < 1e-2
        assert np.linalg.norm(v[6 * b + 3:6 * b + 6]) < 5e-3 and np.linalg.norm(v[6 * b:6 * b + 3]) < 1e-2


def test_mass_hit_ground():  # contact.rs:607-676: impact + sliding distance
    m, r, vx, h_ground, mu = 1.0, 0.1, 2.0, -0.3, 0.5
    d = models.ball(m, r)
    d.add_contact_point(1, (0, 0, 0))
    d.add_halfspace((0, 0, 1), h_ground, alpha=1.0, mu=mu)
    q, v = oracle_of(d).simulate(pose_q(), [0, 0, 0, vx, 0, 0], 2.0, 1e-3, RK4, history=False)
    assert abs(q[6] - h_ground) < 1e-2
    assert np.linalg.norm(v[3:6]) < 1e-2 and np.linalg.norm(v[0:3]) < 1e-3
    t_hit = math.sqrt(-h_ground * 2.0 / GRAVITY)
    vx_after = vx - GRAVITY * t_hit * mu
    assert vx_after > 0.0
    acc = -GRAVITY * mu
    t_slide = vx_after / -acc
    x_expect = t_hit * vx + vx_after * t_slide + acc * t_slide ** 2 / 2.0
    assert abs(q[4] - x_expect) < 1e-2


def test_rimless_wheel_limit_cycle():  # contact.rs:679-729 (RK4, dt=1/600, 20 s)
    o = oracle_of(models.rimless_wheel_on_slope())
    q, v, hq, hv = o.simulate(pose_q(), [0, 0, 0, 1.0, 0, 0], 20.0, 1.0 / 600.0, RK4)
    omega_y = hv[:, 1]
    assert omega_y[-1] > 0.0
    assert omega_y.max() < 0.75


def test_ball_fall():  # joint/floating.rs ball_fall: free fall under gravity (SemiImplicitEuler)
    o = oracle_of(models.ball())
    q, v = o.simulate(pose_q(), np.zeros(6), 1.0, 1e-3, SIE, history=False)
    n = simulate_step_count(1.0, 1e-3)
    assert abs(v[5] + GRAVITY * n * 1e-3) < 1e-9  # world-aligned body frame: v_z = -g t exactly per step
    assert abs(q[6] + 0.5 * GRAVITY * (n * 1e-3) ** 2) < 1e-2


def test_acrobot_swingup():  # control/swingup.rs:129-188 (RK4, dt=1e-2, 50 s, "very relaxed check")
    m, l = 1.0, 7.0
    o = oracle_of(models.double_pendulum_horizontal(m, l, axis=(0, -1, 0)))
    n = simulate_step_count(50.0, 1e-2)
    q, v, hq, hv = o.rollout([-PI / 2.0 + 0.1, 0.0], [0.0, 0.0], 1e-2, n, RK4, controller=2, params=(m, l),
                             history=True)
    q1 = np.mod(hq[:, 0], 2 * PI)
    q2 = np.mod(hq[:, 1], 2 * PI)
    swungup = (np.abs(q1 - PI / 2.0) < 0.5) & (np.abs(q2) < 0.5) & (np.abs(hv[:, 0]) < 0.3) & (np.abs(hv[:, 1]) < 0.3)
    assert swungup.any()


def test_cart_pole_swingup():  # control/swingup.rs:190-261 (RK4, dt=1e-2, 30 s)
    m_cart, m_pole, l_pole = 3.0, 5.0, 7.0
    o = oracle_of(models.cart_pole(m_cart, 1.0, m_pole, l_pole, (0, -1, 0)))
    n = simulate_step_count(30.0, 1e-2)
    q, v, hq, hv = o.rollout([0.0, 0.1], [0.0, 0.0], 1e-2, n, RK4, controller=3, params=(m_cart, m_pole, l_pole),
                             history=True)
    q2 = np.mod(hq[:, 1], 2 * PI)
    swungup = (np.abs(hq[:, 0]) < 1e-1) & (np.abs(q2 - PI) < 1e-1) & (np.abs(hv[:, 0]) < 1e-1) & (np.abs(hv[:, 1]) < 1e-1)
    assert swungup.any()
    assert abs(q[0]) < 1e-1 and abs(v[0]) < 1e-1
    E = o.kinetic_energy(q, v) - m_pole * GRAVITY * l_pole * math.cos(q[1])  # energy.rs:34-41
    assert abs(E - m_pole * l_pole * GRAVITY) < 2.0


def test_acrobot_lqr():  # control/lqr.rs:78-124 (RK4, dt = 1e-2, 100 s): the hard-coded gain holds the upright acrobot, 1e-3
    m, l = 5.0, 7.0
    o = oracle_of(models.double_pendulum_hanging(m, l))
    K = np.array([-14067.26123453, -4689.08739542, -15265.74479887, -5803.13757768])  # lqr.rs:10-15
    q1_upright = -PI
    q, v = np.array([q1_upright - 0.03, 0.03]), np.array([0.03, 0.03])
    for _ in range(simulate_step_count(100.0, 0.01)):
        xbar = np.array([q[0] - q1_upright, q[1], v[0], v[1]])
        q, v = o.step(q, v, [0.0, -float(K @ xbar)], 0.01, RK4)
    np.testing.assert_allclose(q, [q1_upright, 0.0], atol=1e-3)
    np.testing.assert_allclose(v, [0.0, 0.0], atol=1e-3)


def test_cart_pole_lqr():  # control/lqr.rs:126-186 (RK4, dt = 1e-2, 50 s): from 0.5 rad off upright back to it, 2e-3 / 1e-3
    o = oracle_of(models.cart_pole(3.0, 1.0, 5.0, 7.0, (0, -1, 0)))
    K = np.array([-1.0, 210.06025784, -5.27501096, 129.54543534])  # lqr.rs:40-43
    q, v = np.array([-1.0, PI + 0.5]), np.array([1.0, 0.5])
    for _ in range(simulate_step_count(50.0, 0.01)):
        xbar = np.array([q[0], q[1] - PI, v[0], v[1]])
        q, v = o.step(q, v, [-float(K @ xbar), 0.0], 0.01, RK4)
    np.testing.assert_allclose(q, [0.0, PI], atol=2e-3)
    np.testing.assert_allclose(v, [0.0, 0.0], atol=1e-3)


def test_acrobot_example_energy_trace():  # examples/acrobot.rs: config 1 (SemiImplicitEuler, dt=1e-3)
    """The swing-up controller pumps the total energy towards m g (l + 2l) (swingup.rs:20)."""
    m, l = 1.0, 7.0
    o = oracle_of(Mechanism.from_model("double_pendulum"))  # defaults = examples/acrobot.rs:14-34
    q, v, hq, hv = o.rollout([0.0, 0.0], [0.0, 0.0], 1e-3, 30000, SIE, controller=2, params=(m, l), history=True)

    def energy(qq, vv):
        pe = m * GRAVITY * (l * math.sin(qq[0]) + (l * math.sin(qq[0]) + l * math.sin(qq[0] + qq[1])))
        return o.kinetic_energy(qq, vv) + pe
    e0, e1 = energy(hq[0], hv[0]), energy(hq[-1], hv[-1])
    target = m * GRAVITY * 3.0 * l
    assert abs(e0) < 1e-12
    assert abs(e1 - target) < abs(e0 - target)


def test_so101_pd_controller_holds_zero_pose():  # control/so101_control.rs:12-34 (PD + clamp)
    m = Mechanism.from_model("so101")
    o = oracle_of(m)
    tau = o.control(np.full(6, 0.02), np.zeros(6), 1, (1000.0, 0.1, 10.0))
    np.testing.assert_allclose(tau, np.full(6, -10.0))  # 1000 * -0.02 = -20 clamped to -10
    tau = o.control(np.full(6, 0.001), np.full(6, 1.0), 1, (1000.0, 0.1, 10.0))
    np.testing.assert_allclose(tau, np.full(6, -1.1))


def test_hopper_1d_example_hops_to_setpoint():  # examples/1D_hopper.rs + control/energy_control.rs:24-101
    """Hopper1DController (stateful, Raibert-style energy control): the body keeps hopping and its
    apex height settles near h_setpoint = 0 (ground at -20, leg lengths 2 + 10)."""
    m = models.hopper1d_on_ground()
    o = oracle_of(m)
    q, v = m.desc().zero_state()
    _, _, hq, _ = o.rollout(q, v, 1.0 / 500.0, 15000, SIE, controller=4, params=(200.0, 0.0, 2.0, 10.0), history=True)
    z = hq[:, 6]
    apex = [z[i] for i in range(1, len(z) - 1) if z[i] > z[i - 1] and z[i] >= z[i + 1]]
    assert len(apex) >= 6 and np.isfinite(hq).all()
    assert abs(apex[-1] - 0.0) < 0.5 and abs(apex[-1] - apex[-2]) < 0.05
    assert z.min() > -12.5  # never collapses through the ground (-20 + 2 + 10 = -8 at rest)


def _two_mass_spring(mu):
    """contact.rs:908-955 / :991-1033 fixture: floating body + prismatic(-z) foot with a joint spring"""
    d = MechanismDesc()
    d.add_body(0, FLOATING, moment=models.sphere_moment(1.0, 1.0), mass=1.0)
    d.add_body(1, PRISMATIC, axis=(0, 0, -1), init_iso=iso((0, 0, -1.0)), moment=models.sphere_moment(1.0, 1.0),
               mass=1.0, spring=(100.0, 0.0))
    d.add_contact_point(2, (0, 0, 0))
    d.add_halfspace((0, 0, 1), -2.0, alpha=1.0, mu=mu)
    return d


def test_spring_drop():  # contact.rs:908-985 (RK4, 4 s): exact zeros + energy
    d = _two_mass_spring(mu=0.0)
    o = oracle_of(d)
    q, v = d.zero_state()
    q, v = o.simulate(q, v, 4.0, 1e-3, RK4, history=False)
    poses = o.poses(q)
    foot, body = poses[1], poses[0]
    assert foot[4] == 0.0 and foot[5] == 0.0  # assert_eq! in the reference
    z_tol = ((1.0 + 1.0) * GRAVITY / 50e3) ** (2.0 / 3.0)
    assert abs(foot[6] - (-2.0)) < 2.0 * z_tol
    np.testing.assert_array_equal(body[:4], [0.0, 0.0, 0.0, 1.0])  # body upright, exactly
    total = o.kinetic_energy(q, v) + o.gravitational_energy(q) + o.spring_energy(q)
    assert abs(total - (1.0 * GRAVITY * -2.0 + 1.0 * GRAVITY * (-2.0 + 1.0))) < 1.5


def test_spring_forward_drop():  # contact.rs:991-1074 (RK4): body leans forward at the first bottom
    d = _two_mass_spring(mu=0.5)
    o = oracle_of(d)
    q, v = d.zero_state()
    v[3] = 0.5
    prev, had_bottom = 0.0, False
    for _ in range(1500):
        q, v = o.step(q, v, None, 1e-3, RK4)
        v_spring = v[6]
        if prev < 0.0 and v_spring >= 0.0:
            had_bottom = True
            quat = q[0:4]
            assert quat[0] == 0.0 and quat[2] == 0.0 and quat[1] > 0.0 and quat[3] > 0.0  # rotation about +y
            assert q[4] > o.poses(q)[1][4]  # body in front of the foot
            break
        prev = v_spring
    assert had_bottom


def test_compass_gait_standing():  # contact.rs:735-832 (RK4, 2 s): stands still on a 5 degree slope
    m_hip, r_hip, m_leg, l_leg = 10.0, 0.3, 5.0, 1.0
    d = MechanismDesc()
    d.add_body(0, FLOATING, moment=np.eye(3) * (m_hip * r_hip * r_hip * 2.0 / 5.0), mass=m_hip)
    mx = m_leg * (l_leg / 2.0) * (l_leg / 2.0)
    for _ in range(2):
        d.add_body(1, REVOLUTE, axis=(0, -1, 0), moment=np.diag([mx, mx, 0.0]), cross_part=(0, 0, -m_leg * l_leg / 2.0),
                   mass=m_leg)
    d.add_contact_point(2, (0, 0, -l_leg))
    d.add_contact_point(3, (0, 0, -l_leg))
    ang = math.radians(5.0)
    n = np.array([math.sin(ang), 0.0, math.cos(ang)])
    d.add_halfspace(n / np.linalg.norm(n), 0.0, alpha=1.0, mu=1.0)
    q, v = d.zero_state()
    q[6] = l_leg
    q[7], q[8] = math.radians(30.0), math.radians(-30.0)
    q, v = oracle_of(d).simulate(q, v, 2.0, 1e-3, RK4, history=False)
    assert q[4] > 0.0 and q[6] > 0.0
    assert v[0] == 0.0 and v[2] == 0.0 and v[4] == 0.0  # assert_eq! in the reference: planar motion stays planar
    assert abs(v[1]) < 1e-3 and abs(v[3]) < 6e-3 and abs(v[5]) < 3e-2


# ---------------------------------------------------------------- hybrid::Articulated
def test_articulated_step_tests_of_the_second_engine():
    """hybrid/articulated/mod.rs:474-577: Articulated::step = free_velocity + integrate is this path's
    semi-implicit Euler step without contact. cart: v stays; pendulum: swings to pi, energy kept."""
    # cart (:474-501)
    d = MechanismDesc()
    d.add_body(0, PRISMATIC, axis=(1, 0, 0), moment=np.diag([0.01 + 0.01, 1.01, 1.01]) / 12.0, mass=1.0)
    o = oracle_of(d)
    q, v = o.rollout([0.0], [1.0], 1e-3, 2000, SIE)
    assert abs(v[0] - 1.0) < 1e-3 and abs(q[0] - 2.0) < 1e-3
    # pendulum (:504-534): sphere of mass 1, radius 1 at (l, 0, 0), revolute about y
    l = 1.0
    com = np.array([l, 0.0, 0.0])
    moment = np.eye(3) * (2.0 / 5.0) + 1.0 * (com @ com * np.eye(3) - np.outer(com, com))
    d = MechanismDesc()
    d.add_body(0, REVOLUTE, axis=(0, 1, 0), moment=moment, cross_part=com, mass=1.0)
    o = oracle_of(d)
    e0 = o.kinetic_energy([0.0], [0.0]) + 1.0 * GRAVITY * 0.0
    q, v, hq, hv = o.rollout([0.0], [0.0], 1e-3, 2000, SIE, history=True)
    assert abs(hq[:, 0].max() - PI) < 1e-3
    e1 = o.kinetic_energy(q, v) + 1.0 * GRAVITY * (-l * math.sin(q[0]))  # COM height
    assert abs(e1 - e0) < 1e-2
    # free_velocity == v + vdot dt of the MechanismState engine when armature = 0 and gravity is on
    q0, v0 = [0.3], [0.7]
    np.testing.assert_allclose(o.free_velocity(q0, v0, 1e-3), np.asarray(v0) + o.dynamics(q0, v0) * 1e-3, rtol=1e-14)


def test_armature_adds_to_the_joint_diagonal():  # hybrid/articulated/mod.rs:243-247
    m, l, arm = 5.0, 7.0, 3.5
    d = models.rod_pendulum(m, l)
    d._armature[0] = arm
    o = oracle_of(d)
    inertia = m * l * l / 3.0
    torque_g = m * GRAVITY * l / 2.0
    # MechanismState's dynamics_continuous never reads the armature (only hybrid::Articulated does)
    assert abs(o.dynamics([0.0], [0.0])[0] - torque_g / inertia) < 1e-12
    assert abs(o.free_velocity([0.0], [0.0], 0.01)[0] - 0.01 * torque_g / (inertia + arm)) < 1e-12
    assert abs(o.free_velocity([0.0], [0.0], 0.01, gravity_enabled=False)[0]) < 1e-15
    # the product's host code carries the armature through the flat description
    from gorilla_physics_b200 import Mechanism
    assert Mechanism.from_desc(d).desc().armature[0] == arm


def _slip_direction():
    a = math.radians(45.0)
    d = np.array([math.sin(a), 0.0, -math.cos(a)])
    return d / np.linalg.norm(d)


def test_SLIP_hopping():  # contact.rs:836-903: SpringContact (stateful leg), SemiImplicitEuler, 3 s
    m = Mechanism.from_model("slip")  # helpers.rs:308-337, SLIP_hopping's parameters
    m.add_halfspace((0, 0, 1), -0.3)
    o = oracle_of(m)
    direction, l_rest = _slip_direction(), 0.2
    q, v = pose_q(), np.array([0, 0, 0, 5.0, 0, 0])
    st = o.spring_state_init()
    e0 = o.kinetic_energy(q, v) + o.gravitational_energy(q)
    dt = 1.0 / 2000.0
    vz_prev, energies, hs = 0.0, [], []
    for _ in range(int(3.0 / dt)):
        q, v, st, flags = o.step_sc(q, v, st, dt)
        assert flags == 0  # never "Spring force is into the halfspace!"
        vz = v[5]
        if vz_prev > 0.0 and vz <= 0.0:  # apex: swing the leg back to the front angle (the test does this by hand)
            energies.append(o.kinetic_energy(q, v) + o.gravitational_energy(q))
            hs.append(q[6])
            st[0, 4:7] = direction
            st[0, 7] = l_rest
        vz_prev = vz
    assert len(energies) > 0
    assert max(abs(e - e0) for e in energies) < 1e-2
    assert abs(hs[-3] - hs[-2]) < 1e-3 and abs(hs[-2] - hs[-1]) < 1e-3


class _SLIPController:
    """control/SLIP_control.rs:6-61 restated test-side: touch-down angle from a PID on the forward speed"""

    def __init__(self, k_q, speed_bound, k_v_p, k_v_i, k_v_d):
        self.k_q, self.speed_bound, self.k_v_p, self.k_v_i, self.k_v_d = k_q, speed_bound, k_v_p, k_v_i, k_v_d
        self.v_integral, self.v_diff_prev = 0.0, 0.0

    def control_to_pos(self, q, v, x_target):
        v_x_target = -self.k_q * (q[4] - x_target)
        if abs(v_x_target) > self.speed_bound:
            v_x_target = math.copysign(self.speed_bound, v_x_target)
        return self.control_to_velocity(q, v, v_x_target)

    def control_to_velocity(self, q, v, v_x_target):
        v_diff = _world_linear_velocity(q, v)[0] - v_x_target
        self.v_integral += v_diff
        degree = self.k_v_p * v_diff + self.k_v_i * self.v_integral + self.k_v_d * (v_diff - self.v_diff_prev)
        self.v_diff_prev = v_diff
        return math.radians(degree)


def _world_linear_velocity(q, v):
    x, y, z, w = q[0:4]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    return R @ v[3:6]


def _slip_controlled_hopping(controller_gains, final_time, angle_of):
    """the common loop of control/SLIP_control.rs:78-131 and :133-189: build_SLIP(m 0.54, r 1, l_rest 0.2, angle 0,
    k 2000) above ground z = -0.5, SemiImplicitEuler at dt = 1/600, at every apex (world v_z turning negative) the
    test swings the leg to the commanded touch-down angle and restores its rest length by hand"""
    l_rest = 0.2
    m = Mechanism.from_model("slip", [0.54, 1.0, l_rest, 0.0, 2000.0])
    m.add_halfspace((0, 0, 1), -0.5)
    o = oracle_of(m)
    ctrl = _SLIPController(*controller_gains)
    q, v = pose_q(), np.zeros(6)
    st = o.spring_state_init()
    dt = 1.0 / 600.0
    vz_prev, vx_at_apex = 0.0, []
    for _ in range(int(final_time / dt)):
        q, v, st, flags = o.step_sc(q, v, st, dt)
        assert flags == 0
        v_lin = _world_linear_velocity(q, v)
        if vz_prev >= 0.0 and v_lin[2] < 0.0:  # apex
            angle = angle_of(ctrl, q, v)
            d = np.array([math.sin(angle), 0.0, -math.cos(angle)])
            st[0, 4:7] = d / np.linalg.norm(d)
            st[0, 7] = l_rest
            vx_at_apex.append(v_lin[0])
        vz_prev = v_lin[2]
    return q, vx_at_apex


def test_SLIP_position_control():  # control/SLIP_control.rs:78-131: hops to x = 2 and stays there, 1e-3 after 10 s
    x_target = 2.0
    q, _ = _slip_controlled_hopping((1.0, 0.5, 20.0, 0.0, 0.0), 10.0, lambda c, q, v: c.control_to_pos(q, v, x_target))
    assert abs(abs(q[4]) - x_target) < 1e-3


def test_SLIP_speed_control():  # control/SLIP_control.rs:133-189: the last three apexes pass at 0.3 m/s, 1e-3
    v_x_target = 0.3
    _, vx = _slip_controlled_hopping((1.0, 0.5, 11.0, 1.5, 1.5), 15.0, lambda c, q, v: c.control_to_velocity(q, v, v_x_target))
    assert len(vx) >= 3
    for k in (1, 2, 3):
        assert abs(vx[-k] - v_x_target) < 1e-3


def test_quadruped_inverse_kinematics():  # control/quadruped_control.rs:318-413
    from tests.controllers_ref import QuadrupedTrottingController as C
    np.testing.assert_allclose(C.inverse_kinematics([np.array([0.0, 0.0, 0.0])], 1.0)[0], [-PI / 2.0, PI], atol=1e-5)
    np.testing.assert_allclose(C.inverse_kinematics([np.array([0.0, 0.0, -1.0])], 1.0)[0], [0.0, 0.0], atol=1e-5)
    for x_t in (-0.0, -0.1, 0.1):
        t1, t2 = C.inverse_kinematics([np.array([x_t, 0.0, -0.5])], 1.0)[0]
        x = -0.5 * math.sin(-t1) + 0.5 * math.sin(t2 + t1)
        z = -0.5 * math.cos(-t1) - 0.5 * math.cos(t2 + t1)
        assert abs(x - x_t) < 1e-5 and abs(z + 0.5) < 1e-5


def test_quadruped_trot_to_position():  # control/quadruped_control.rs:415-478
    """14-dof floating base + 12 contact points, SemiImplicitEuler, dt = 1/3000, 3 s: the config-5 twin
    the reference actually tests. The gait controller is restated test-side (tests/controllers_ref.py)."""
    from tests.controllers_ref import QuadrupedTrottingController, quadruped_initial_state
    dt, target_x = 1.0 / (60.0 * 50.0), -0.2
    o = oracle_of(models.quadruped_on_ground())
    q, v = quadruped_initial_state()
    ctrl = QuadrupedTrottingController(dt, target_x, -0.8)
    for _ in range(int(3.0 / dt)):
        tau = ctrl.control(q, v)
        q, v = o.step(q, v, tau, dt, SIE)
    assert abs(q[4] - target_x) < 1e-1
    # body velocity expressed in the world-aligned frame of the test: rotation.inverse() * v_lin
    # (the reference applies the inverse rotation to the body-frame velocity; restated as written)
    x, y, z, w = q[0:4]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    assert abs((R.T @ v[3:6])[0]) < 3e-1


def _pitch(quat_xyzw):
    """nalgebra Rotation3::euler_angles().1 = -asin(R[2][0]) (nalgebra 0.33.2)"""
    x, y, z, w = quat_xyzw
    return -math.asin(2.0 * (x * z - w * y))


def _hopper_test_loop(o, q, v, dt, num_steps, hip_torque):
    """Shared frame of control/hopper_control.rs:169-237 / :272-383: mechanical stop on the leg spring,
    hip torque from `hip_torque(s, q, v, poses, twists)`, step, re-arm the spring at the bottom point."""
    prev_body_vz = 0.0
    for s in range(num_steps):
        tau = np.zeros(8)
        spring_q, spring_v = q[7], v[6]
        if spring_q > 0.0:
            tau[6] = -1e5 * (spring_q - 0.0) - 125.0 * spring_v  # mechanical_stop, energy_control.rs:20-22
        poses, twists = o.poses(q), o.body_twists(q, v)
        tau[7] = hip_torque(s, q, v, poses, twists)
        q, v = o.step(q, v, tau, dt, SIE)
        poses, twists = o.poses(q), o.body_twists(q, v)
        body_vel = twists[2, 3:6] + np.cross(twists[2, 0:3], poses[2, 4:7])
        if prev_body_vz < 0.0 and body_vel[2] >= 0.0:
            q[7] = -0.42
        prev_body_vz = body_vel[2]
    return q, v, poses, twists, body_vel


def test_hopper_foot_placement_position_control():  # control/hopper_control.rs:141-241 (SemiImplicitEuler, 10 s)
    m_foot, m_hip, m_body, l_foot_to_hip = 1.0, 0.5, 9.5, 1.0
    mech = Mechanism.from_model("hopper", [m_foot, 1.0, m_hip, 1.0, m_body, 4.0, l_foot_to_hip])
    mech.add_halfspace((0, 0, 1), 0.0, alpha=1.0, mu=1.0)
    o = oracle_of(mech)
    q, v = mech.desc().zero_state()
    q[4:7] = (0.5, 0.0, 0.5)

    def hip_torque(s, q, v, poses, twists):
        body_vx = (twists[2, 3:6] + np.cross(twists[2, 0:3], poses[2, 4:7]))[0]
        body_x = poses[2, 4]
        leg_angle = _pitch(poses[0, 0:4])
        leg_angular_v = twists[0, 1]
        vx_d = -math.copysign(1.0, body_x) * min(abs(body_x) * 10.0, 1.0)
        x_err = -0.1 * (vx_d - body_vx)
        target = -math.asin((m_body + m_hip + m_foot) * x_err / ((l_foot_to_hip + q[7]) * (m_body + m_hip)))
        return 2000.0 * (leg_angle - target) + 200.0 * leg_angular_v

    q, v, poses, *_ = _hopper_test_loop(o, q, v, 1e-3, int(10.0 / 1e-3), hip_torque)
    assert abs(poses[2, 4]) < 0.05


def test_hopper_servo_attitude_position_control():  # control/hopper_control.rs:243-395 (SemiImplicitEuler, 20 s)
    m_foot, m_hip, m_body, l_foot_to_hip = 1.0, 0.5, 9.5, 1.0
    mech = Mechanism.from_model("hopper", [m_foot, 1.0, m_hip, 1.0, m_body, 1.0, l_foot_to_hip])
    mech.add_halfspace((0, 0, 1), 0.0, alpha=1.0, mu=2.0)
    o = oracle_of(mech)
    q, v = mech.desc().zero_state()
    q[4:7] = (-3.0, 0.0, 0.5)
    dt = 5e-4
    mem = {"prev_foot_z": 0.0, "ground_hit_time": 0.0, "contact_duration": 0.0}

    def hip_torque(s, q, v, poses, twists):
        foot_z = poses[0, 6]
        t = s * dt
        if mem["prev_foot_z"] >= 0.0 and foot_z < 0.0:
            mem["ground_hit_time"] = t
        if mem["prev_foot_z"] < 0.0 and foot_z >= 0.0:
            mem["contact_duration"] = t - mem["ground_hit_time"]
        mem["prev_foot_z"] = foot_z
        if foot_z >= 0.0:  # flight: foot placement
            body_vx = (twists[2, 3:6] + np.cross(twists[2, 0:3], poses[2, 4:7]))[0]
            body_x = poses[2, 4]
            leg_angle = _pitch(poses[0, 0:4])
            leg_angular_v = twists[0, 1]
            w = l_foot_to_hip + q[7]
            dvx_max = 0.2
            vx_desired = -math.copysign(1.0, body_x) * min(abs(body_x) * 0.1, 0.2)
            dvx = vx_desired - body_vx
            x_err = -0.25 * math.copysign(1.0, dvx) * min(abs(dvx), dvx_max)
            if vx_desired < body_vx - dvx_max:
                vs = body_vx - dvx_max
            elif vx_desired > body_vx + dvx_max:
                vs = body_vx + dvx_max
            else:
                vs = vx_desired
            x_stance = vs * (0.4 if mem["contact_duration"] == 0.0 else mem["contact_duration"])
            x_touchdown = (m_body + m_hip + m_foot) * x_err / (m_body + m_hip) + x_stance / 2.0
            return 2000.0 * (leg_angle - (-math.asin(x_touchdown / w))) + 200.0 * leg_angular_v
        body_angle = _pitch(poses[2, 0:4])  # stance: servo the body attitude
        return 1200.0 * -body_angle + 60.0 * -twists[2, 1]

    q, v, poses, twists, body_vel = _hopper_test_loop(o, q, v, dt, int(20.0 / dt), hip_torque)
    assert abs(poses[2, 4]) < 1e-1 and abs(_pitch(poses[2, 0:4])) < 1e-1
    assert abs(body_vel[0]) < 1e-1 and abs(twists[2, 1]) < 1e-1


def test_pendulum_gravity_inversion():  # control/mod.rs:146-194 (RK4, dt=1e-2, 200 s)
    m, l = 5.0, 7.0
    o = oracle_of(models.hanging_rod_pendulum(m, l))
    n = simulate_step_count(200.0, 1e-2)
    q, v = o.rollout([0.1], [0.0], 1e-2, n, RK4, controller=5)
    assert abs(q[0] - PI) < 1e-3, "Pendulum should swing to the top"
    assert abs(v[0]) < 1e-4, "Pendulum should stop at the top"


def test_pendulum_energy_shaping():  # control/mod.rs:196-240 (SemiImplicitEuler, dt=1e-2, 50 s)
    m, l = 5.0, 7.0
    o = oracle_of(models.hanging_rod_pendulum(m, l))
    n = simulate_step_count(50.0, 1e-2)
    q, v, hq, hv = o.rollout([0.0], [0.1], 1e-2, n, SIE, controller=6, history=True)
    assert hq[:, 0].max() > 3.0, "Pendulum should swing to near the top"


def test_pendulum_swing_up_and_balance():  # control/mod.rs:98-105
    m, l = 5.0, 7.0
    o = oracle_of(models.hanging_rod_pendulum(m, l))
    # (the reference has no test of the combined law; |q - pi| is taken without wrapping, so it only
    # balances on the first approach. Checked here: which branch it takes, and the two closed forms.)
    tau_shape = o.control([1.0], [0.5], 6, [0.0])
    tau_inv = o.control([PI - 0.1], [0.5], 5, [0.0])
    assert o.control([1.0], [0.5], 7, [0.0])[0] == tau_shape[0]
    assert o.control([PI - 0.1], [0.5], 7, [0.0])[0] == tau_inv[0]
    # closed forms (control/mod.rs:61-66, :88-95) with l_c = l / 2, J about the joint axis = m l^2 / 3
    assert tau_inv[0] == pytest.approx(2.0 * m * GRAVITY * (l / 2) * math.sin(PI - 0.1) - 10.0 * 0.5, rel=1e-14)
    ke = 0.5 * (m * l * l / 3.0) * 0.25
    assert tau_shape[0] == pytest.approx(-0.1 * 0.5 * (ke - m * GRAVITY * (l / 2) * math.cos(1.0) - m * GRAVITY * l / 2), rel=1e-13)
