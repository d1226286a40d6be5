"""Shared test fixtures: mechanisms used by the reference's own tests, built either through the
product's model builders (gp_model_create) or ad hoc through MechanismDesc, plus helpers that
hand the same description to the oracle."""
import math

import numpy as np

from gorilla_physics_b200 import (FIXED, FLOATING, PRISMATIC, REVOLUTE, Mechanism, MechanismDesc, iso,
                                  quat_from_euler, quat_from_scaled_axis)

GRAVITY = 9.81
PI = math.pi


def oracle_of(mech_or_desc):
    # (imported here, not at module level: bench.py's own arm takes its mechanisms from this module and must
    # not pull the oracle in; only its cpu_baseline / --impl reference legs may)
    from oracle.binding import OracleMechanism
    desc = mech_or_desc.desc() if isinstance(mech_or_desc, Mechanism) else mech_or_desc
    return OracleMechanism(desc)


def rod_pendulum(m=5.0, l=7.0, rod_to_world=None, axis=(0.0, 1.0, 0.0), point_mass=False):
    """reference dynamics.rs:883-1035 fixtures"""
    d = MechanismDesc()
    f = 1.0 if point_mass else 1.0 / 3.0
    cp = (m * l, 0.0, 0.0) if point_mass else (m * l / 2.0, 0.0, 0.0)
    d.add_body(0, REVOLUTE, axis=axis, init_iso=iso() if rod_to_world is None else rod_to_world,
               moment=np.diag([0.0, f * m * l * l, f * m * l * l]), cross_part=cp, mass=m)
    return d


def hanging_rod_pendulum(m=5.0, l=7.0):
    """reference control/mod.rs:147-165 fixture: rod hanging along -z from the joint, axis +y, q = 0 at the bottom"""
    d = MechanismDesc()
    d.add_body(0, REVOLUTE, axis=(0.0, 1.0, 0.0), init_iso=iso(),
               moment=np.diag([m * l * l / 3.0, m * l * l / 3.0, 0.0]), cross_part=(0.0, 0.0, -m * l / 2.0), mass=m)
    return d


def double_pendulum_horizontal(m=5.0, l=7.0, axis=(0.0, 1.0, 0.0)):
    """reference dynamics.rs:1038-1087 / examples/acrobot.rs (axis -y, m=1)"""
    d = MechanismDesc()
    mom = np.diag([0.0, m * l * l, m * l * l])
    d.add_body(0, REVOLUTE, axis=axis, moment=mom, cross_part=(m * l, 0, 0), mass=m)
    d.add_body(1, REVOLUTE, axis=axis, init_iso=iso((l, 0, 0)), moment=mom, cross_part=(m * l, 0, 0), mass=m)
    return d


def double_pendulum_hanging(m=3.0, l=5.0, axis=(0.0, 1.0, 0.0)):
    """reference dynamics.rs:1091-1130, energy.rs:82-127"""
    d = MechanismDesc()
    mom = np.diag([m * l * l, m * l * l, 0.0])
    d.add_body(0, REVOLUTE, axis=axis, moment=mom, cross_part=(0, 0, -m * l), mass=m)
    d.add_body(1, REVOLUTE, axis=axis, init_iso=iso((0, 0, -l)), moment=mom, cross_part=(0, 0, -m * l), mass=m)
    return d


def cart_pole(m_cart, l_cart, m_pole, l_pole, axis_pole):
    """reference simulate.rs:218-272, energy.rs:130-178"""
    d = MechanismDesc()
    d.add_body(0, PRISMATIC, axis=(1, 0, 0),
               moment=np.diag([0.0, m_cart * l_cart * l_cart / 12.0, m_cart * l_cart * l_cart / 12.0]), mass=m_cart)
    d.add_body(1, REVOLUTE, axis=axis_pole, moment=np.diag([m_pole * l_pole * l_pole, m_pole * l_pole * l_pole, 0.0]),
               cross_part=(0, 0, -l_pole * m_pole), mass=m_pole)
    return d


def sphere_moment(m, r):
    return np.eye(3) * (2.0 / 5.0 * m * r * r)


def ball(m=5.0, r=1.0):
    d = MechanismDesc()
    d.add_body(0, FLOATING, moment=sphere_moment(m, r), mass=m)
    return d


def motor_turning_mass():
    """reference dynamics.rs:1194-1251"""
    d = MechanismDesc()
    d.add_body(0, FLOATING, moment=sphere_moment(1.0, 1.0), mass=1.0)
    d.add_body(1, REVOLUTE, axis=(0, 0, 1), moment=np.diag([1.0, 0.0, 1.0]), cross_part=(0.0, -1.0, 0.0), mass=1.0)
    return d


def mass_matrix_fixture():
    """reference mechanism.rs:711-786"""
    m_body, w_body, h_body = 2.0, 1.0, 0.1
    mx = (w_body * w_body + h_body * h_body) * m_body / 12.0
    mz = (w_body * w_body + w_body * w_body) * m_body / 12.0
    m_leg, w_leg, h_leg = 1.0, 0.1, 1.0
    lx = m_leg * ((w_leg * w_leg + h_leg * h_leg) / 12.0 + (h_leg / 2.0 * h_leg / 2.0))
    lz = (w_leg * w_leg + w_leg * w_leg) * m_leg / 12.0
    d = MechanismDesc()
    d.add_body(0, FLOATING, moment=np.diag([mx, mx, mz]), mass=m_body)
    d.add_body(1, REVOLUTE, axis=(0, 1, 0), moment=np.diag([lx, lx, lz]), cross_part=(0, 0, -h_leg / 2.0 * m_leg),
               mass=m_leg)
    d.add_body(2, REVOLUTE, axis=(0, 1, 0), init_iso=iso((0, 0, -h_leg)), moment=np.diag([lx, lx, lz]),
               cross_part=(0, 0, -h_leg / 2.0 * m_leg), mass=m_leg)
    return d


def supports_fixture():
    """reference mechanism.rs:793-832:   3       5
                                          |       |
                                          2 - 1 - 4 """
    d = MechanismDesc()
    s = sphere_moment(1.0, 1.0)
    d.add_body(0, FLOATING, moment=s, mass=1.0)
    d.add_body(1, REVOLUTE, axis=(0, 1, 0), init_iso=iso((-1, 0, 0)), moment=s, mass=1.0)
    d.add_body(2, REVOLUTE, axis=(0, 1, 0), init_iso=iso((0, 0, 1)), moment=s, mass=1.0)
    d.add_body(1, REVOLUTE, axis=(0, 1, 0), init_iso=iso((1, 0, 0)), moment=s, mass=1.0)
    d.add_body(4, REVOLUTE, axis=(0, 1, 0), init_iso=iso((0, 0, 1)), moment=s, mass=1.0)
    return d


def spring_pair():
    """reference dynamics.rs:1133-1190 spring_on_frictionless_ground"""
    m, r, l_init = 1.0, 0.1, 2.0
    d = MechanismDesc()
    d.add_body(0, FLOATING, moment=sphere_moment(m, r), mass=m)
    d.add_body(1, PRISMATIC, axis=(1, 0, 0), moment=sphere_moment(m, r), mass=m, spring=(50.0, l_init / 3.0))
    d.add_halfspace((0, 0, 1), 0.0, alpha=1.0, mu=0.0)
    return d, l_init


def pose_q(rpy=(0.0, 0.0, 0.0), t=(0.0, 0.0, 0.0)):
    return np.concatenate([quat_from_euler(*rpy), np.asarray(t, dtype=float)])


# ---- the benchmark / parity workloads of SURVEY.md §8d: owned by the package (gorilla_physics_b200/workloads.py)
from gorilla_physics_b200.workloads import (acrobot, hopper1d_on_ground, navbot_with_contact,  # noqa: E402,F401
                                            quadruped_on_ground, rimless_wheel_on_slope, so101_with_contact)


def random_tree(seed: int, n_bodies: int, max_dof: int = 24, contact: bool = True):
    """A random mechanism for the run-time-topology kernel: random parents (any earlier body or the
    world), mixed joint types (floating joints may hang off other bodies, fixed joints may sit in
    the middle of a chain), random unit axes, random joint origins, random positive-definite
    inertias, prismatic joint springs, contact points and one or two halfspaces."""
    rng = np.random.default_rng(seed)
    d = MechanismDesc()
    dof = 0
    for i in range(n_bodies):
        parent = int(rng.integers(0, i + 1))
        choices = [REVOLUTE, REVOLUTE, PRISMATIC, FIXED]
        if dof + 6 <= max_dof - (n_bodies - i - 1):
            choices.append(FLOATING)
        jt = int(rng.choice(choices))
        if dof + (6 if jt == FLOATING else 1) > max_dof:
            jt = FIXED
        dof += {FIXED: 0, REVOLUTE: 1, PRISMATIC: 1, FLOATING: 6}[jt]
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        aa = rng.normal(size=3) * 0.7
        origin = iso(rng.uniform(-0.3, 0.3, size=3), quat_from_scaled_axis(aa))
        m = float(rng.uniform(0.2, 2.0))
        com = rng.uniform(-0.1, 0.1, size=3)
        a = rng.normal(size=(3, 3))
        j_com = a @ a.T * 0.01 + np.eye(3) * 0.01
        moment = j_com + m * (com @ com * np.eye(3) - np.outer(com, com))
        spring = (float(rng.uniform(10, 100)), float(rng.uniform(-0.2, 0.2))) if (jt == PRISMATIC and rng.random() < 0.5) else None
        d.add_body(parent, jt, axis=axis, init_iso=origin, moment=moment, cross_part=m * com, mass=m, spring=spring)
    if contact:
        for _ in range(int(rng.integers(1, 9))):
            d.add_contact_point(int(rng.integers(1, n_bodies + 1)), rng.uniform(-0.2, 0.2, size=3),
                                k=float(rng.choice([50e3, 10e3, 75e3])))
        d.add_halfspace((0, 0, 1), -0.2, alpha=0.9, mu=0.5)
        if rng.random() < 0.5:
            n = np.array([0.3, 0.1, 1.0])
            d.add_halfspace(n / np.linalg.norm(n), -0.4, alpha=1.0, mu=1.0)
    return d


def cube_in_corner() -> Mechanism:
    """two halfspaces (ground + tilted wall): exercises the multi-halfspace contact mode"""
    m = Mechanism.from_model("cube")
    m.add_halfspace((0, 0, 1), -0.45, alpha=0.9, mu=0.5)
    n = np.array([1.0, 0.0, 0.2])
    m.add_halfspace(n / np.linalg.norm(n), -0.45, alpha=1.0, mu=0.3)
    return m


def maximum_size_mechanism() -> MechanismDesc:
    """Every limit of the C ABI at once (include/gorilla_b200.h GP_MAX_*): 16 bodies, 24 velocity dofs
    (two floating joints, the second hanging off the first, + 12 single-dof joints + 2 fixed leaves),
    32 contact points, 4 halfspaces."""
    rng = np.random.default_rng(2026)
    d = MechanismDesc()

    def inertia():
        m = float(rng.uniform(0.3, 1.5))
        com = rng.uniform(-0.05, 0.05, size=3)
        a = rng.normal(size=(3, 3))
        return dict(moment=a @ a.T * 0.01 + np.eye(3) * 0.02 + m * (com @ com * np.eye(3) - np.outer(com, com)),
                    cross_part=m * com, mass=m)

    d.add_body(0, FLOATING, **inertia())
    d.add_body(1, FLOATING, init_iso=iso((0.1, 0.0, 0.2)), **inertia())
    for i in range(12):
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        parent = int(rng.integers(1, d.n_bodies + 1))
        jt = PRISMATIC if i % 4 == 3 else REVOLUTE
        d.add_body(parent, jt, axis=axis, init_iso=iso(rng.uniform(-0.2, 0.2, size=3), quat_from_scaled_axis(rng.normal(size=3) * 0.5)),
                   spring=(40.0, 0.05) if jt == PRISMATIC else None, **inertia())
    d.add_body(5, FIXED, init_iso=iso((0.0, 0.1, 0.0)), **inertia())
    d.add_body(15, FIXED, init_iso=iso((0.05, 0.0, 0.0)), **inertia())
    assert d.n_bodies == 16 and d.n_v == 24
    for c in range(32):
        d.add_contact_point(1 + c % 16, rng.uniform(-0.15, 0.15, size=3), k=float(rng.choice([50e3, 20e3])))
    d.add_halfspace((0, 0, 1), -0.3)
    n = np.array([0.2, 0.0, 1.0])
    d.add_halfspace(n / np.linalg.norm(n), -0.35, alpha=1.0, mu=1.0)
    n = np.array([0.0, -0.3, 1.0])
    d.add_halfspace(n / np.linalg.norm(n), -0.4, alpha=0.5, mu=0.2)
    d.add_halfspace((1, 0, 0), -0.6, alpha=0.9, mu=0.0)
    return d
