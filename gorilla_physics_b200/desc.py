"""Flat mechanism description — the Python view of `gp_mechanism_desc` (include/gorilla_b200.h).

Host-side mirror of what the reference passes to `MechanismState::new(treejoints, bodies)`
(src/mechanism.rs:62-148) plus `add_contact_point` / `add_halfspace` (:379-392). Joint i's
child body is body i; ids are 1-based, 0 is the world. Pure numpy; no GPU needed.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np

FIXED, REVOLUTE, PRISMATIC, FLOATING = 0, 1, 2, 3
JOINT_NQ = {FIXED: 0, REVOLUTE: 1, PRISMATIC: 1, FLOATING: 7}
JOINT_NV = {FIXED: 0, REVOLUTE: 1, PRISMATIC: 1, FLOATING: 6}

IDENTITY_ISO = (0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0)  # quaternion x,y,z,w + translation


def quat_mul(a, b):
    """Hamilton product of quaternions given as (x, y, z, w)."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz,
    ])


def quat_from_euler(roll, pitch, yaw):
    """UnitQuaternion::from_euler_angles as used by Transform3D::new_xyz_rpy
    (src/spatial/transform.rs:77-94): R = Rz(yaw) Ry(pitch) Rx(roll). Returns x,y,z,w."""
    sr, cr = math.sin(roll * 0.5), math.cos(roll * 0.5)
    sp, cp = math.sin(pitch * 0.5), math.cos(pitch * 0.5)
    sy, cy = math.sin(yaw * 0.5), math.cos(yaw * 0.5)
    return np.array([
        sr * cp * cy - cr * sp * sy,
        cr * sp * cy + sr * cp * sy,
        cr * cp * sy - sr * sp * cy,
        cr * cp * cy + sr * sp * sy,
    ])


def quat_from_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    s, c = math.sin(angle / 2.0), math.cos(angle / 2.0)
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, c])


def quat_from_scaled_axis(aa):
    aa = np.asarray(aa, dtype=np.float64)
    n = float(np.linalg.norm(aa))
    if n == 0.0:
        return np.array([0.0, 0.0, 0.0, 1.0])
    return quat_from_axis_angle(aa / n, n)


def iso(translation=(0.0, 0.0, 0.0), quat=(0.0, 0.0, 0.0, 1.0)):
    return np.concatenate([np.asarray(quat, dtype=np.float64), np.asarray(translation, dtype=np.float64)])


def iso_xyz_rpy(xyz, rpy):
    """Transform3D::new_xyz_rpy (src/spatial/transform.rs:77-94)."""
    return iso(xyz, quat_from_euler(*rpy))


class MechanismDesc:
    """Mutable builder + array view of one mechanism."""

    def __init__(self):
        self._parent, self._jtype, self._axis, self._iso = [], [], [], []
        self._moment, self._cross, self._mass = [], [], []
        self._has_spring, self._spring_k, self._spring_l = [], [], []
        self._armature = []
        self._sc_body, self._sc_l_rest, self._sc_dir, self._sc_k = [], [], [], []
        self._cp_body, self._cp_loc, self._cp_k = [], [], []
        self._hs_point, self._hs_normal, self._hs_alpha, self._hs_mu = [], [], [], []
        self.names: list[str] = []

    # ---- construction ------------------------------------------------------
    def add_body(self, parent: int, joint_type: int, *, axis=(0.0, 0.0, 1.0), init_iso=IDENTITY_ISO,
                 moment=None, cross_part=(0.0, 0.0, 0.0), mass=0.0,
                 spring: Optional[Sequence[float]] = None, armature: float = 0.0, name: str = "") -> int:
        """Append joint+body; returns the new 1-based body id. `parent` must already exist
        (the reference resolves parents by frame name and panics otherwise, mechanism.rs:98-116)."""
        body_id = len(self._parent) + 1
        if not (0 <= parent < body_id):
            raise ValueError(f"joint {body_id} has no parent body {parent}")
        if joint_type not in JOINT_NQ:
            raise ValueError(f"unknown joint type {joint_type}")
        moment = np.zeros((3, 3)) if moment is None else np.asarray(moment, dtype=np.float64).reshape(3, 3)
        self._parent.append(int(parent))
        self._jtype.append(int(joint_type))
        self._axis.append(np.asarray(axis, dtype=np.float64).reshape(3))
        self._iso.append(np.asarray(init_iso, dtype=np.float64).reshape(7))
        self._moment.append(moment.reshape(9))
        self._cross.append(np.asarray(cross_part, dtype=np.float64).reshape(3))
        self._mass.append(float(mass))
        self._has_spring.append(0 if spring is None else 1)
        self._spring_k.append(0.0 if spring is None else float(spring[0]))
        self._spring_l.append(0.0 if spring is None else float(spring[1]))
        self._armature.append(float(armature))
        self.names.append(name or f"body{body_id}")
        return body_id

    def add_contact_point(self, body: int, location, k: float = 50e3):
        """ContactPoint::new / new_with_k (src/contact.rs:24-38) + add_contact_point (mechanism.rs:384)."""
        if not (1 <= body <= len(self._parent)):
            raise ValueError("contact point frame does not match a body")
        self._cp_body.append(int(body))
        self._cp_loc.append(np.asarray(location, dtype=np.float64).reshape(3))
        self._cp_k.append(float(k))

    def add_spring_contact(self, body: int, l_rest: float, direction, k: float):
        """SpringContact::new + add_spring_contact (src/contact.rs:83-94, mechanism.rs:394-401): an ideal spring
        leg attached at the body origin, pointing along `direction` (unit, body frame)."""
        if not (1 <= body <= len(self._parent)):
            raise ValueError("spring contact frame does not match a body")
        self._sc_body.append(int(body))
        self._sc_l_rest.append(float(l_rest))
        self._sc_dir.append(np.asarray(direction, dtype=np.float64).reshape(3))
        self._sc_k.append(float(k))

    def add_halfspace(self, normal, distance: float, alpha: float = 0.9, mu: float = 0.5):
        """HalfSpace::new / new_with_params (src/collision/halfspace.rs:15-37): point = normal * distance."""
        n = np.asarray(normal, dtype=np.float64).reshape(3)
        self._hs_point.append(n * distance)
        self._hs_normal.append(n)
        self._hs_alpha.append(float(alpha))
        self._hs_mu.append(float(mu))

    # ---- array view (gp_mechanism_desc fields) --------------------------------
    @property
    def n_bodies(self):
        return len(self._parent)

    @property
    def n_contact_points(self):
        return len(self._cp_body)

    @property
    def n_halfspaces(self):
        return len(self._hs_point)

    @property
    def parent(self):
        return np.asarray(self._parent, dtype=np.int32)

    @property
    def joint_type(self):
        return np.asarray(self._jtype, dtype=np.int32)

    @property
    def axis(self):
        return np.asarray(self._axis, dtype=np.float64).reshape(-1, 3)

    @property
    def init_iso(self):
        return np.asarray(self._iso, dtype=np.float64).reshape(-1, 7)

    @property
    def moment(self):
        return np.asarray(self._moment, dtype=np.float64).reshape(-1, 9)

    @property
    def cross_part(self):
        return np.asarray(self._cross, dtype=np.float64).reshape(-1, 3)

    @property
    def mass(self):
        return np.asarray(self._mass, dtype=np.float64)

    @property
    def has_spring(self):
        return np.asarray(self._has_spring, dtype=np.int32)

    @property
    def spring_k(self):
        return np.asarray(self._spring_k, dtype=np.float64)

    @property
    def spring_l(self):
        return np.asarray(self._spring_l, dtype=np.float64)

    @property
    def armature(self):
        """reflected drivetrain inertia on the joint's own diagonal of M (reference revolute.rs:29)"""
        return np.asarray(self._armature, dtype=np.float64)

    @property
    def n_spring_contacts(self):
        return len(self._sc_body)

    @property
    def sc_body(self):
        return np.asarray(self._sc_body, dtype=np.int32)

    @property
    def sc_l_rest(self):
        return np.asarray(self._sc_l_rest, dtype=np.float64)

    @property
    def sc_direction(self):
        return np.asarray(self._sc_dir, dtype=np.float64).reshape(-1, 3)

    @property
    def sc_k(self):
        return np.asarray(self._sc_k, dtype=np.float64)

    @property
    def cp_body(self):
        return np.asarray(self._cp_body, dtype=np.int32)

    @property
    def cp_location(self):
        return np.asarray(self._cp_loc, dtype=np.float64).reshape(-1, 3)

    @property
    def cp_k(self):
        return np.asarray(self._cp_k, dtype=np.float64)

    @property
    def hs_point(self):
        return np.asarray(self._hs_point, dtype=np.float64).reshape(-1, 3)

    @property
    def hs_normal(self):
        return np.asarray(self._hs_normal, dtype=np.float64).reshape(-1, 3)

    @property
    def hs_alpha(self):
        return np.asarray(self._hs_alpha, dtype=np.float64)

    @property
    def hs_mu(self):
        return np.asarray(self._hs_mu, dtype=np.float64)

    @property
    def n_q(self):
        return sum(JOINT_NQ[t] for t in self._jtype)

    @property
    def n_v(self):
        return sum(JOINT_NV[t] for t in self._jtype)

    def q_offsets(self):
        off, out = 0, []
        for t in self._jtype:
            out.append(off)
            off += JOINT_NQ[t]
        return out

    def v_offsets(self):
        off, out = 0, []
        for t in self._jtype:
            out.append(off)
            off += JOINT_NV[t]
        return out

    def zero_state(self):
        """MechanismState::new initial condition (mechanism.rs:71-88): zeros, identity pose."""
        q = np.zeros(self.n_q)
        for t, o in zip(self._jtype, self.q_offsets()):
            if t == FLOATING:
                q[o + 3] = 1.0
        return q, np.zeros(self.n_v)

    def copy(self) -> "MechanismDesc":
        import copy
        return copy.deepcopy(self)

    @classmethod
    def from_arrays(cls, **kw) -> "MechanismDesc":
        d = cls()
        nb = int(kw["n_bodies"])
        hs = kw.get("has_spring")
        for i in range(nb):
            spring = None
            if hs is not None and int(hs[i]):
                spring = (float(kw["spring_k"][i]), float(kw["spring_l"][i]))
            d.add_body(int(kw["parent"][i]), int(kw["joint_type"][i]), axis=np.asarray(kw["axis"]).reshape(-1, 3)[i],
                       init_iso=np.asarray(kw["init_iso"]).reshape(-1, 7)[i],
                       moment=np.asarray(kw["moment"]).reshape(-1, 9)[i],
                       cross_part=np.asarray(kw["cross_part"]).reshape(-1, 3)[i], mass=float(kw["mass"][i]),
                       spring=spring, armature=float(kw["armature"][i]) if kw.get("armature") is not None else 0.0)
        for c in range(int(kw.get("n_contact_points", 0))):
            d.add_contact_point(int(kw["cp_body"][c]), np.asarray(kw["cp_location"]).reshape(-1, 3)[c],
                                float(kw["cp_k"][c]))
        for h in range(int(kw.get("n_halfspaces", 0))):
            d._hs_point.append(np.asarray(kw["hs_point"], dtype=np.float64).reshape(-1, 3)[h].copy())
            d._hs_normal.append(np.asarray(kw["hs_normal"], dtype=np.float64).reshape(-1, 3)[h].copy())
            d._hs_alpha.append(float(kw["hs_alpha"][h]))
            d._hs_mu.append(float(kw["hs_mu"][h]))
        return d
