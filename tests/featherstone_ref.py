"""A second, independent CPU restatement of the hot path, used ONLY to cross-check the oracle.

The oracle (oracle/gp_oracle.cpp) follows the reference's own formulation: world-frame twists,
wrenches and inertias, a world->body->world Coriolis term, a partial-pivot LU. This module derives the
same quantities the textbook way (Featherstone, "Rigid Body Dynamics Algorithms", 2008): 6x6 Pluecker
transforms, RNEA and CRBA in BODY coordinates with dense numpy matrices, `numpy.linalg.solve`. Nothing
is shared with the oracle or with the CUDA kernels beyond the flat mechanism description, so agreement
pins the models the reference itself never tests (SO-101, navbot: SURVEY.md §8c "parity unpinned by the
reference") to an independent derivation of the same physics.

Conventions of the description (gorilla_physics_b200/desc.py, reference joint/*.rs, joint/mod.rs:186-303):
x_parent = R x_child + t with (R, t) = init_iso * joint motion; spatial vectors are [angular; linear];
floating joints carry q = (quat xyzw, t) and the body-frame twist as v; inertias are about the body
frame origin (moment, cross_part = m c, mass); gravity is 9.81 along -z of the world.
"""
from __future__ import annotations

import math

import numpy as np

FIXED, REVOLUTE, PRISMATIC, FLOATING = 0, 1, 2, 3
GRAVITY = 9.81


def skew(a):
    return np.array([[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]])


def quat_to_rot(x, y, z, w):
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def rodrigues(axis, angle):
    k = skew(axis)
    return np.eye(3) + math.sin(angle) * k + (1.0 - math.cos(angle)) * (k @ k)


def plucker(R, t):
    """motion transform parent -> child coordinates for x_parent = R x_child + t (RBDA eq. 2.24)"""
    E = R.T
    X = np.zeros((6, 6))
    X[:3, :3] = E
    X[3:, 3:] = E
    X[3:, :3] = -E @ skew(t)
    return X


def crm(v):
    out = np.zeros((6, 6))
    out[:3, :3] = skew(v[:3])
    out[3:, 3:] = skew(v[:3])
    out[3:, :3] = skew(v[3:])
    return out


def crf(v):
    return -crm(v).T


class Model:
    def __init__(self, desc):
        self.nb = int(desc.n_bodies)
        self.parent = [int(p) - 1 for p in desc.parent]  # -1 = world
        self.jtype = [int(t) for t in desc.joint_type]
        self.axis = np.asarray(desc.axis, dtype=float).reshape(-1, 3)
        iso = np.asarray(desc.init_iso, dtype=float).reshape(-1, 7)
        self.R0 = [quat_to_rot(*iso[i, :4]) for i in range(self.nb)]
        self.t0 = [iso[i, 4:7].copy() for i in range(self.nb)]
        moment = np.asarray(desc.moment, dtype=float).reshape(-1, 3, 3)
        cross = np.asarray(desc.cross_part, dtype=float).reshape(-1, 3)
        mass = np.asarray(desc.mass, dtype=float)
        self.I = []
        for i in range(self.nb):
            I = np.zeros((6, 6))
            I[:3, :3] = moment[i]
            I[:3, 3:] = skew(cross[i])
            I[3:, :3] = skew(cross[i]).T
            I[3:, 3:] = mass[i] * np.eye(3)
            self.I.append(I)
        self.has_spring = np.asarray(desc.has_spring)
        self.spring_k = np.asarray(desc.spring_k, dtype=float)
        self.spring_l = np.asarray(desc.spring_l, dtype=float)
        arm = getattr(desc, "armature", None)
        self.armature = np.zeros(self.nb) if arm is None or len(arm) != self.nb else np.asarray(arm, dtype=float)
        nq = {FIXED: 0, REVOLUTE: 1, PRISMATIC: 1, FLOATING: 7}
        nv = {FIXED: 0, REVOLUTE: 1, PRISMATIC: 1, FLOATING: 6}
        self.qoff, self.voff = [], []
        a = b = 0
        for t in self.jtype:
            self.qoff.append(a)
            self.voff.append(b)
            a += nq[t]
            b += nv[t]
        self.n_q, self.n_v = a, b
        self.cp_body = [int(x) - 1 for x in desc.cp_body] if desc.n_contact_points else []
        self.cp_loc = np.asarray(desc.cp_location, dtype=float).reshape(-1, 3)
        self.cp_k = np.asarray(desc.cp_k, dtype=float)
        self.hs_point = np.asarray(desc.hs_point, dtype=float).reshape(-1, 3)
        self.hs_normal = np.asarray(desc.hs_normal, dtype=float).reshape(-1, 3)
        self.hs_alpha = np.asarray(desc.hs_alpha, dtype=float)
        self.hs_mu = np.asarray(desc.hs_mu, dtype=float)

    def subspace(self, i):
        t = self.jtype[i]
        if t == REVOLUTE:
            return np.concatenate([self.axis[i], np.zeros(3)]).reshape(6, 1)
        if t == PRISMATIC:
            return np.concatenate([np.zeros(3), self.axis[i]]).reshape(6, 1)
        if t == FLOATING:
            return np.eye(6)
        return np.zeros((6, 0))

    def joint_pose(self, i, q):
        """(R, t): x_parent = R x_child + t"""
        t = self.jtype[i]
        R, p = self.R0[i], self.t0[i]
        if t == REVOLUTE:
            return R @ rodrigues(self.axis[i], q[self.qoff[i]]), p
        if t == PRISMATIC:
            return R, p + R @ (self.axis[i] * q[self.qoff[i]])
        if t == FLOATING:
            o = self.qoff[i]
            return R @ quat_to_rot(*q[o:o + 4]), p + R @ q[o + 4:o + 7]
        return R, p


def contact_force_law(z, zdot_vel, n, k_a, alpha, mu, k_b=50e3, v_slip=1e-3):
    """Hunt-Crossley normal force + regularised Coulomb friction on one point, world frame
    (what reference contact.rs:260-302, :321-338 computes, written from the formulas in SURVEY.md §8a C2)."""
    if not z > 0.0:  # inside the 1e-8 margin but not penetrating: z^1.5 is NaN in the reference, max(NaN, 0) = 0
        return np.zeros(3)
    v = zdot_vel
    k = k_a * k_b / (k_a + k_b)
    zd = -float(v @ n)
    zn = z ** 1.5
    normal = max(1.5 * alpha * k * zn * zd + k * zn, 0.0)
    vt = v + zd * n
    speed = float(np.linalg.norm(vt))
    if speed == 0.0:
        return normal * n
    mu_eff = mu * min(1.0, speed / v_slip)
    return normal * n - mu_eff * normal * vt / speed


def dynamics(model: Model, q, v, tau=None, gravity=GRAVITY):
    """-> dict(vdot, mass_matrix, bias, contact_forces[n_cp, 3] world frame)"""
    nb, nv = model.nb, model.n_v
    q = np.asarray(q, dtype=float)
    v = np.asarray(v, dtype=float)
    tau = np.zeros(nv) if tau is None else np.asarray(tau, dtype=float).copy()
    X, S, vel, acc, Rw, pw = [], [], [], [], [], []
    a_world = np.array([0, 0, 0, 0, 0, gravity])  # the world "accelerates upwards"
    for i in range(nb):
        R, t = model.joint_pose(i, q)
        Xi = plucker(R, t)
        Si = model.subspace(i)
        k = Si.shape[1]
        vj = Si @ v[model.voff[i]:model.voff[i] + k]
        p = model.parent[i]
        if p < 0:
            vi = vj
            ai = Xi @ a_world
            Rw.append(R)
            pw.append(t)
        else:
            vi = Xi @ vel[p] + vj
            ai = Xi @ acc[p] + crm(vi) @ vj
            Rw.append(Rw[p] @ R)
            pw.append(pw[p] + Rw[p] @ t)
        X.append(Xi)
        S.append(Si)
        vel.append(vi)
        acc.append(ai)
    # external (contact) wrenches in body coordinates
    fext = [np.zeros(6) for _ in range(nb)]
    cf = np.zeros((len(model.cp_body), 3))
    for c, b in enumerate(model.cp_body):
        loc = model.cp_loc[c]
        x_w = Rw[b] @ loc + pw[b]
        v_w = Rw[b] @ (vel[b][3:] + np.cross(vel[b][:3], loc))
        for h in range(len(model.hs_point)):
            n = model.hs_normal[h]
            d = float((x_w - model.hs_point[h]) @ n)
            if d <= 1e-8:
                cf[c] += contact_force_law(-d, v_w, n, model.cp_k[c], model.hs_alpha[h], model.hs_mu[h])
        f_b = Rw[b].T @ cf[c]
        fext[b] += np.concatenate([np.cross(loc, f_b), f_b])
    # RNEA with zero joint accelerations -> bias
    f = [model.I[i] @ acc[i] + crf(vel[i]) @ (model.I[i] @ vel[i]) - fext[i] for i in range(nb)]
    bias = np.zeros(nv)
    for i in reversed(range(nb)):
        k = S[i].shape[1]
        bias[model.voff[i]:model.voff[i] + k] = S[i].T @ f[i]
        if model.parent[i] >= 0:
            f[model.parent[i]] = f[model.parent[i]] + X[i].T @ f[i]
    # CRBA
    Ic = [I.copy() for I in model.I]
    for i in reversed(range(nb)):
        if model.parent[i] >= 0:
            Ic[model.parent[i]] += X[i].T @ Ic[i] @ X[i]
    H = np.zeros((nv, nv))
    for i in range(nb):
        k = S[i].shape[1]
        if k == 0:
            continue
        F = Ic[i] @ S[i]
        oi = model.voff[i]
        H[oi:oi + k, oi:oi + k] = S[i].T @ F
        j = i
        while model.parent[j] >= 0:
            F = X[j].T @ F
            j = model.parent[j]
            kj = S[j].shape[1]
            if kj:
                oj = model.voff[j]
                H[oi:oi + k, oj:oj + kj] = F.T @ S[j]
                H[oj:oj + kj, oi:oi + k] = S[j].T @ F
    for i in range(nb):
        # (no armature here: MechanismState's dynamics never reads it, only hybrid::Articulated::free_velocity does)
        if model.jtype[i] == PRISMATIC and model.has_spring[i]:
            tau[model.voff[i]] += -model.spring_k[i] * (q[model.qoff[i]] - model.spring_l[i])
    vdot = np.linalg.solve(H, tau - bias) if nv else np.zeros(0)
    return {"vdot": vdot, "mass_matrix": H, "bias": bias, "contact_forces": cf}


def potential_energy(model: Model, q, gravity=GRAVITY):
    """true gravitational potential energy: sum over bodies of g * (m * origin_z + (R_world * mc)_z)
    (the reference's gravitational_energy uses the frame-origin height only, mechanism.rs:352-362)"""
    q = np.asarray(q, dtype=float)
    Rw, pw, pe = [], [], 0.0
    for i in range(model.nb):
        R, t = model.joint_pose(i, q)
        p = model.parent[i]
        if p < 0:
            Rw.append(R)
            pw.append(t)
        else:
            Rw.append(Rw[p] @ R)
            pw.append(pw[p] + Rw[p] @ t)
        mass = model.I[i][3, 3]
        mc = np.array([model.I[i][2, 4], model.I[i][0, 5], model.I[i][1, 3]])  # from skew(mc) in the upper-right block
        pe += gravity * (mass * pw[i][2] + (Rw[i] @ mc)[2])
    return pe


def kinetic_energy(model: Model, q, v):
    """1/2 v^T H v"""
    H = dynamics(model, q, v)["mass_matrix"]
    v = np.asarray(v, dtype=float)
    return 0.5 * float(v @ H @ v)


def semi_implicit_euler_step(model: Model, q, v, tau, dt):
    """One step of the reference's semi-implicit Euler, written from its definition (SURVEY.md §8a I1,
    integrators.rs:25-39, :276-319): v+ = v + vdot dt; scalar joints q+ = q + v+ dt; floating joints
    quat+ = normalize(quat + 1/2 quat * (0, w+) dt) (w+ the NEW body-frame angular velocity),
    t+ = t + R(quat) v_lin+ dt with the OLD rotation."""
    q = np.asarray(q, dtype=float)
    v = np.asarray(v, dtype=float)
    vdot = dynamics(model, q, v, tau)["vdot"]
    v1 = v + vdot * dt
    q1 = q.copy()
    for i in range(model.nb):
        t = model.jtype[i]
        qo, vo = model.qoff[i], model.voff[i]
        if t in (REVOLUTE, PRISMATIC):
            q1[qo] = q[qo] + v1[vo] * dt
        elif t == FLOATING:
            x, y, z, w = q[qo:qo + 4]
            wx, wy, wz = v1[vo:vo + 3]
            # Hamilton product quat * (wx, wy, wz, 0), halved
            dq = 0.5 * np.array([w * wx + y * wz - z * wy,
                                 w * wy + z * wx - x * wz,
                                 w * wz + x * wy - y * wx,
                                 -x * wx - y * wy - z * wz])
            qn = q[qo:qo + 4] + dq * dt
            q1[qo:qo + 4] = qn / np.linalg.norm(qn)
            q1[qo + 4:qo + 7] = q[qo + 4:qo + 7] + quat_to_rot(x, y, z, w) @ v1[vo + 3:vo + 6] * dt
    return q1, v1
