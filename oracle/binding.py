"""ctypes binding of the CPU parity oracle (oracle/libgp_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package never imports this module.

The oracle restates the reference's f64 CPU algorithm (see gp_oracle.h for the file:line
map). It takes the same flat mechanism description as the product's C ABI, as any object
with numpy-array attributes named like the fields of gp_mechanism_desc.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libgp_oracle.so"

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class _Desc(C.Structure):
    _fields_ = [
        ("n_bodies", C.c_int32),
        ("parent", _ip),
        ("joint_type", _ip),
        ("axis", _dp),
        ("init_iso", _dp),
        ("moment", _dp),
        ("cross_part", _dp),
        ("mass", _dp),
        ("has_spring", _ip),
        ("spring_k", _dp),
        ("spring_l", _dp),
        ("n_contact_points", C.c_int32),
        ("cp_body", _ip),
        ("cp_location", _dp),
        ("cp_k", _dp),
        ("n_halfspaces", C.c_int32),
        ("hs_point", _dp),
        ("hs_normal", _dp),
        ("hs_alpha", _dp),
        ("hs_mu", _dp),
        ("armature", _dp),
        ("n_spring_contacts", C.c_int32),
        ("sc_body", _ip),
        ("sc_l_rest", _dp),
        ("sc_direction", _dp),
        ("sc_k", _dp),
    ]


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (g++, no -march=native)."""
    src_newer = (not _LIB_PATH.exists()) or any(
        (_HERE / f).stat().st_mtime > _LIB_PATH.stat().st_mtime for f in ("gp_oracle.cpp", "gp_oracle.h")
    )
    if force or src_newer:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = configure(C.CDLL(str(_LIB_PATH)))
    return _lib


def configure(L):
    """Declare the argument types of the gpo_* entry points on a loaded library (or on a proxy that
    maps the names, see tools/count_reference_flops.py)."""
    if True:
        vp = C.c_void_p
        L.gpo_mechanism_create.argtypes = [C.POINTER(_Desc), C.POINTER(vp)]
        L.gpo_mechanism_destroy.argtypes = [vp]
        L.gpo_mechanism_destroy.restype = None
        L.gpo_n_q.argtypes = [vp]
        L.gpo_n_v.argtypes = [vp]
        L.gpo_supports.argtypes = [vp, _ip]
        L.gpo_supports.restype = None
        L.gpo_dynamics.argtypes = [vp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.gpo_step.argtypes = [vp, _dp, _dp, _dp, C.c_double, C.c_int]
        L.gpo_control.argtypes = [vp, _dp, _dp, C.c_int, _dp, _dp]
        L.gpo_rollout.argtypes = [vp, _dp, _dp, _dp, C.c_double, C.c_int64, C.c_int, C.c_int, _dp, _dp, _dp]
        L.gpo_simulate_step_count.argtypes = [C.c_double, C.c_double]
        L.gpo_simulate_step_count.restype = C.c_int64
        L.gpo_batch_rollout.argtypes = [vp, _dp, _dp, _dp, C.c_int64, C.c_double, C.c_int64, C.c_int,
                                        C.c_int, _dp, C.c_int]
        L.gpo_batch_dynamics.argtypes = [vp, _dp, _dp, _dp, C.c_int64, _dp, _dp, C.c_int]
        L.gpo_n_spring_contacts.argtypes = [vp]
        L.gpo_spring_state_init.argtypes = [vp, _dp]
        L.gpo_spring_state_init.restype = None
        L.gpo_step_sc.argtypes = [vp, _dp, _dp, _dp, C.c_double, _dp]
        L.gpo_dynamics_sc.argtypes = [vp, _dp, _dp, _dp, _dp, _dp]
        L.gpo_free_velocity.argtypes = [vp, _dp, _dp, _dp, C.c_double, C.c_int, _dp]
        L.gpo_kinetic_energy.argtypes = [vp, _dp, _dp]
        L.gpo_kinetic_energy.restype = C.c_double
        L.gpo_gravitational_energy.argtypes = [vp, _dp]
        L.gpo_gravitational_energy.restype = C.c_double
        L.gpo_spring_energy.argtypes = [vp, _dp]
        L.gpo_spring_energy.restype = C.c_double
        L.gpo_poses.argtypes = [vp, _dp, _dp]
        L.gpo_poses.restype = None
        L.gpo_body_twists.argtypes = [vp, _dp, _dp, _dp]
        L.gpo_body_twists.restype = None
        L.gpo_simple_double_pendulum.argtypes = [C.c_double] * 8 + [_dp]
        L.gpo_simple_double_pendulum.restype = None
        L.gpo_quat_from_euler.argtypes = [C.c_double] * 3 + [_dp]
        L.gpo_quat_from_euler.restype = None
        L.gpo_quat_from_axis_angle.argtypes = [_dp, C.c_double, _dp]
        L.gpo_quat_from_axis_angle.restype = None
        L.gpo_quat_from_scaled_axis.argtypes = [_dp, _dp]
        L.gpo_quat_from_scaled_axis.restype = None
        L.gpo_twist_transform.argtypes = [_dp, _dp, _dp]
        L.gpo_twist_transform.restype = None
    return L


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


class OracleMechanism:
    """One mechanism inside the oracle. `desc` has the gp_mechanism_desc fields as arrays."""

    def __init__(self, desc):
        L = lib()
        nb = int(desc.n_bodies)
        keep = self._keep = {}
        keep["parent"] = _i32(desc.parent)
        keep["joint_type"] = _i32(desc.joint_type)
        keep["axis"] = _f64(desc.axis, (nb, 3))
        keep["init_iso"] = _f64(desc.init_iso, (nb, 7))
        keep["moment"] = _f64(desc.moment, (nb, 9))
        keep["cross_part"] = _f64(desc.cross_part, (nb, 3))
        keep["mass"] = _f64(desc.mass, (nb,))
        hs = getattr(desc, "has_spring", None)
        keep["has_spring"] = _i32(hs if hs is not None else np.zeros(nb))
        sk = getattr(desc, "spring_k", None)
        sl = getattr(desc, "spring_l", None)
        keep["spring_k"] = _f64(sk if sk is not None else np.zeros(nb), (nb,))
        keep["spring_l"] = _f64(sl if sl is not None else np.zeros(nb), (nb,))
        nc = int(getattr(desc, "n_contact_points", 0))
        keep["cp_body"] = _i32(desc.cp_body if nc else np.zeros(0))
        keep["cp_location"] = _f64(desc.cp_location if nc else np.zeros((0, 3)), (nc, 3))
        keep["cp_k"] = _f64(desc.cp_k if nc else np.zeros(0), (nc,))
        nh = int(getattr(desc, "n_halfspaces", 0))
        keep["hs_point"] = _f64(desc.hs_point if nh else np.zeros((0, 3)), (nh, 3))
        keep["hs_normal"] = _f64(desc.hs_normal if nh else np.zeros((0, 3)), (nh, 3))
        keep["hs_alpha"] = _f64(desc.hs_alpha if nh else np.zeros(0), (nh,))
        keep["hs_mu"] = _f64(desc.hs_mu if nh else np.zeros(0), (nh,))
        arm = getattr(desc, "armature", None)
        keep["armature"] = _f64(arm if arm is not None and len(arm) == nb else np.zeros(nb), (nb,))
        ns = int(getattr(desc, "n_spring_contacts", 0))
        keep["sc_body"] = _i32(desc.sc_body if ns else np.zeros(0))
        keep["sc_l_rest"] = _f64(desc.sc_l_rest if ns else np.zeros(0), (ns,))
        keep["sc_direction"] = _f64(desc.sc_direction if ns else np.zeros((0, 3)), (ns, 3))
        keep["sc_k"] = _f64(desc.sc_k if ns else np.zeros(0), (ns,))
        d = _Desc()
        d.n_bodies = nb
        d.n_contact_points = nc
        d.n_halfspaces = nh
        d.n_spring_contacts = ns
        self.n_sc = ns
        for name, arr in keep.items():
            ptr_t = _ip if arr.dtype == np.int32 else _dp
            setattr(d, name, arr.ctypes.data_as(ptr_t))
        h = C.c_void_p()
        rc = L.gpo_mechanism_create(C.byref(d), C.byref(h))
        if rc != 0:
            raise ValueError("oracle: malformed mechanism description")
        self._h = h
        self.n_bodies = nb
        self.n_q = L.gpo_n_q(h)
        self.n_v = L.gpo_n_v(h)
        self.n_cp = nc
        # body-major order of the contact points (order of the flat contact_forces output)
        order = np.argsort(keep["cp_body"], kind="stable") if nc else np.zeros(0, dtype=np.int64)
        self.cp_order = order

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().gpo_mechanism_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- single environment -------------------------------------------------
    def dynamics(self, q, v, tau=None, want="vdot"):
        """dynamics_continuous (dynamics.rs:322). Returns vdot, or a dict with
        vdot / contact_forces / mass_matrix / bias when want == 'all'."""
        q = _f64(q, (self.n_q,))
        v = _f64(v, (self.n_v,))
        tau = None if tau is None else _f64(tau, (self.n_v,))
        vdot = np.zeros(self.n_v)
        cf = np.zeros((self.n_cp, 3))
        M = np.zeros((self.n_v, self.n_v))
        c = np.zeros(self.n_v)
        rc = lib().gpo_dynamics(self._h, _d(q), _d(v), _d(tau), _d(vdot), _d(cf), _d(M), _d(c))
        if rc != 0:
            raise ArithmeticError("oracle: singular mass matrix")
        if want == "all":
            return {"vdot": vdot, "contact_forces": cf, "mass_matrix": M, "bias": c}
        return vdot

    def step(self, q, v, tau=None, dt=1e-3, integrator=0):
        q = _f64(q, (self.n_q,)).copy()
        v = _f64(v, (self.n_v,)).copy()
        tau = None if tau is None else _f64(tau, (self.n_v,))
        rc = lib().gpo_step(self._h, _d(q), _d(v), _d(tau), dt, integrator)
        if rc == 2:
            raise ValueError("oracle: unsupported integrator")
        return q, v

    def control(self, q, v, controller, params):
        q = _f64(q, (self.n_q,))
        v = _f64(v, (self.n_v,))
        p = _f64(params)
        tau = np.zeros(self.n_v)
        lib().gpo_control(self._h, _d(q), _d(v), controller, _d(p), _d(tau))
        return tau

    def rollout(self, q, v, dt, n_steps, integrator=0, tau=None, controller=0, params=(), history=False):
        q = _f64(q, (self.n_q,)).copy()
        v = _f64(v, (self.n_v,)).copy()
        tau = None if tau is None else _f64(tau, (self.n_v,))
        p = _f64(params if len(params) else [0.0])
        hq = np.zeros((n_steps + 1, self.n_q)) if history else None
        hv = np.zeros((n_steps + 1, self.n_v)) if history else None
        lib().gpo_rollout(self._h, _d(q), _d(v), _d(tau), dt, n_steps, integrator, controller, _d(p),
                          _d(hq), _d(hv))
        if history:
            return q, v, hq, hv
        return q, v

    def simulate(self, q, v, final_time, dt, integrator=0, tau=None, controller=0, params=(), history=True):
        """simulate() (simulate.rs:87): the step count follows the reference's f64 loop."""
        n = int(lib().gpo_simulate_step_count(final_time, dt))
        return self.rollout(q, v, dt, n, integrator, tau, controller, params, history)

    def spring_state_init(self):
        """unregistered SpringContact state, [n_sc, 8] = (registered halfspace, contact xyz, direction xyz, l_rest)"""
        st = np.zeros((self.n_sc, 8))
        lib().gpo_spring_state_init(self._h, _d(st))
        return st

    def step_sc(self, q, v, sc_state, dt, tau=None):
        """step() with SemiImplicitEuler on a mechanism with spring contacts; returns (q, v, sc_state, flags)"""
        q = _f64(q, (self.n_q,)).copy()
        v = _f64(v, (self.n_v,)).copy()
        st = _f64(sc_state, (self.n_sc, 8)).copy()
        tau = None if tau is None else _f64(tau, (self.n_v,))
        flags = lib().gpo_step_sc(self._h, _d(q), _d(v), _d(tau), dt, _d(st))
        return q, v, st, flags

    def free_velocity(self, q, v, dt, tau=None, gravity_enabled=True):
        """Articulated::free_velocity (hybrid/articulated/mod.rs:124-197)"""
        q = _f64(q, (self.n_q,))
        v = _f64(v, (self.n_v,))
        tau = None if tau is None else _f64(tau, (self.n_v,))
        out = np.zeros(self.n_v)
        rc = lib().gpo_free_velocity(self._h, _d(q), _d(v), _d(tau), dt, 1 if gravity_enabled else 0, _d(out))
        if rc != 0:
            raise ArithmeticError("oracle: mass matrix not positive definite")
        return out

    def kinetic_energy(self, q, v):
        q = _f64(q, (self.n_q,))
        v = _f64(v, (self.n_v,))
        return lib().gpo_kinetic_energy(self._h, _d(q), _d(v))

    def gravitational_energy(self, q):
        q = _f64(q, (self.n_q,))
        return lib().gpo_gravitational_energy(self._h, _d(q))

    def spring_energy(self, q):
        q = _f64(q, (self.n_q,))
        return lib().gpo_spring_energy(self._h, _d(q))

    def poses(self, q):
        q = _f64(q, (self.n_q,))
        out = np.zeros((self.n_bodies, 7))
        lib().gpo_poses(self._h, _d(q), _d(out))
        return out

    def body_twists(self, q, v):
        q = _f64(q, (self.n_q,))
        v = _f64(v, (self.n_v,))
        out = np.zeros((self.n_bodies, 6))
        lib().gpo_body_twists(self._h, _d(q), _d(v), _d(out))
        return out

    def supports(self):
        out = np.zeros((self.n_bodies, self.n_bodies), dtype=np.int32)
        lib().gpo_supports(self._h, out.ctypes.data_as(_ip))
        return out

    # ---- many environments (env-major arrays) ----------------------------------
    def batch_dynamics(self, q, v, tau=None, n_threads=None):
        q = _f64(q)
        n = q.shape[0]
        q = q.reshape(n, self.n_q)
        v = _f64(v, (n, self.n_v))
        tau = None if tau is None else _f64(tau, (n, self.n_v))
        vdot = np.zeros((n, self.n_v))
        cf = np.zeros((n, self.n_cp, 3))
        nt = n_threads or os.cpu_count() or 1
        rc = lib().gpo_batch_dynamics(self._h, _d(q), _d(v), _d(tau), n, _d(vdot), _d(cf), nt)
        if rc != 0:
            raise ArithmeticError("oracle: singular mass matrix")
        return vdot, cf

    def batch_rollout(self, q, v, dt, n_steps, integrator=0, tau=None, controller=0, params=(),
                      n_threads=None):
        q = _f64(q)
        n = q.shape[0]
        q = q.reshape(n, self.n_q).copy()
        v = _f64(v, (n, self.n_v)).copy()
        tau = None if tau is None else _f64(tau, (n, self.n_v))
        p = _f64(params if len(params) else [0.0])
        nt = n_threads or os.cpu_count() or 1
        lib().gpo_batch_rollout(self._h, _d(q), _d(v), _d(tau), n, dt, n_steps, integrator, controller,
                                _d(p), nt)
        return q, v


def simple_double_pendulum(m1, m2, l1, l2, q1, q2, q1dot, q2dot):
    out = np.zeros(2)
    lib().gpo_simple_double_pendulum(m1, m2, l1, l2, q1, q2, q1dot, q2dot, _d(out))
    return out


def quat_from_euler(r, p, y):
    out = np.zeros(4)
    lib().gpo_quat_from_euler(r, p, y, _d(out))
    return out


def quat_from_axis_angle(axis, angle):
    out = np.zeros(4)
    a = _f64(axis, (3,))
    lib().gpo_quat_from_axis_angle(_d(a), angle, _d(out))
    return out


def quat_from_scaled_axis(aa):
    out = np.zeros(4)
    a = _f64(aa, (3,))
    lib().gpo_quat_from_scaled_axis(_d(a), _d(out))
    return out


def twist_transform(iso, twist):
    out = np.zeros(6)
    i = _f64(iso, (7,))
    t = _f64(twist, (6,))
    lib().gpo_twist_transform(_d(i), _d(t), _d(out))
    return out


def simulate_step_count(final_time, dt):
    return int(lib().gpo_simulate_step_count(final_time, dt))
