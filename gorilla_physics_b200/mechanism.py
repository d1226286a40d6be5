"""Host-side mirror of the reference's Mechanism / MechanismState / step() / simulate() API,
batched over many independent environments and executed by the sm_100a kernels behind the C ABI.

Reference surface mirrored (one-for-all/gorilla-physics):
    MechanismState::new / update / add_halfspace / add_contact_point / kinetic_energy /
    gravitational_energy / spring_energy / poses           src/mechanism.rs:62-417
    step(state, dt, tau, integrator)                       src/simulate.rs:20-83
    simulate(state, final_time, dt, control_fn, integ)     src/simulate.rs:87-112
    dynamics_continuous(state, tau)                        src/dynamics.rs:322-364
    enum Integrator                                        src/integrators.rs:17-23

Arrays are numpy, env-major: q is [n_envs, n_q] in the reference's flat packing
(floating joint: qx,qy,qz,qw,tx,ty,tz), v / tau / vdot are [n_envs, n_v]. Host pointers may
also be given as integer addresses (e.g. a pinned torch tensor's data_ptr()).
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Callable, Optional, Sequence

import numpy as np

from . import _abi
from ._abi import GpMechanismDesc, GpStateDist, check, dp, ip, lib
from .desc import MechanismDesc


class Integrator(enum.IntEnum):
    """enum Integrator, reference src/integrators.rs:17-23 (same order)."""
    SemiImplicitEuler = 0
    RungeKutta2 = 1
    RungeKutta4 = 2
    VelocityStepping = 3      # needs the SOCP contact solver: GP_ERR_UNSUPPORTED
    CCDVelocityStepping = 4   # idem


class Controller(enum.IntEnum):
    """Closed-form controllers evaluated inside the step kernel (enum gp_controller)."""
    NONE = 0
    SO101_PD = 1           # reference control/so101_control.rs:12-34, params [kp, kd, clamp]
    ACROBOT_SWINGUP = 2    # reference control/swingup.rs:9-69, params [m, l]
    CARTPOLE_SWINGUP = 3   # reference control/swingup.rs:76-110, params [m_c, m_p, l]
    HOPPER_1D = 4          # reference control/energy_control.rs:24-101 (stateful), params
    #                        [k_spring, h_setpoint, body_leg_length, leg_foot_length]
    PENDULUM_GRAVITY_INVERSION = 5   # reference control/mod.rs:57-67 (single revolute pendulum, no params)
    PENDULUM_ENERGY_SHAPING = 6      # reference control/mod.rs:78-96
    PENDULUM_SWINGUP_BALANCE = 7     # reference control/mod.rs:98-105
    QUADRUPED_TROT = 8     # reference control/quadruped_control.rs:10-266 (stateful, 9 values), params [dt, target_x, default_foot_z]


class KernelMode(enum.IntEnum):
    """enum gp_kernel_mode: which device kernels a mechanism runs on."""
    AUTO = 0      # shipped specialisation, else run-time-compiled (NVRTC), else generic
    GENERIC = 1   # always the run-time-topology kernel
    JIT = 2       # always run-time-compiled, even where a shipped specialisation matches
    SHIPPED = 3   # shipped specialisation, else generic: never compile at run time


# gp_mechanism_precompile kinds
JIT_STEP_SIE, JIT_STEP_RK, JIT_DYNAMICS, JIT_ENERGY, JIT_STEP_TAU_SEQ = 1, 2, 4, 8, 16


def jit_available() -> bool:
    """NVRTC could be loaded (and GP_JIT is not 0): unknown trees get their own compiled kernels."""
    return bool(lib().gp_jit_available())


def jit_cache_dir() -> str:
    buf = C.create_string_buffer(4096)
    lib().gp_jit_cache_dir(buf, len(buf))
    return buf.value.decode()


def _f64(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return a.reshape(shape) if shape is not None else a


def _ptr(a):
    """numpy array / int address / None -> c_void_p"""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return C.c_void_p(a.ctypes.data)


class Mechanism:
    """One mechanism description living behind a gp_mechanism handle."""

    def __init__(self, handle: C.c_void_p):
        self._h = handle

    # ---- construction -------------------------------------------------------
    @classmethod
    def from_desc(cls, desc: MechanismDesc, kernel: "KernelMode | None" = None) -> "Mechanism":
        """MechanismState::new(treejoints, bodies), reference mechanism.rs:62."""
        keep = {
            "parent": np.ascontiguousarray(desc.parent, dtype=np.int32),
            "joint_type": np.ascontiguousarray(desc.joint_type, dtype=np.int32),
            "axis": _f64(desc.axis), "init_iso": _f64(desc.init_iso), "moment": _f64(desc.moment),
            "cross_part": _f64(desc.cross_part), "mass": _f64(desc.mass),
            "has_spring": np.ascontiguousarray(desc.has_spring, dtype=np.int32),
            "spring_k": _f64(desc.spring_k), "spring_l": _f64(desc.spring_l),
            "cp_body": np.ascontiguousarray(desc.cp_body, dtype=np.int32),
            "cp_location": _f64(desc.cp_location), "cp_k": _f64(desc.cp_k),
            "hs_point": _f64(desc.hs_point), "hs_normal": _f64(desc.hs_normal),
            "hs_alpha": _f64(desc.hs_alpha), "hs_mu": _f64(desc.hs_mu),
            "armature": _f64(desc.armature),
            "sc_body": np.ascontiguousarray(desc.sc_body, dtype=np.int32),
            "sc_l_rest": _f64(desc.sc_l_rest), "sc_direction": _f64(desc.sc_direction), "sc_k": _f64(desc.sc_k),
        }
        d = GpMechanismDesc()
        d.n_bodies = desc.n_bodies
        d.n_contact_points = desc.n_contact_points
        d.n_halfspaces = desc.n_halfspaces
        d.n_spring_contacts = desc.n_spring_contacts
        for name, arr in keep.items():
            setattr(d, name, arr.ctypes.data_as(ip if arr.dtype == np.int32 else dp))
        h = C.c_void_p()
        check(lib().gp_mechanism_create(C.byref(d), C.byref(h)))
        m = cls(h)
        if kernel is not None:
            m.set_kernel_mode(kernel)
        return m

    @classmethod
    def from_model(cls, name: str, params: Sequence[float] = ()) -> "Mechanism":
        """The reference's builders (helpers.rs, builders/mod.rs, navbot_builder.rs); see
        gp_model_create in include/gorilla_b200.h for names and parameter lists."""
        p = _f64(list(params)) if len(params) else None
        h = C.c_void_p()
        check(lib().gp_model_create(name.encode(), None if p is None else p.ctypes.data_as(dp), len(params),
                                    C.byref(h)))
        return cls(h)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().gp_mechanism_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_kernel_mode(self, mode: "KernelMode") -> "Mechanism":
        """gp_mechanism_set_kernel_mode: shipped / run-time-compiled / run-time-topology kernels."""
        check(lib().gp_mechanism_set_kernel_mode(self._h, int(mode)))
        return self

    def precompile(self, kinds: int = JIT_STEP_SIE | JIT_DYNAMICS) -> int:
        """Compile this mechanism's run-time-specialised kernels into the on-disk cache (needs no GPU).
        Returns how many were not cached yet."""
        n = C.c_int(0)
        check(lib().gp_mechanism_precompile(self._h, int(kinds), C.byref(n)))
        return n.value

    # ---- queries ---------------------------------------------------------------
    @property
    def n_bodies(self):
        return lib().gp_mechanism_n_bodies(self._h)

    @property
    def n_q(self):
        return lib().gp_mechanism_n_q(self._h)

    @property
    def n_v(self):
        return lib().gp_mechanism_n_v(self._h)

    @property
    def n_contact_points(self):
        return lib().gp_mechanism_n_contact_points(self._h)

    @property
    def n_halfspaces(self):
        return lib().gp_mechanism_n_halfspaces(self._h)

    @property
    def kernel_variant(self) -> str:
        return lib().gp_mechanism_kernel_variant(self._h).decode()

    def desc(self) -> MechanismDesc:
        """Copy of the flat description (contact points in body-major order)."""
        d = GpMechanismDesc()
        check(lib().gp_mechanism_get_desc(self._h, C.byref(d)))
        nb, nc, nh = d.n_bodies, d.n_contact_points, d.n_halfspaces

        def arr(ptr, n, dtype):
            if n == 0:
                return np.zeros(0, dtype=dtype)
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)

        ns = d.n_spring_contacts
        md = MechanismDesc.from_arrays(
            n_bodies=nb, parent=arr(d.parent, nb, np.int32), joint_type=arr(d.joint_type, nb, np.int32),
            axis=arr(d.axis, 3 * nb, np.float64), init_iso=arr(d.init_iso, 7 * nb, np.float64),
            moment=arr(d.moment, 9 * nb, np.float64), cross_part=arr(d.cross_part, 3 * nb, np.float64),
            mass=arr(d.mass, nb, np.float64), has_spring=arr(d.has_spring, nb, np.int32),
            spring_k=arr(d.spring_k, nb, np.float64), spring_l=arr(d.spring_l, nb, np.float64),
            n_contact_points=nc, cp_body=arr(d.cp_body, nc, np.int32),
            cp_location=arr(d.cp_location, 3 * nc, np.float64), cp_k=arr(d.cp_k, nc, np.float64),
            n_halfspaces=nh, hs_point=arr(d.hs_point, 3 * nh, np.float64),
            hs_normal=arr(d.hs_normal, 3 * nh, np.float64), hs_alpha=arr(d.hs_alpha, nh, np.float64),
            hs_mu=arr(d.hs_mu, nh, np.float64), armature=arr(d.armature, nb, np.float64))
        for k in range(ns):
            md.add_spring_contact(int(d.sc_body[k]), float(d.sc_l_rest[k]),
                                  [d.sc_direction[3 * k], d.sc_direction[3 * k + 1], d.sc_direction[3 * k + 2]], float(d.sc_k[k]))
        return md

    def supports(self) -> np.ndarray:
        """supports[j-1][i-1] == 1 iff joint j supports body i (reference mechanism.rs:118-125)."""
        nb = self.n_bodies
        out = np.zeros((nb, nb), dtype=np.int32)
        check(lib().gp_mechanism_supports(self._h, out.ctypes.data_as(ip)))
        return out

    # ---- add_halfspace / add_contact_point, reference mechanism.rs:379-392 ---------
    def add_halfspace(self, normal, distance: float, alpha: float = 0.9, mu: float = 0.5):
        """HalfSpace::new / new_with_params (reference collision/halfspace.rs:15-37)."""
        n = _f64(normal, (3,))
        point = _f64(n * distance)
        check(lib().gp_mechanism_add_halfspace(self._h, point.ctypes.data_as(dp), n.ctypes.data_as(dp), alpha, mu))

    def add_spring_contact(self, body: int, l_rest: float, direction, k: float):
        """SpringContact::new + add_spring_contact (reference contact.rs:83-94, mechanism.rs:394-401)"""
        d = _f64(direction, (3,))
        check(lib().gp_mechanism_add_spring_contact(self._h, body, l_rest, d.ctypes.data_as(dp), k))

    @property
    def n_spring_contacts(self):
        return lib().gp_mechanism_n_spring_contacts(self._h)

    def add_contact_point(self, body: int, location, k: float = 50e3):
        loc = _f64(location, (3,))
        check(lib().gp_mechanism_add_contact_point(self._h, body, loc.ctypes.data_as(dp), k))


class MechanismState:
    """n_envs independent copies of one MechanismState (reference mechanism.rs:41-58), resident in
    HBM as structure-of-arrays planes. Single-owner, like `&mut MechanismState`."""

    def __init__(self, mechanism: Mechanism, n_envs: int = 1, device: int = 0):
        self.mechanism = mechanism
        self.n_envs = int(n_envs)
        self.n_q, self.n_v = mechanism.n_q, mechanism.n_v
        h = C.c_void_p()
        check(lib().gp_batch_create(mechanism._h, self.n_envs, device, C.byref(h)))
        self._h = h
        self.device = device
        self._owned = True

    @classmethod
    def _borrowed(cls, mechanism: Mechanism, handle, owner) -> "MechanismState":
        """view of a gp_batch that something else owns (a shard of a gp_sharded): never destroyed from here;
        `owner` is kept alive as long as the view is"""
        st = cls.__new__(cls)
        st.mechanism = mechanism
        st._h = C.c_void_p(handle)
        st.n_envs = int(lib().gp_batch_n_envs(st._h))
        st.n_q, st.n_v = mechanism.n_q, mechanism.n_v
        st.device = int(lib().gp_batch_device(st._h))
        st._owned = False
        st._owner = owner
        return st

    def __del__(self):
        try:
            if getattr(self, "_h", None) and getattr(self, "_owned", False):
                lib().gp_batch_destroy(self._h)
            self._h = None
        except Exception:
            pass

    # ---- raw handles (zero-copy interop, event timing) ----------------------------
    @property
    def ld(self) -> int:
        return lib().gp_batch_ld(self._h)

    @property
    def q_ptr(self) -> int:
        return lib().gp_batch_q_device(self._h)

    @property
    def v_ptr(self) -> int:
        return lib().gp_batch_v_device(self._h)

    @property
    def tau_ptr(self) -> int:
        return lib().gp_batch_tau_device(self._h)

    @property
    def stream(self) -> int:
        return lib().gp_batch_stream(self._h)

    @property
    def step_lanes(self) -> int:
        """threads per environment of a SemiImplicitEuler step of this batch: 1, or 2 (warp pairs: small batches of
        trees that have two halves)"""
        return lib().gp_batch_step_lanes(self._h)

    @property
    def launch_count(self) -> int:
        return lib().gp_batch_launch_count(self._h)

    def synchronize(self):
        check(lib().gp_batch_sync(self._h))

    # ---- state ---------------------------------------------------------------------
    def update(self, q=None, v=None):
        """MechanismState::update(q, v), reference mechanism.rs:209. Rows are environments; a single
        row is broadcast to every environment."""
        qa = None if q is None else self._rows(q, self.n_q)
        va = None if v is None else self._rows(v, self.n_v)
        check(lib().gp_batch_set_state(self._h, _ptr(qa), _ptr(va)))

    def update_from_host_ptr(self, q_addr: Optional[int], v_addr: Optional[int]):
        check(lib().gp_batch_set_state(self._h, _ptr(q_addr), _ptr(v_addr)))

    def _rows(self, a, k):
        if isinstance(a, (int, np.integer)):
            return a
        a = _f64(a)
        if a.ndim == 1:
            a = np.broadcast_to(a.reshape(1, k), (self.n_envs, k))
        return np.ascontiguousarray(a.reshape(self.n_envs, k))

    def state(self):
        q = np.empty((self.n_envs, self.n_q))
        v = np.empty((self.n_envs, self.n_v))
        check(lib().gp_batch_get_state(self._h, _ptr(q), _ptr(v)))
        return q, v

    @property
    def q(self):
        return self.state()[0]

    @property
    def v(self):
        return self.state()[1]

    def set_tau(self, tau):
        """None -> zero torques (reference simulate.rs:27-48)."""
        ta = None if tau is None else self._rows(tau, self.n_v)
        check(lib().gp_batch_set_tau(self._h, _ptr(ta)))

    def spring_contact_state(self):
        """[n_envs, NS, 8] = (registered halfspace, contact xyz, direction xyz, l_rest) per spring contact"""
        ns = self.mechanism.n_spring_contacts
        out = np.zeros((self.n_envs, ns, 8))
        if ns:
            check(lib().gp_batch_get_spring_contact_state(self._h, _ptr(out)))
        return out

    def set_spring_contact_state(self, state=None):
        sa = None if state is None else np.ascontiguousarray(np.asarray(state, dtype=np.float64).reshape(self.n_envs, -1))
        check(lib().gp_batch_set_spring_contact_state(self._h, _ptr(sa)))

    def set_controller_state(self, state=None, k: int = 2):
        """per-environment controller state, k values each: Controller.HOPPER_1D (leg_length_setpoint,
        v_vertical_prev), Controller.QUADRUPED_TROT k = 9 (ticks + 1, four feet's (x, z)); None resets to 0
        (a fresh controller)."""
        sa = None if state is None else self._rows(state, k)
        check(lib().gp_batch_set_controller_state_n(self._h, _ptr(sa), int(k)))

    def controller_state(self, k: int = 2):
        out = np.empty((self.n_envs, k))
        check(lib().gp_batch_get_controller_state_n(self._h, _ptr(out), int(k)))
        return out

    def randomize(self, seed: int, q_range=(-1.0, 1.0), v_range=(-1.0, 1.0), base_t=(0.0, 0.0, 0.0),
                  t_jitter=(0.0, 0.0, 0.0), rpy_jitter=0.0, base_v=(0.0,) * 6, v_jitter=0.0):
        d = GpStateDist()
        d.q_lo, d.q_hi = q_range
        d.v_lo, d.v_hi = v_range
        d.base_t = (C.c_double * 3)(*base_t)
        d.t_jitter = (C.c_double * 3)(*t_jitter)
        d.rpy_jitter = rpy_jitter
        d.base_v = (C.c_double * 6)(*base_v)
        d.v_jitter = v_jitter
        check(lib().gp_batch_randomize(self._h, seed, C.byref(d)))

    # ---- dynamics ------------------------------------------------------------------
    def dynamics(self, tau="keep", contact_forces: bool = False):
        """dynamics_continuous, reference dynamics.rs:322-364. Returns vdot [n_envs, n_v]
        (and the per-contact-point world-frame forces [n_envs, NC, 3])."""
        if not (isinstance(tau, str) and tau == "keep"):
            self.set_tau(tau)
        vdot = np.empty((self.n_envs, self.n_v))
        nc = self.mechanism.n_contact_points
        cf = np.zeros((self.n_envs, nc, 3)) if contact_forces else None
        check(lib().gp_batch_dynamics(self._h, _ptr(vdot), _ptr(cf) if (cf is not None and nc) else None))
        return (vdot, cf) if contact_forces else vdot

    def free_velocity(self, dt: float, gravity_enabled: bool = True, tau="keep"):
        """Articulated::free_velocity (reference hybrid/articulated/mod.rs:124-197): v + M^-1 (tau - c) dt,
        armature on the diagonal of M, no contact forces. Does not advance the state."""
        if not (isinstance(tau, str) and tau == "keep"):
            self.set_tau(tau)
        out = np.empty((self.n_envs, self.n_v))
        check(lib().gp_batch_free_velocity(self._h, dt, 1 if gravity_enabled else 0, _ptr(out)))
        return out

    def mass_matrix(self):
        """mass_matrix (reference mechanism.rs:637-696) and dynamics_bias (dynamics.rs:233-251)."""
        M = np.empty((self.n_envs, self.n_v, self.n_v))
        c = np.empty((self.n_envs, self.n_v))
        check(lib().gp_batch_mass_matrix(self._h, _ptr(M), _ptr(c)))
        return M, c

    def step(self, dt: float, tau="keep", integrator: Integrator = Integrator.SemiImplicitEuler,
             n_steps: int = 1, controller: Controller = Controller.NONE, ctrl_params: Sequence[float] = ()):
        """step() (reference simulate.rs:20-83) applied n_steps times in one kernel launch. Asynchronous."""
        if not (isinstance(tau, str) and tau == "keep"):
            self.set_tau(tau)
        p = _f64(list(ctrl_params)) if len(ctrl_params) else None
        check(lib().gp_batch_step(self._h, dt, int(integrator), int(n_steps), int(controller),
                                  None if p is None else p.ctypes.data_as(dp), len(ctrl_params)))

    def step_tau_sequence(self, dt: float, tau_seq, integrator: Integrator = Integrator.SemiImplicitEuler,
                          n_steps: Optional[int] = None):
        """n_steps of step() with a different torque vector before each: the reference's per-step control closure
        (simulate.rs:87-112) for torques known ahead of the rollout. tau_seq: [n_steps, n_envs, n_v] (numpy, or the
        address of such a host buffer together with n_steps). One launch per block of steps; returns when done."""
        if isinstance(tau_seq, (int, np.integer)):
            if n_steps is None:
                raise ValueError("n_steps is required with a raw address")
            ptr = C.c_void_p(int(tau_seq))
        else:
            a = _f64(tau_seq)
            if a.ndim == 2 and self.n_envs == 1:
                a = a[:, None, :]
            if a.ndim != 3 or a.shape[1:] != (self.n_envs, self.n_v):
                raise ValueError(f"tau_seq must be [n_steps, {self.n_envs}, {self.n_v}]")
            n_steps = a.shape[0]
            ptr = _ptr(a)
        check(lib().gp_batch_step_tau_sequence(self._h, dt, int(integrator), int(n_steps), ptr))

    def step_tau_sequence_device(self, dt: float, tau_seq_dev_ptr: int, n_steps: int,
                                 integrator: Integrator = Integrator.SemiImplicitEuler):
        """as step_tau_sequence, the sequence already on the device as [n_steps][n_v][ld] planes; asynchronous"""
        check(lib().gp_batch_step_tau_sequence_device(self._h, dt, int(integrator), int(n_steps),
                                                      C.c_void_p(int(tau_seq_dev_ptr))))

    def simulate(self, final_time: float, dt: float, q, v, tau=None,
                 integrator: Integrator = Integrator.SemiImplicitEuler, controller: Controller = Controller.NONE,
                 ctrl_params: Sequence[float] = (), history: bool = False):
        """simulate() (reference simulate.rs:87-112) through host buffers: H2D, rollout, D2H in one call.
        q / v are updated in place when they are float64 C-contiguous arrays (or raw addresses).
        Returns (n_steps, qs, vs) with qs/vs the [n_steps+1, n_envs, ·] histories when history=True."""
        qa, va = self._rows(q, self.n_q), self._rows(v, self.n_v)
        ta = None if tau is None else self._rows(tau, self.n_v)
        p = _f64(list(ctrl_params)) if len(ctrl_params) else None
        n_steps = C.c_int64()
        hq = hv = None
        if history:
            n = int(lib().gp_simulate_step_count(final_time, dt))
            if n < 0:
                raise ValueError(f"simulate: final_time={final_time!r} / dt={dt!r} gives no countable number of steps")
            hq = np.empty((n + 1, self.n_envs, self.n_q))
            hv = np.empty((n + 1, self.n_envs, self.n_v))
        check(lib().gp_batch_simulate(self._h, _ptr(qa), _ptr(va), _ptr(ta), final_time, dt, int(integrator),
                                      int(controller), None if p is None else p.ctypes.data_as(dp),
                                      len(ctrl_params), C.byref(n_steps), _ptr(hq), _ptr(hv)))
        if history:
            return n_steps.value, hq, hv
        return n_steps.value, qa, va

    # ---- diagnostics ------------------------------------------------------------------
    def energies(self):
        ke, pe, se = (np.empty(self.n_envs) for _ in range(3))
        check(lib().gp_batch_energy(self._h, _ptr(ke), _ptr(pe), _ptr(se)))
        return ke, pe, se

    def kinetic_energy(self):
        return self.energies()[0]

    def gravitational_energy(self):
        return self.energies()[1]

    def spring_energy(self):
        return self.energies()[2]

    def reduce_diagnostics(self, comm: "Optional[Communicator]" = None) -> np.ndarray:
        """[sum KE, sum PE, sum spring energy, flagged environments] of this batch, all-reduced over the ranks of
        `comm` inside the library (NCCL on the batch's stream, gp_batch_reduce_diagnostics); without a
        communicator: this batch's own sums. The path's only exchange, end of rollout."""
        out = np.zeros(4)
        check(lib().gp_batch_reduce_diagnostics(self._h, comm._h if comm is not None else None, out.ctypes.data_as(dp)))
        return out

    def energy_sums_device(self, out_dev_ptr: int):
        check(lib().gp_batch_energy_sums_device(self._h, C.c_void_p(int(out_dev_ptr))))

    def poses(self):
        out = np.empty((self.n_envs, self.mechanism.n_bodies, 7))
        check(lib().gp_batch_poses(self._h, _ptr(out)))
        return out

    def status(self):
        out = np.empty(self.n_envs, dtype=np.uint32)
        check(lib().gp_batch_status(self._h, _ptr(out)))
        return out

    def clear_status(self):
        check(lib().gp_batch_clear_status(self._h))


# ---- free functions with the reference's names -----------------------------------------------
def step(state: MechanismState, dt: float, tau=None, integrator: Integrator = Integrator.SemiImplicitEuler):
    """step(state, dt, tau, integrator) -> (q, v), reference simulate.rs:20-83. An empty / None tau
    means zero torques (simulate.rs:27-48)."""
    if tau is not None and np.size(tau) == 0:
        tau = None
    if tau is not None and np.asarray(tau).shape[-1] != state.n_v:
        # reference: assert_eq!(tau.len(), state.v.len(), ...) simulate.rs:39-45
        raise ValueError(f"joint torques vector length {np.asarray(tau).shape[-1]} and joint velocity "
                         f"vector length {state.n_v} differ!")
    state.step(dt, tau=tau, integrator=integrator, n_steps=1)
    return state.state()


def simulate(state: MechanismState, final_time: float, dt: float,
             control_fn: Optional[Callable[[MechanismState], Optional[np.ndarray]]] = None,
             integrator: Integrator = Integrator.SemiImplicitEuler):
    """simulate(state, final_time, dt, control_fn, integrator) -> (qs, vs) including the initial state,
    reference simulate.rs:87-112. control_fn runs on the host between steps (the reference's closure);
    use MechanismState.step(controller=...) for the in-kernel controllers."""
    q, v = state.state()
    qs, vs = [q], [v]
    t = 0.0
    while t < final_time:
        tau = control_fn(state) if control_fn is not None else None
        q, v = step(state, dt, tau, integrator)
        qs.append(q)
        vs.append(v)
        t += dt
    return np.stack(qs), np.stack(vs)


def measure_fp64_peak(device: int = 0, seconds: float = 1.0) -> float:
    out = C.c_double()
    check(lib().gp_measure_fp64_peak(device, seconds, C.byref(out)))
    return out.value


def measure_fp64_peak_trace(device: int = 0, seconds: float = 1.0, max_samples: int = 4096):
    """DFMA-chain probe, one (seconds of load so far, TFLOP/s) sample per launch: the burst figure at the start,
    the sustained one at the end (gp_measure_fp64_peak_trace)."""
    t = np.zeros(max_samples)
    f = np.zeros(max_samples)
    n = C.c_int()
    dp = C.POINTER(C.c_double)
    check(lib().gp_measure_fp64_peak_trace(device, seconds, t.ctypes.data_as(dp), f.ctypes.data_as(dp), max_samples,
                                           C.byref(n)))
    return t[:n.value], f[:n.value]


def nccl_available() -> bool:
    return bool(lib().gp_nccl_available())


class Communicator:
    """NCCL communicator owned by the library (gp_comm_*), one rank per GPU. Rank 0 calls
    Communicator.unique_id() and hands the 128 bytes to the other ranks by its own means."""

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().gp_comm_unique_id(buf))
        return buf.raw

    def __init__(self, rank: int, world: int, unique_id: bytes, device: int):
        if len(unique_id) != 128:
            raise ValueError("the unique id is 128 bytes")
        h = C.c_void_p()
        check(lib().gp_comm_create(int(rank), int(world), C.create_string_buffer(unique_id, 128), int(device), C.byref(h)))
        self._h = h
        self.rank, self.world = rank, world

    def close(self):
        if getattr(self, "_h", None):
            lib().gp_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
