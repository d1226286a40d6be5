"""Developer tool (not a test, not the product): run the device dynamics code compiled for the host
(tools/host_debug.cu, -DGP_HOST_DEBUG) against the oracle, to chase kernel bugs without a GPU."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from gorilla_physics_b200 import _abi
from gorilla_physics_b200.mechanism import Mechanism
from oracle.binding import OracleMechanism
from tests import models
from tests.test_parity_gpu import WORKLOADS, random_states

MODEL_STATES = {"biped": dict(base_t=(0, 0, 0.6), t_jitter=0.1, rpy_jitter=0.3, q_range=0.5),
                "leg": dict(base_t=(0, 0, 0.6), t_jitter=0.1, rpy_jitter=0.3, q_range=0.5),
                "leg_from_foot": dict(base_t=(0, 0, 0.02), t_jitter=0.05, rpy_jitter=0.3, q_range=0.5)}

dbg = C.CDLL(sys.argv[2] if len(sys.argv) > 2 else "/tmp/proto/libgpdbg.so")
dp = C.POINTER(C.c_double)
for n in ("gp_model_create", "gp_mechanism_create", "gp_mechanism_add_halfspace", "gp_mechanism_add_contact_point"):
    getattr(dbg, n).restype = C.c_int
dbg.gp_mechanism_create.argtypes = [C.POINTER(_abi.GpMechanismDesc), C.POINTER(C.c_void_p)]
dbg.gpdbg_dynamics.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, dp, dp]
dbg.gpdbg_dynamics_static.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, dp, dp]
HAVE_PAIRS = hasattr(dbg, "gpdbg_dynamics_pair")  # a -DGP_HOST_PAIRS build (tests/test_device_code_on_host.py)
if HAVE_PAIRS:
    dbg.gpdbg_dynamics_pair.argtypes = [C.c_void_p, dp, dp, dp, dp, dp]


def run(name):
    if name.startswith("random_tree:"):
        # random mechanisms (mixed joint types, fixed joints inside chains, floating joints off other bodies): the
        # run-time-topology instantiation only, the static one has no specialisation for them
        seed = int(name.split(":")[1])
        desc, kw = models.random_tree(2000 + seed, 2 + (5 * seed + 3) % 8), {}
        if desc.n_v == 0:  # (fixed joints only: nothing to solve)
            return 0.0
    elif name.startswith("model:"):
        # the cuboid-built models of the reference (biped / leg / leg_from_foot) on the ground, around it; their
        # compile-time-topology instantiation exists only in a build with that tree's SpecCustom macros
        mech = Mechanism.from_model(name.split(":")[1])
        mech.add_halfspace((0, 0, 1), 0.0)
        desc, kw = mech.desc(), MODEL_STATES[name.split(":")[1]]
    else:
        factory, kw, _ = WORKLOADS[name]
        desc = factory().desc()
    # rebuild the mechanism inside the debug library from the flat description
    keep = {}
    d = _abi.GpMechanismDesc()
    d.n_bodies, d.n_contact_points, d.n_halfspaces = desc.n_bodies, desc.n_contact_points, desc.n_halfspaces
    for f in ("parent", "joint_type", "has_spring", "cp_body"):
        keep[f] = np.ascontiguousarray(getattr(desc, f), dtype=np.int32)
        setattr(d, f, keep[f].ctypes.data_as(_abi.ip))
    for f in ("axis", "init_iso", "moment", "cross_part", "mass", "spring_k", "spring_l", "cp_location", "cp_k",
              "hs_point", "hs_normal", "hs_alpha", "hs_mu"):
        keep[f] = np.ascontiguousarray(getattr(desc, f), dtype=np.float64)
        setattr(d, f, keep[f].ctypes.data_as(dp))
    h = C.c_void_p()
    assert dbg.gp_mechanism_create(C.byref(d), C.byref(h)) == 0
    orc = OracleMechanism(desc)
    q, v = random_states(desc, 8, seed=1234, **kw)
    tau = np.random.default_rng(7).uniform(-1, 1, size=(8, desc.n_v))
    worst = 0.0
    for e in range(8):
        nv = desc.n_v
        vdot = np.zeros(nv); H = np.zeros((nv, nv)); b = np.zeros(nv); cf = np.zeros((max(1, desc.n_contact_points), 3))
        qq = np.ascontiguousarray(q[e]); vv = np.ascontiguousarray(v[e]); tt = np.ascontiguousarray(tau[e])
        dbg.gpdbg_dynamics(h, qq.ctypes.data_as(dp), vv.ctypes.data_as(dp), tt.ctypes.data_as(dp),
                           vdot.ctypes.data_as(dp), H.ctypes.data_as(dp), b.ctypes.data_as(dp), cf.ctypes.data_as(dp))
        vs = np.zeros(nv); Hs = np.zeros((nv, nv)); bs = np.zeros(nv); cfs = np.zeros_like(cf)
        rc = dbg.gpdbg_dynamics_static(h, qq.ctypes.data_as(dp), vv.ctypes.data_as(dp), tt.ctypes.data_as(dp),
                                       vs.ctypes.data_as(dp), Hs.ctypes.data_as(dp), bs.ctypes.data_as(dp), cfs.ctypes.data_as(dp))
        ref = orc.dynamics(q[e], v[e], tau[e], want="all")
        if rc >= 0:
            es = np.abs(vs - ref["vdot"]).max() / max(np.abs(ref["vdot"]).max(), 1e-9)
            ecf = np.abs(cfs - cf).max()
            if e == 0 or es > 1e-9:
                print(f"{name} env {e}: static-topology vdot err {es:.2e}, contact force vs generic {ecf:.1e}")
            worst = max(worst, es)
        if HAVE_PAIRS:
            # the warp-pair mapping: the tree's two halves as two threads that meet at the exchange barriers
            vp = np.zeros(nv)
            mism = C.c_double(0.0)
            rc = dbg.gpdbg_dynamics_pair(h, qq.ctypes.data_as(dp), vv.ctypes.data_as(dp), tt.ctypes.data_as(dp),
                                         vp.ctypes.data_as(dp), C.byref(mism))
            if rc >= 0:
                ep = np.abs(vp - ref["vdot"]).max() / max(np.abs(ref["vdot"]).max(), 1e-9)
                if e == 0 or ep > 1e-9:
                    print(f"{name} env {e}: warp-pair vdot err {ep:.2e}, root copies differ by {mism.value:.1e}")
                worst = max(worst, ep, 1.0 if mism.value != 0.0 else 0.0)
        eM = np.abs(H - ref["mass_matrix"]).max() / np.abs(ref["mass_matrix"]).max()
        eb = np.abs(b - ref["bias"]).max() / max(np.abs(ref["bias"]).max(), 1e-9)
        ev = np.abs(vdot - ref["vdot"]).max() / max(np.abs(ref["vdot"]).max(), 1e-9)
        worst = max(worst, ev)
        if e == 0 or ev > 1e-9:
            print(f"{name} env {e}: M err {eM:.2e}  bias err {eb:.2e}  vdot err {ev:.2e}")
            if eM > 1e-9:
                np.set_printoptions(precision=4, linewidth=200, suppress=True)
                print("H (device code):\n", H, "\nM (oracle):\n", ref["mass_matrix"])
            if eb > 1e-9:
                print("bias dev", b, "\nbias ref", ref["bias"])
    return worst


if __name__ == "__main__":
    names = [sys.argv[1]] if len(sys.argv) > 1 and sys.argv[1] != "all" else list(WORKLOADS) + [f"random_tree:{k}" for k in range(12)]
    if len(names) == 1 and names[0].startswith("random_trees:"):  # random_trees:<first>:<count> in one process
        _, first, count = names[0].split(":")
        names = [f"random_tree:{k}" for k in range(int(first), int(first) + int(count))]
    for n in names:
        print(n, "worst vdot err", run(n))
