#!/usr/bin/env python
"""GPU box: how well conditioned are the joint-space mass matrices the step kernels factorise WITHOUT pivoting?

The reference solves M vdot = tau - c with a dense partial-pivot LU (dynamics.rs:255-276); the kernels use a sparse
L^T D L factorisation (Cholesky family, no pivoting), which is backward stable for symmetric positive definite
matrices whatever their condition number - the error grows with cond(M) for ANY solver, pivoting or not. This
prints, per benchmark workload, cond_2(M) over states drawn from the bench's own distribution (initial states and
the states after a 1 s rollout), the smallest pivot D relative to the largest diagonal entry, and the observed
vdot error against the oracle's LU solve on the same states.

    python tools/cond_mass_matrix.py [workload ...]  > profiles/r2_mass_matrix_conditioning.json
"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from gorilla_physics_b200 import WORKLOADS, MechanismState  # noqa: E402
from oracle.binding import OracleMechanism  # noqa: E402

N = 4096
out = {"_how": "python tools/cond_mass_matrix.py on a B200: gp_batch_mass_matrix of 4096 environments per workload, numpy.linalg.cond "
               "(2-norm) and LDL^T pivots; states: gp_batch_randomize with the bench's distribution, then again after 6000 time steps"}
for name in sys.argv[1:] or [w for w in WORKLOADS if WORKLOADS[w].controller.name == "NONE" and not WORKLOADS[w].settle_steps]:
    w = WORKLOADS[name]
    mech = w.mechanism()
    st = MechanismState(mech, N)
    st.randomize(7, **w.randomize)
    orc = OracleMechanism(mech.desc())
    rec = {}
    for label, steps in (("initial", 0), ("after_6000_steps", 6000)):
        if steps:
            st.step(w.dt, n_steps=steps)
        H, _ = st.mass_matrix()
        q, v = st.state()
        ok = np.isfinite(H).all(axis=(1, 2)) & np.isfinite(q).all(axis=1) & np.isfinite(v).all(axis=1)
        H, q, v = H[ok], q[ok], v[ok]
        cond = np.linalg.cond(H)
        # pivots of the unpivoted LDL^T (numpy Cholesky: D = diag(L)^2)
        L = np.linalg.cholesky(H)
        piv = np.diagonal(L, axis1=1, axis2=2) ** 2
        rel_piv = piv.min(axis=1) / np.diagonal(H, axis1=1, axis2=2).max(axis=1)
        vdot = st.dynamics(tau=None)[ok]
        ref, _ = orc.batch_dynamics(q, v)
        scale = np.maximum(np.abs(ref).max(axis=1), 1e-9)
        err = np.abs(vdot - ref).max(axis=1) / scale
        rec[label] = {"envs": int(ok.sum()), "cond_median": float(np.median(cond)), "cond_p99": float(np.quantile(cond, 0.99)),
                      "cond_max": float(cond.max()), "smallest_pivot_over_largest_diagonal_min": float(rel_piv.min()),
                      "vdot_rel_err_vs_oracle_LU_max": float(err.max()), "vdot_rel_err_p99": float(np.quantile(err, 0.99)),
                      "cond_times_eps_max": float(cond.max() * 2.2e-16)}
    out[name] = rec
    print(name, json.dumps(rec), file=sys.stderr)
print(json.dumps(out, indent=1))
