#!/usr/bin/env python
"""Developer tool (GPU box): the static specialisation of a mechanism against its twin (a massless fixed leaf
appended, so that no shipped specialisation matches) on the run-time-topology ("generic") kernel and on a kernel
compiled at run time for the twin's tree (NVRTC, gp_jit.cpp).   python tools/bench_generic.py [workload ...]
One JSON line per (workload, flavour) -> stdout."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from gorilla_physics_b200 import WORKLOADS, KernelMode, MechanismState, jit_available  # noqa: E402
from tests.test_parity_gpu import generic_twin  # noqa: E402

for w in sys.argv[1:] or ["so101_contact", "navbot_contact", "double_pendulum"]:
    wl = WORKLOADS[w]
    n, dt, rnd = min(wl.n_envs, 65536), wl.dt, wl.randomize
    flavours = [("static", wl.mechanism()), ("generic", generic_twin(wl.mechanism()))]
    if jit_available():
        flavours.append(("jit", generic_twin(wl.mechanism(), KernelMode.JIT)))
    base = None
    for label, mech in flavours:
        st = MechanismState(mech, n)
        st.randomize(1, **rnd)
        st.step(dt, n_steps=32)
        st.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            st.step(dt, n_steps=64)
        st.synchronize()
        el = time.perf_counter() - t0
        rate = n * 64 * 5 / el
        base = base or rate
        print(json.dumps({"workload": w, "n_envs": n, "flavour": label, "kernel": mech.kernel_variant,
                          "env_steps_per_sec": rate, "static_over_this": base / rate}), flush=True)
