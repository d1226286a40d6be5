#!/usr/bin/env python
"""Developer tool (GPU box): what the run-time-topology ("generic") kernel costs against the static
specialisation of the same mechanism.   python tools/bench_generic.py [workload ...]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import bench  # noqa: E402
from gorilla_physics_b200 import MechanismState  # noqa: E402
from tests.test_parity_gpu import generic_twin  # noqa: E402

for w in sys.argv[1:] or ["so101_contact", "navbot_contact", "double_pendulum"]:
    n, dt, rnd = bench.WORKLOADS[w]
    n = min(n, 65536)
    for label, mech in (("static", bench.make_mechanism(w)), ("generic", generic_twin(bench.make_mechanism(w)))):
        st = MechanismState(mech, n)
        st.randomize(1, **rnd)
        st.step(dt, n_steps=32)
        st.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            st.step(dt, n_steps=64)
        st.synchronize()
        el = time.perf_counter() - t0
        print(f"{w:16s} {label:8s} {mech.kernel_variant:20s} {n * 64 * 5 / el:.3e} env-steps/s")
