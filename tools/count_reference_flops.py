#!/usr/bin/env python
"""Floating-point operations per environment time step of the REFERENCE's formulation, counted by running
the oracle compiled with a counting scalar type (oracle/gp_oracle_count.cpp; SURVEY.md §8d asks for this
"CountingDouble" figure), next to what the CUDA kernels execute (ncu, profiles/flop_counts.json) and the
algorithmic bytes. Writes profiles/roofline.json. CPU only:   python tools/count_reference_flops.py"""
import ctypes as C
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import binding  # noqa: E402


class _Proxy:
    """the counting library exports gpc_* where the oracle exports gpo_*"""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        return getattr(self._real, name.replace("gpo_", "gpc_", 1))


def main():
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "count"], check=True)
    real = C.CDLL(str(ROOT / "oracle" / "libgp_oracle_count.so"))
    real.gpc_reset_counters.restype = None
    real.gpc_read_counters.restype = None
    real.gpc_read_counters.argtypes = [C.POINTER(C.c_longlong)]
    binding._lib = binding.configure(_Proxy(real))  # OracleMechanism now drives the counting build

    from gorilla_physics_b200 import WORKLOADS  # workload table and mechanisms of the benchmark
    from tests.test_parity_gpu import random_states

    executed = json.loads((ROOT / "profiles" / "flop_counts.json").read_text())
    state_kw = {
        "rimless_wheel": dict(base_t=(0, 0, -10.5), t_jitter=0.5, rpy_jitter=0.3),
        "quadruped": dict(base_t=(0, 0, 0.8), t_jitter=0.01, rpy_jitter=0.1, q_range=0.2),
        "navbot_contact": dict(base_t=(0, 0, 0.075), t_jitter=0.01, rpy_jitter=0.1, q_range=0.2),
        "hopper_1d": dict(base_t=(0, 0, 2.5), t_jitter=1.0, rpy_jitter=0.0, q_range=0.0),
        "biped": dict(base_t=(0, 0, 0.72), t_jitter=0.02, rpy_jitter=0.1, q_range=0.3),
    }
    out = {"_how": "python tools/count_reference_flops.py: oracle/gp_oracle_count.cpp (the oracle with a counting scalar "
                   "type, reference operation order, no FMA) stepping 16 environments of each bench workload for 2000 "
                   "semi-implicit-Euler steps from the parity tests' seeded states; counts are per environment time step, "
                   "every add/sub, mul, div, sqrt, sin/cos, pow = 1. executed_* = what the CUDA step kernel executes "
                   "(ncu, profiles/flop_counts.json, 2*DFMA + DADD + DMUL averaged over the bench's timed launches). "
                   "bytes = (n_q + n_v) * 8 read + written per environment per LAUNCH (state stays in registers across "
                   "the fused steps); per env-step = / 128 fused steps.",
           "workloads": {}}
    for name, wl in WORKLOADS.items():
        if wl.controller.name != "NONE" or wl.settle_steps:
            continue  # the count is per mechanism: variants of a counted workload add nothing
        n_envs, dt = wl.n_envs, wl.dt
        mech = wl.mechanism()
        desc = mech.desc()
        orc = binding.OracleMechanism(desc)
        n, steps = 16, 2000
        q, v = random_states(desc, n, seed=1, **state_kw.get(name, {}))
        real.gpc_reset_counters()
        orc.batch_rollout(q, v, dt, steps, n_threads=1)  # one thread: the counters are thread-local
        cnt = (C.c_longlong * 6)()
        real.gpc_read_counters(cnt)
        per = [c / (n * steps) for c in cnt]
        ref_flop = sum(per)
        exe = executed["flop_per_env_step"].get(name)
        bytes_launch = (desc.n_q + desc.n_v) * 8 * 2
        out["workloads"][name] = {
            "n_envs_per_gpu": n_envs, "n_q": desc.n_q, "n_v": desc.n_v, "n_contact_points": int(desc.n_contact_points),
            "reference_formulation_flop_per_env_step": round(ref_flop, 1),
            "reference_formulation_ops": {"add": round(per[0], 1), "mul": round(per[1], 1), "div": round(per[2], 2),
                                          "sqrt": round(per[3], 2), "sincos": round(per[4], 2), "pow": round(per[5], 2)},
            "executed_flop_per_env_step": exe,
            "executed_fp64_inst_per_env_step": executed["fp64_inst_per_env_step"].get(name),
            "executed_over_reference": round(exe / ref_flop, 3) if exe else None,
            "algorithmic_bytes_per_env_per_launch": bytes_launch,
            "algorithmic_bytes_per_env_step_at_128_fused": bytes_launch / 128.0,
            "dram_bytes_per_launch_ncu": executed["dram_bytes_per_launch"].get(name),
        }
        print(name, out["workloads"][name]["reference_formulation_flop_per_env_step"], exe)
    (ROOT / "profiles" / "roofline.json").write_text(json.dumps(out, indent=1))
    print("wrote profiles/roofline.json")


if __name__ == "__main__":
    main()
