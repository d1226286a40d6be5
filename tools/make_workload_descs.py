#!/usr/bin/env python
"""Freezes the flat mechanism descriptions (gp_mechanism_desc fields) of the benchmark workloads into
tests/golden/workload_descs.json, so that bench.py's reference arm and cpu_baseline leg can hand them to the
oracle WITHOUT loading the product library (a wrong literal in csrc/gp_models.cpp must not be common to both
arms). tests/test_host_cpu.py checks the product's builders against this file, and the SO-101 / navbot entries
against the literals extracted from the reference sources (tests/golden/model_literals.json).

    python tools/make_workload_descs.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

FIELDS = ("parent", "joint_type", "axis", "init_iso", "moment", "cross_part", "mass", "has_spring", "spring_k", "spring_l",
          "cp_body", "cp_location", "cp_k", "hs_point", "hs_normal", "hs_alpha", "hs_mu")


def desc_to_dict(d):
    out = {"n_bodies": int(d.n_bodies), "n_contact_points": int(d.n_contact_points), "n_halfspaces": int(d.n_halfspaces)}
    for f in FIELDS:
        out[f] = np.asarray(getattr(d, f)).tolist()
    return out


def main():
    from gorilla_physics_b200 import WORKLOADS
    out = {"_how": "python tools/make_workload_descs.py (Mechanism.desc() of every gorilla_physics_b200.workloads entry; "
                   "floats are repr-exact)"}
    seen = {}
    for name, w in WORKLOADS.items():
        key = w.factory.__name__ if w.factory.__name__ != "<lambda>" else name
        d = desc_to_dict(w.mechanism().desc())
        # workloads that share a mechanism share an entry
        for k, v in seen.items():
            if v == d:
                out[name] = {"same_as": k}
                break
        else:
            seen[name] = d
            out[name] = d
        del key
    (ROOT / "tests" / "golden" / "workload_descs.json").write_text(json.dumps(out, indent=1))
    print("wrote", len(out) - 1, "workloads")


if __name__ == "__main__":
    main()
