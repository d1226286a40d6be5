#!/usr/bin/env python
"""Writes tests/golden/oracle_frozen.json: frozen oracle outputs (vdot, contact forces, a 50-step
semi-implicit-Euler rollout) for the models the reference holds no known-answer test for. Run it only
after tests/test_oracle_independent.py's cross-check against the independent derivation passes; the
file then keeps the oracle from drifting.   python tools/make_oracle_golden.py"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle.binding import OracleMechanism  # noqa: E402
from tests.test_oracle_independent import CASES, states  # noqa: E402

FROZEN = ("so101", "so101_contact", "navbot", "navbot_contact", "quadruped", "hopper_1d", "rimless_wheel")
DT = {"so101": 1 / 6000, "so101_contact": 1 / 6000, "navbot": 1 / 6000, "navbot_contact": 1 / 6000,
      "quadruped": 1 / 3000, "hopper_1d": 1 / 500, "rimless_wheel": 1 / 600}


def main():
    out = {"_how": "python tools/make_oracle_golden.py (oracle/gp_oracle.cpp, g++ -O2 -ffp-contract=off)", "cases": {}}
    for name in FROZEN:
        factory, kw, _ = CASES[name]
        desc = factory()
        orc = OracleMechanism(desc)
        q, v, tau = states(desc, 3, seed=2026, **kw)
        samples = []
        for e in range(3):
            a = orc.dynamics(q[e], v[e], tau[e], want="all")
            samples.append({"q": q[e].tolist(), "v": v[e].tolist(), "tau": tau[e].tolist(),
                            "vdot": a["vdot"].tolist(), "contact_forces": a["contact_forces"].tolist()})
        q1, v1 = orc.rollout(q[0], v[0], DT[name], 50)
        out["cases"][name] = {"samples": samples,
                              "rollout": {"q0": q[0].tolist(), "v0": v[0].tolist(), "dt": DT[name], "steps": 50,
                                          "q1": q1.tolist(), "v1": v1.tolist()}}
    path = ROOT / "tests" / "golden" / "oracle_frozen.json"
    path.write_text(json.dumps(out, indent=0))
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
