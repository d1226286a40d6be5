#!/usr/bin/env python
"""Prints the make variables that give a mechanism its own static kernels (csrc/Makefile CUSTOM_*,
gp_topology.cuh SpecCustom).   python tools/custom_topo.py <model name> [model parameters ...]
From Python: custom_topo_vars(mechanism.desc(), "my_robot")."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def custom_topo_vars(desc, name):
    parents = ",".join(str(int(p) - 1) for p in desc.parent)
    joints = ",".join(str(int(t)) for t in desc.joint_type)
    axes = ",".join("1" if (int(t) in (1, 2) and tuple(float(x) for x in a) == (0.0, 0.0, 1.0)) else "0"
                    for t, a in zip(desc.joint_type, desc.axis))
    return (f"CUSTOM_NB={desc.n_bodies} CUSTOM_PARENTS={parents} CUSTOM_JOINTS={joints} CUSTOM_AXES={axes} "
            f"CUSTOM_NAME={name}")


if __name__ == "__main__":
    from gorilla_physics_b200 import Mechanism
    model = sys.argv[1]
    m = Mechanism.from_model(model, [float(x) for x in sys.argv[2:]])
    print(f"make -C gorilla_physics_b200/csrc -j8 BUILD=../lib/obj_{model} OUT=../lib/alt/libgorilla_b200_{model}.so "
          + custom_topo_vars(m.desc(), model + "_custom"))
