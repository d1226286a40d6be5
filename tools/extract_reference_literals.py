#!/usr/bin/env python
"""Extract the physics literals of the reference's SO-101 and navbot builders into
tests/golden/model_literals.json (run in the build container, where /root/reference exists).

Parses src/builders/mod.rs (build_so101*) and src/builders/navbot_builder.rs (build_navbot*):
per body m, com, COM-frame inertia entries; per joint the parent frame, xyz/rpy origin, joint type
and axis. The JSON is the golden fixture tests/test_host_cpu.py::test_model_literals_match_reference_sources
checks the product's builders against. Nothing at run time reads /root/reference.
"""
import json
import re
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
NUM = r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+)"


def num(s):
    return float(s.rstrip("."))


def body_literals(src, fn_name):
    m = re.search(r"fn %s\b.*?\n}\n" % re.escape(fn_name), src, re.S)
    assert m, fn_name
    body = m.group(0)
    # ignore commented-out alternatives
    body = "\n".join(l for l in body.splitlines() if not l.strip().startswith("//"))
    out = {"m": num(re.search(r"let m = (%s);" % NUM, body).group(1))}
    com = re.search(r"let com = vector!\[(%s), (%s), (%s)\];" % (NUM, NUM, NUM), body)
    out["com"] = [num(com.group(k)) for k in (1, 2, 3)]
    for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz"):
        out[k] = num(re.search(r"let %s = (%s);" % (k, NUM), body).group(1))
    return out


def joint_origins(build_src):
    """frame -> (parent_frame, xyz, rpy) from Transform3D::new_xyz_rpy calls; identity transforms too."""
    names = dict(re.findall(r'let (\w+)_frame = "(\w+)";', build_src))  # var stem -> frame name
    origins = {}
    for m in re.finditer(r"Transform3D::new_xyz_rpy\(\s*(\w+)_frame,\s*(\w+)_frame,\s*&vec!\[(.*?)\],\s*&vec!\[(.*?)\],?\s*\)",
                         build_src, re.S):
        child, parent = names[m.group(1)], names[m.group(2)]
        xyz = [num(x) for x in re.findall(NUM, m.group(3))]
        rpy = [num(x) for x in re.findall(NUM, m.group(4))]
        origins[child] = (parent, xyz, rpy)
    for m in re.finditer(r"Transform3D::identity\((\w+)_frame, WORLD_FRAME\)", build_src):
        origins[names[m.group(1)]] = ("world", None, None)
    return names, origins


def joint_list(build_src, var_to_frame_from_transform):
    m = re.search(r"let treejoints = vec!\[(.*?)\];", build_src, re.S)
    text = "\n".join(l for l in m.group(1).splitlines() if not l.strip().startswith("//"))
    joints = []
    for jm in re.finditer(r"Joint::(\w+)Joint\(\w+Joint::new\(\s*(\w+?)(?:,\s*(-?)Vector3::(\w)_axis\(\))?,?\s*\)\)", text, re.S):
        kind, tvar, neg, ax = jm.groups()
        joints.append((kind, tvar, neg, ax))
    return joints


def model(src, build_fn, body_fn_of):
    bm = re.search(r"pub fn %s\b.*?\n}\n" % build_fn, src, re.S).group(0)
    names, origins = joint_origins(bm)
    # transform variable -> child frame:  let shoulder_to_base = Transform3D::...(shoulder_frame, ...
    tvars = {}
    for m in re.finditer(r"let (\w+) = Transform3D::(?:new_xyz_rpy|identity)\(\s*&?(\w+)_frame", bm):
        tvars[m.group(1)] = names[m.group(2)]
    joints = joint_list(bm, tvars)
    order = [tvars[t] for _, t, _, _ in joints]
    kind_id = {"Fixed": 0, "Revolute": 1, "Prismatic": 2, "Floating": 3}
    bodies = []
    for (kind, tvar, neg, ax), frame in zip(joints, order):
        parent, xyz, rpy = origins[frame]
        lit = body_literals(src, body_fn_of(frame))
        axis = None
        if ax:
            axis = [0.0, 0.0, 0.0]
            axis["xyz".index(ax)] = -1.0 if neg else 1.0
        lit.update({"name": frame, "parent": 0 if parent == "world" else order.index(parent) + 1,
                    "joint_type": kind_id[kind], "xyz": xyz, "rpy": rpy, "axis": axis})
        bodies.append(lit)
    return {"bodies": bodies}


def main():
    so = (REF / "src/builders/mod.rs").read_text()
    nav = (REF / "src/builders/navbot_builder.rs").read_text()
    out = {
        "_source": "one-for-all/gorilla-physics src/builders/mod.rs, src/builders/navbot_builder.rs",
        "so101": model(so, "build_so101", lambda f: f"build_so101_{f}_body"),
        "navbot": model(nav, "build_navbot", lambda f: f"build_navbot_{f}"),
    }
    dst = Path(__file__).resolve().parent.parent / "tests" / "golden" / "model_literals.json"
    dst.write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", dst, {k: len(v["bodies"]) for k, v in out.items() if k != "_source"})


if __name__ == "__main__":
    main()
