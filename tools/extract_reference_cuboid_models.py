#!/usr/bin/env python
"""Evaluate the reference's cuboid-built models - build_biped (src/builders/biped_builder.rs), build_leg and
build_leg_from_foot (src/builders/leg_builder.rs) - into tests/golden/cuboid_models.json (run in the build
container, where /root/reference exists; nothing at run time reads /root/reference).

Unlike SO-101 / navbot (tools/extract_reference_literals.py) these builders are arithmetic over a few lengths and
masses, so this script EVALUATES the Rust statements instead of collecting literals: `let x = <arithmetic>;`,
`vector![..]`, `(-)Vector3::?_axis()`, `RigidBody::new_cuboid[_at]`, `add_cuboid_contacts[_with]`,
`Transform3D::identity / move_x / move_z / move_xyz`, and the `treejoints` / `bodies` vectors. The JSON keeps, per
body in joint order: frame, parent (0 = world), joint type, axis, joint origin xyz, the cuboid's m, w, d, h and
centre, and the contact points in the order the reference adds them.
tests/test_host_cpu.py::test_cuboid_models_match_reference_sources holds the product's builders against it.
"""
import json
import re
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
KIND = {"fixed": 0, "revolute": 1, "prismatic": 2, "floating": 3}


def split_args(s):
    """top-level comma split (brackets and parentheses nest)"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


class Env:
    def __init__(self):
        self.num, self.vec, self.axis, self.frame, self.body, self.xform = {}, {}, {}, {}, {}, {}

    def f(self, expr):
        return float(eval(expr, {"__builtins__": {}}, dict(self.num)))  # arithmetic over earlier `let`s only

    def v(self, expr):
        expr = expr.lstrip("&").strip()
        m = re.fullmatch(r"vector!\[(.*)\]", expr, re.S)
        if m:
            return [self.f(a) for a in split_args(m.group(1))]
        return list(self.vec[expr])

    def fr(self, expr):
        return "world" if expr == "WORLD_FRAME" else self.frame[expr]


def evaluate(src, fn_name):
    body = re.search(r"pub fn %s\b.*?\n}\n" % fn_name, src, re.S).group(0)
    body = re.sub(r"//[^\n]*", "", body)
    body = body[body.index("{") + 1:]
    env = Env()
    joints = bodies = None
    for stmt in (s.strip() for s in body.split(";")):
        if not stmt:
            continue
        m = re.fullmatch(r"(\w+)\.add_cuboid_contacts(_with)?\((.*)\)", stmt, re.S)
        if m:
            b, args = env.body[m.group(1)], split_args(m.group(3))
            if m.group(2):
                c, (w, d, h) = env.v(args[0]), (env.f(a) for a in args[1:4])
                for i in (-1.0, 1.0):      # rigid_body.rs:251-264
                    for j in (-1.0, 1.0):
                        for k in (-1.0, 1.0):
                            b["contacts"].append([c[0] + i * w / 2.0, c[1] + j * d / 2.0, c[2] + k * h / 2.0])
            else:
                w, d, h = (env.f(a) for a in args[0:3])
                for sz in (1.0, -1.0):     # rigid_body.rs:216-249
                    for sy in (1.0, -1.0):
                        for sx in (-1.0, 1.0):
                            b["contacts"].append([sx * w / 2.0, sy * d / 2.0, sz * h / 2.0])
            continue
        m = re.fullmatch(r"let (?:mut )?(\w+)(?:\s*:\s*[^=]+?)?\s*=\s*(.*)", stmt, re.S)
        if not m:
            continue  # the closing `MechanismState::new(treejoints, bodies) }`
        name, rhs = m.group(1), m.group(2).strip()
        if name == "treejoints":
            joints = []
            for j in split_args(re.fullmatch(r"vec!\[(.*)\]", rhs, re.S).group(1)):
                jm = re.fullmatch(r"Joint::new_(\w+)\((.*)\)", j, re.S) or \
                    re.fullmatch(r"Joint::(\w+)Joint\(\w+::new\((.*)\)\)", j, re.S)
                a = split_args(jm.group(2))
                joints.append((KIND[jm.group(1).lower()], a[0], env.axis[a[1]] if len(a) > 1 else None))
        elif name == "bodies":
            bodies = split_args(re.fullmatch(r"vec!\[(.*)\]", rhs, re.S).group(1))
        elif rhs.startswith('"'):
            env.frame[name] = rhs.strip('"')
        elif rhs.startswith("vector!"):
            env.vec[name] = env.v(rhs)
        elif re.fullmatch(r"-?\s*Vector3::[xyz]_axis\(\)", rhs):
            ax = [0.0, 0.0, 0.0]
            ax["xyz".index(rhs[rhs.index("::") + 2])] = -1.0 if rhs.startswith("-") else 1.0
            env.axis[name] = ax
        elif rhs.startswith("RigidBody::new_cuboid"):
            args = split_args(rhs[rhs.index("(") + 1:rhs.rindex(")")])
            com = [0.0, 0.0, 0.0]
            if rhs.startswith("RigidBody::new_cuboid_at"):
                com, args = env.v(args[0]), args[1:]
            env.body[name] = {"frame": env.fr(args[4]), "m": env.f(args[0]), "w": env.f(args[1]), "d": env.f(args[2]),
                              "h": env.f(args[3]), "com": com, "contacts": []}
        elif rhs.startswith("Transform3D::"):
            kind = rhs[len("Transform3D::"):rhs.index("(")]
            args = split_args(rhs[rhs.index("(") + 1:rhs.rindex(")")])
            xyz = {"identity": lambda: [0.0, 0.0, 0.0], "move_x": lambda: [env.f(args[2]), 0.0, 0.0],
                   "move_z": lambda: [0.0, 0.0, env.f(args[2])],
                   "move_xyz": lambda: [env.f(args[2]), env.f(args[3]), env.f(args[4])]}[kind]()
            env.xform[name] = (env.fr(args[0]), env.fr(args[1]), xyz)
        else:
            env.num[name] = env.f(rhs)
    assert joints and bodies and len(joints) == len(bodies), fn_name
    order = [env.body[b]["frame"] for b in bodies]
    out = []
    for (jt, tvar, axis), bname in zip(joints, bodies):
        child, parent, xyz = env.xform[tvar]
        b = dict(env.body[bname])
        assert child == b["frame"], (fn_name, tvar, child, b["frame"])  # mechanism.rs:91-96: joint i <-> body i
        b.update({"parent": 0 if parent == "world" else order.index(parent) + 1, "joint_type": jt, "axis": axis, "xyz": xyz})
        out.append(b)
    return {"bodies": out}


def main():
    biped = (REF / "src/builders/biped_builder.rs").read_text()
    leg = (REF / "src/builders/leg_builder.rs").read_text()
    out = {
        "_source": "one-for-all/gorilla-physics src/builders/biped_builder.rs, src/builders/leg_builder.rs "
                   "(evaluated by tools/extract_reference_cuboid_models.py)",
        "biped": evaluate(biped, "build_biped"),
        "leg": evaluate(leg, "build_leg"),
        "leg_from_foot": evaluate(leg, "build_leg_from_foot"),
    }
    dst = Path(__file__).resolve().parent.parent / "tests" / "golden" / "cuboid_models.json"
    dst.write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", dst, {k: (len(v["bodies"]), sum(len(b["contacts"]) for b in v["bodies"])) for k, v in out.items() if k != "_source"})


if __name__ == "__main__":
    main()
