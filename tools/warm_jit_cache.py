#!/usr/bin/env python3
"""Compiles, ahead of time and without a GPU, the run-time-specialised kernels that the GPU test-suite,
smoke() and the bench's JIT workloads launch, into gorilla_physics_b200/lib/jit_cache (which travels to the
GPU box with the library). Without it the same compilations happen on first launch (NVRTC, 5-70 s each).

    python tools/warm_jit_cache.py [-j N]
"""
import argparse
import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

SIE, RK, DYN, ENERGY = 1, 2, 4, 8


def mechanisms():
    """name -> (factory of the mechanism on its JIT kernel, kinds)"""
    import numpy as np

    from gorilla_physics_b200 import KernelMode, Mechanism
    from tests import models
    from tests.test_parity_gpu import WORKLOADS, generic_twin

    out = {}
    for name, (factory, _, _) in WORKLOADS.items():
        out["twin:" + name] = (lambda f=factory: generic_twin(f(), KernelMode.JIT), SIE | RK | DYN)
    out["twin:cube_in_corner"] = (lambda: generic_twin(models.cube_in_corner(), KernelMode.JIT), SIE | DYN)

    def slip():
        m = Mechanism.from_model("slip")
        m.add_halfspace((0, 0, 1), -0.3)
        return generic_twin(m, KernelMode.JIT)
    out["twin:slip"] = (slip, SIE | DYN)
    for seed in range(6):
        nb = 1 + (3 * seed + 2) % 8
        out[f"random_tree:{seed}"] = (lambda s=seed, n=nb: Mechanism.from_desc(models.random_tree(1000 + s, n)), SIE | RK | DYN)
    out["maximum_size"] = (lambda: Mechanism.from_desc(models.maximum_size_mechanism(), kernel=KernelMode.JIT), SIE | RK | DYN)
    # bench.py --workload jit_* (the twins again, plus their energy kernels for the diagnostics)
    for name in ("so101_contact", "navbot_contact"):
        out["bench:" + name] = (lambda f=WORKLOADS[name][0]: generic_twin(f(), KernelMode.JIT), SIE | DYN | ENERGY)
    # the reference's cuboid-built trees on the ground (tests/test_zz_widening_gpu.py, bench.py --workload biped)
    def grounded(name):
        m = Mechanism.from_model(name)
        m.add_halfspace((0, 0, 1), 0.0)
        return m
    out["model:biped"] = (lambda: grounded("biped"), SIE | RK | DYN | ENERGY)
    out["model:leg"] = (lambda: grounded("leg"), SIE | RK | DYN)
    out["model:leg_from_foot"] = (lambda: grounded("leg_from_foot"), SIE | RK | DYN)
    return out


def work(task):
    name, kind = task
    factory, _ = mechanisms()[name]
    t = time.time()
    m = factory()
    n = m.precompile(kind)
    return name, kind, m.kernel_variant, n, time.time() - t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=max(1, min(16, os.cpu_count() or 1)))
    a = ap.parse_args()
    from gorilla_physics_b200 import jit_available, jit_cache_dir
    if not jit_available():
        print("NVRTC not loadable: nothing to do")
        return
    tasks = [(name, k) for name, (_, kinds) in mechanisms().items() for k in (SIE, RK, DYN, ENERGY) if kinds & k]
    t0 = time.time()
    compiled = 0
    with mp.get_context("spawn").Pool(a.j) as pool:
        for name, kind, variant, n, dt in pool.imap_unordered(work, tasks):
            compiled += n
            if n:
                print(f"  {name:28s} {variant:24s} kind {kind}: {dt:5.1f} s", flush=True)
    print(f"{len(tasks)} kernels, {compiled} compiled in {time.time() - t0:.0f} s -> {jit_cache_dir()}")


if __name__ == "__main__":
    main()
