"""gorilla_physics_b200 — B200-native batched articulated-dynamics stepper.

Drop-in for the Mechanism / MechanismState / step() / simulate() path of
one-for-all/gorilla-physics, executed by hand-written sm_100a CUDA kernels behind a C ABI
(include/gorilla_b200.h, gorilla_physics_b200/lib/libgorilla_b200.so). No CPU fallback.
"""
from .desc import (FIXED, FLOATING, PRISMATIC, REVOLUTE, MechanismDesc, iso, iso_xyz_rpy, quat_from_axis_angle,
                   quat_from_euler, quat_from_scaled_axis)
from .mechanism import (Communicator, Controller, Integrator, KernelMode, Mechanism, MechanismState, jit_available, jit_cache_dir,
                        measure_fp64_peak, measure_fp64_peak_trace, nccl_available, simulate, step)
from .sharding import ShardedMechanismState, shard_range
from .workloads import WORKLOADS, Workload

GRAVITY = 9.81  # reference src/lib.rs:39

__all__ = [
    "FIXED", "REVOLUTE", "PRISMATIC", "FLOATING", "MechanismDesc", "iso", "iso_xyz_rpy", "quat_from_euler",
    "quat_from_axis_angle", "quat_from_scaled_axis", "Mechanism", "MechanismState", "Integrator", "Controller",
    "step", "simulate", "measure_fp64_peak", "GRAVITY", "ShardedMechanismState", "shard_range", "KernelMode",
    "jit_available", "jit_cache_dir", "measure_fp64_peak_trace", "WORKLOADS", "Workload", "Communicator", "nccl_available",
]
