"""ctypes binding of the C ABI (include/gorilla_b200.h) — loads gorilla_physics_b200/lib/libgorilla_b200.so.

Loading needs no GPU (the library links the CUDA runtime statically); every compute entry
point fails with GP_ERR_NO_DEVICE when no CUDA device is usable. There is no CPU fallback:
if the shared library is missing this module raises instead of substituting anything.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
# GP_LIB_PATH lets a developer point at an experimental build of the same library
LIB_PATH = Path(os.environ.get("GP_LIB_PATH", _HERE / "lib" / "libgorilla_b200.so"))

GP_OK, GP_ERR_INVALID, GP_ERR_UNSUPPORTED, GP_ERR_NO_DEVICE, GP_ERR_CUDA, GP_ERR_LIMIT, GP_ERR_JIT = range(7)

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
vp = C.c_void_p


class GpMechanismDesc(C.Structure):
    _fields_ = [
        ("n_bodies", C.c_int32),
        ("parent", ip),
        ("joint_type", ip),
        ("axis", dp),
        ("init_iso", dp),
        ("moment", dp),
        ("cross_part", dp),
        ("mass", dp),
        ("has_spring", ip),
        ("spring_k", dp),
        ("spring_l", dp),
        ("n_contact_points", C.c_int32),
        ("cp_body", ip),
        ("cp_location", dp),
        ("cp_k", dp),
        ("n_halfspaces", C.c_int32),
        ("hs_point", dp),
        ("hs_normal", dp),
        ("hs_alpha", dp),
        ("hs_mu", dp),
        ("armature", dp),
        ("n_spring_contacts", C.c_int32),
        ("sc_body", ip),
        ("sc_l_rest", dp),
        ("sc_direction", dp),
        ("sc_k", dp),
    ]


class GpStateDist(C.Structure):
    _fields_ = [
        ("q_lo", C.c_double), ("q_hi", C.c_double), ("v_lo", C.c_double), ("v_hi", C.c_double),
        ("base_t", C.c_double * 3), ("t_jitter", C.c_double * 3), ("rpy_jitter", C.c_double),
        ("base_v", C.c_double * 6), ("v_jitter", C.c_double),
    ]


# every symbol include/gorilla_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "gp_abi_version": (C.c_int, []),
    "gp_last_error": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "gp_device_count": (C.c_int, []),
    "gp_mechanism_create": (C.c_int, [C.POINTER(GpMechanismDesc), C.POINTER(vp)]),
    "gp_mechanism_destroy": (None, [vp]),
    "gp_mechanism_n_bodies": (C.c_int, [vp]),
    "gp_mechanism_n_q": (C.c_int, [vp]),
    "gp_mechanism_n_v": (C.c_int, [vp]),
    "gp_mechanism_n_contact_points": (C.c_int, [vp]),
    "gp_mechanism_n_halfspaces": (C.c_int, [vp]),
    "gp_mechanism_get_desc": (C.c_int, [vp, C.POINTER(GpMechanismDesc)]),
    "gp_mechanism_add_halfspace": (C.c_int, [vp, dp, dp, C.c_double, C.c_double]),
    "gp_mechanism_add_contact_point": (C.c_int, [vp, C.c_int32, dp, C.c_double]),
    "gp_mechanism_add_spring_contact": (C.c_int, [vp, C.c_int32, C.c_double, dp, C.c_double]),
    "gp_mechanism_n_spring_contacts": (C.c_int, [vp]),
    "gp_mechanism_supports": (C.c_int, [vp, ip]),
    "gp_mechanism_kernel_variant": (C.c_char_p, [vp]),
    "gp_mechanism_set_kernel_mode": (C.c_int, [vp, C.c_int]),
    "gp_jit_available": (C.c_int, []),
    "gp_jit_cache_dir": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "gp_mechanism_precompile": (C.c_int, [vp, C.c_uint, C.POINTER(C.c_int)]),
    "gp_model_create": (C.c_int, [C.c_char_p, dp, C.c_int, C.POINTER(vp)]),
    "gp_batch_create": (C.c_int, [vp, C.c_int64, C.c_int, C.POINTER(vp)]),
    "gp_batch_destroy": (None, [vp]),
    "gp_batch_n_envs": (C.c_int64, [vp]),
    "gp_batch_ld": (C.c_int64, [vp]),
    "gp_batch_device": (C.c_int, [vp]),
    "gp_batch_q_device": (vp, [vp]),
    "gp_batch_v_device": (vp, [vp]),
    "gp_batch_tau_device": (vp, [vp]),
    "gp_batch_stream": (vp, [vp]),
    "gp_batch_sync": (C.c_int, [vp]),
    "gp_batch_step_lanes": (C.c_int, [vp]),
    "gp_batch_launch_count": (C.c_int64, [vp]),
    "gp_batch_set_state": (C.c_int, [vp, vp, vp]),
    "gp_batch_get_state": (C.c_int, [vp, vp, vp]),
    "gp_batch_set_tau": (C.c_int, [vp, vp]),
    "gp_batch_set_spring_contact_state": (C.c_int, [vp, vp]),
    "gp_batch_get_spring_contact_state": (C.c_int, [vp, vp]),
    "gp_batch_set_controller_state": (C.c_int, [vp, vp]),
    "gp_batch_get_controller_state": (C.c_int, [vp, vp]),
    "gp_batch_set_controller_state_n": (C.c_int, [vp, vp, C.c_int]),
    "gp_batch_get_controller_state_n": (C.c_int, [vp, vp, C.c_int]),
    "gp_batch_randomize": (C.c_int, [vp, C.c_uint64, C.POINTER(GpStateDist)]),
    "gp_batch_dynamics": (C.c_int, [vp, vp, vp]),
    "gp_batch_free_velocity": (C.c_int, [vp, C.c_double, C.c_int, vp]),
    "gp_batch_mass_matrix": (C.c_int, [vp, vp, vp]),
    "gp_batch_step": (C.c_int, [vp, C.c_double, C.c_int, C.c_int, C.c_int, dp, C.c_int]),
    "gp_batch_step_tau_sequence": (C.c_int, [vp, C.c_double, C.c_int, C.c_int, vp]),
    "gp_batch_step_tau_sequence_device": (C.c_int, [vp, C.c_double, C.c_int, C.c_int, vp]),
    "gp_batch_simulate": (C.c_int, [vp, vp, vp, vp, C.c_double, C.c_double, C.c_int, C.c_int, dp, C.c_int,
                                    C.POINTER(C.c_int64), vp, vp]),
    "gp_simulate_step_count": (C.c_int64, [C.c_double, C.c_double]),
    "gp_batch_energy": (C.c_int, [vp, vp, vp, vp]),
    "gp_batch_energy_sums_device": (C.c_int, [vp, vp]),
    "gp_batch_poses": (C.c_int, [vp, vp]),
    "gp_batch_status": (C.c_int, [vp, vp]),
    "gp_batch_clear_status": (C.c_int, [vp]),
    "gp_sharded_create": (C.c_int, [vp, C.c_int64, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
    "gp_sharded_destroy": (None, [vp]),
    "gp_sharded_n_shards": (C.c_int, [vp]),
    "gp_sharded_n_envs": (C.c_int64, [vp]),
    "gp_sharded_shard": (vp, [vp, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "gp_sharded_set_state": (C.c_int, [vp, vp, vp]),
    "gp_sharded_get_state": (C.c_int, [vp, vp, vp]),
    "gp_sharded_set_tau": (C.c_int, [vp, vp]),
    "gp_sharded_step": (C.c_int, [vp, C.c_double, C.c_int, C.c_int, C.c_int, dp, C.c_int]),
    "gp_sharded_sync": (C.c_int, [vp]),
    "gp_sharded_simulate": (C.c_int, [vp, vp, vp, vp, C.c_double, C.c_double, C.c_int, C.c_int, dp, C.c_int,
                                      C.POINTER(C.c_int64)]),
    "gp_sharded_status": (C.c_int, [vp, vp]),
    "gp_sharded_energy_sums": (C.c_int, [vp, dp]),
    "gp_nccl_available": (C.c_int, []),
    "gp_comm_unique_id": (C.c_int, [C.c_char_p]),
    "gp_comm_create": (C.c_int, [C.c_int, C.c_int, C.c_char_p, C.c_int, C.POINTER(vp)]),
    "gp_comm_destroy": (None, [vp]),
    "gp_batch_reduce_diagnostics": (C.c_int, [vp, vp, dp]),
    "gp_measure_fp64_peak": (C.c_int, [C.c_int, C.c_double, dp]),
    "gp_measure_fp64_peak_trace": (C.c_int, [C.c_int, C.c_double, dp, dp, C.c_int, C.POINTER(C.c_int)]),
}

_lib = None


class GorillaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[gp status {code}] {message}")
        self.code = code


def lib() -> C.CDLL:
    """Load the shared library (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). gorilla_physics_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(2048)
    lib().gp_last_error(buf, len(buf))
    return buf.value.decode("utf-8", "replace")


def check(rc: int):
    if rc != GP_OK:
        raise GorillaError(rc, last_error())
