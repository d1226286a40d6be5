"""Environment sharding across the GPUs of one box.

Environments are independent, so the step path needs no collective: rank g owns the contiguous
range [g*N/G, (g+1)*N/G) and keeps its state resident on its own device for the whole rollout.
The only exchange is an optional end-of-rollout reduction of a few diagnostic scalars
(sum KE, sum PE, sum spring energy, flagged environments) with torch.distributed
(NCCL over NVLink on the GPU box; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous environment range [lo, hi) of `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return n_total * rank // world, n_total * (rank + 1) // world


def shard_sizes(n_total: int, world: int):
    return [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]


def reduce_diagnostics(sums, group=None):
    """All-reduce (sum) the 4-vector written by MechanismState.energy_sums_device / a CPU tensor of
    the same layout. No-op without an initialised process group. Returns the tensor."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


class ShardedMechanismState:
    """The environments of one batch spread over several GPUs of the box FROM ONE PROCESS (the reference is
    one): the library's gp_sharded (csrc/gp_sharded.cpp) - one resident batch, stream and contiguous environment
    range (`shard_range`) per device, no exchange on the step path, one worker thread per device inside the
    library for the calls that move host data. `bench.py` and the multi-process tests use the process-per-GPU
    form instead. `shards` are views of the per-device batches for everything else MechanismState offers.
    """

    def __init__(self, mechanism, n_envs: int, devices):
        import ctypes as C

        from ._abi import check, lib
        from .mechanism import MechanismState
        self.devices = list(devices)
        if not self.devices:
            raise ValueError("no devices")
        self.n_envs = int(n_envs)
        self.mechanism = mechanism
        self.n_q, self.n_v = mechanism.n_q, mechanism.n_v
        if self.n_envs < len(self.devices):
            raise ValueError(f"{n_envs} environments do not cover {len(self.devices)} devices")
        h = C.c_void_p()
        ids = (C.c_int * len(self.devices))(*self.devices)
        check(lib().gp_sharded_create(mechanism._h, self.n_envs, ids, len(self.devices), C.byref(h)))
        self._h = h
        self.ranges, self.shards = [], []
        for g in range(lib().gp_sharded_n_shards(self._h)):
            lo, hi = C.c_int64(), C.c_int64()
            b = lib().gp_sharded_shard(self._h, g, C.byref(lo), C.byref(hi))
            self.ranges.append((lo.value, hi.value))
            self.shards.append(MechanismState._borrowed(mechanism, b, self))
        assert self.ranges == [shard_range(self.n_envs, r, len(self.devices)) for r in range(len(self.devices))]

    def __del__(self):
        try:
            from ._abi import lib
            if getattr(self, "_h", None):
                for s in self.shards:
                    s._h = None
                lib().gp_sharded_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _rows(self, a, k, what):
        import numpy as np
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        if a.size == k and self.n_envs > 1:
            a = np.ascontiguousarray(np.broadcast_to(a.reshape(1, k), (self.n_envs, k)))
        if a.size != self.n_envs * k:
            raise ValueError(f"{what} must hold {self.n_envs} x {k} values")
        return a.reshape(self.n_envs, k)

    @staticmethod
    def _p(a):
        import ctypes as C
        return None if a is None else C.c_void_p(a.ctypes.data)

    def update(self, q, v):
        from ._abi import check, lib
        qa, va = self._rows(q, self.n_q, "q"), self._rows(v, self.n_v, "v")
        check(lib().gp_sharded_set_state(self._h, self._p(qa), self._p(va)))

    def set_tau(self, tau):
        """per-environment torques [n_envs, n_v] (each shard receives its own rows), one row for all, or None = zeros"""
        from ._abi import check, lib
        ta = None if tau is None else self._rows(tau, self.n_v, "tau")
        check(lib().gp_sharded_set_tau(self._h, self._p(ta)))

    def step(self, dt, tau="keep", integrator=0, n_steps=1, controller=0, ctrl_params=()):
        import ctypes as C

        import numpy as np

        from ._abi import check, lib
        if not (isinstance(tau, str) and tau == "keep"):
            self.set_tau(tau)
        p = np.asarray(list(ctrl_params), dtype=np.float64) if len(ctrl_params) else None
        check(lib().gp_sharded_step(self._h, dt, int(integrator), int(n_steps), int(controller),
                                    None if p is None else p.ctypes.data_as(C.POINTER(C.c_double)), len(ctrl_params)))

    def synchronize(self):
        from ._abi import check, lib
        check(lib().gp_sharded_sync(self._h))

    def state(self):
        import numpy as np

        from ._abi import check, lib
        q, v = np.empty((self.n_envs, self.n_q)), np.empty((self.n_envs, self.n_v))
        check(lib().gp_sharded_get_state(self._h, self._p(q), self._p(v)))
        return q, v

    def simulate(self, final_time, dt, q, v, tau=None, integrator=0, controller=0, ctrl_params=()):
        """simulate() through host buffers on every device at once; q / v (float64, C-contiguous,
        [n_envs, .]) are updated in place. Returns the number of steps."""
        import ctypes as C

        import numpy as np

        from ._abi import check, lib
        for name, a in (("q", q), ("v", v)):
            if not (isinstance(a, np.ndarray) and a.flags.c_contiguous and a.dtype == np.float64):
                raise ValueError(f"{name} must be a C-contiguous float64 array (it is updated in place)")
        if q.size != self.n_envs * self.n_q or v.size != self.n_envs * self.n_v:
            raise ValueError("q / v do not match the batch")
        ta = None if tau is None else self._rows(tau, self.n_v, "tau")
        p = np.asarray(list(ctrl_params), dtype=np.float64) if len(ctrl_params) else None
        n = C.c_int64()
        check(lib().gp_sharded_simulate(self._h, self._p(q), self._p(v), self._p(ta), final_time, dt, int(integrator),
                                        int(controller), None if p is None else p.ctypes.data_as(C.POINTER(C.c_double)),
                                        len(ctrl_params), C.byref(n)))
        return n.value

    def status(self):
        import numpy as np

        from ._abi import check, lib
        out = np.zeros(self.n_envs, dtype=np.uint32)
        check(lib().gp_sharded_status(self._h, self._p(out)))
        return out

    def energy_sums(self):
        """(sum KE, sum PE, sum spring energy) over every device: the end-of-rollout diagnostic; in one
        process the "reduction" is a host-side sum of one 4-vector per device (gp_sharded_energy_sums)."""
        import ctypes as C

        import numpy as np

        from ._abi import check, lib
        out = np.zeros(4)
        check(lib().gp_sharded_energy_sums(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return float(out[0]), float(out[1]), float(out[2])
