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
    """The environments of one batch spread over several GPUs of the box FROM ONE PROCESS: one
    `MechanismState` (its own device, stream and resident planes) per device, contiguous environment
    ranges (`shard_range`), no exchange on the step path. This is how a single-process host (the
    reference is one) drives the C ABI on a multi-GPU box; `bench.py` and the tests use the
    process-per-GPU form with `torch.distributed` instead. `step` only enqueues (the devices run
    concurrently); calls that move host data run one worker thread per device (ctypes releases the GIL).
    """

    def __init__(self, mechanism, n_envs: int, devices):
        from concurrent.futures import ThreadPoolExecutor

        from .mechanism import MechanismState
        self.devices = list(devices)
        if not self.devices:
            raise ValueError("no devices")
        self.n_envs = int(n_envs)
        self.ranges = [shard_range(self.n_envs, r, len(self.devices)) for r in range(len(self.devices))]
        if any(hi <= lo for lo, hi in self.ranges):
            raise ValueError(f"{n_envs} environments do not cover {len(self.devices)} devices")
        self.shards = [MechanismState(mechanism, hi - lo, device=d) for (lo, hi), d in zip(self.ranges, self.devices)]
        self.n_q, self.n_v = self.shards[0].n_q, self.shards[0].n_v
        self._pool = ThreadPoolExecutor(max_workers=len(self.devices))

    def _each(self, fn):
        return list(self._pool.map(lambda a: fn(*a), [(s, lo, hi) for s, (lo, hi) in zip(self.shards, self.ranges)]))

    def update(self, q, v):
        import numpy as np
        q = np.ascontiguousarray(np.asarray(q, dtype=np.float64).reshape(self.n_envs, self.n_q))
        v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(self.n_envs, self.n_v))
        self._each(lambda s, lo, hi: s.update(q[lo:hi], v[lo:hi]))

    def _slice_kw(self, kw, lo, hi):
        """per-environment keyword arguments (tau of shape [n_envs, n_v]) go to each shard as its own rows"""
        import numpy as np
        out = dict(kw)
        tau = out.get("tau")
        if tau is not None and not isinstance(tau, (int, np.integer)):
            t = np.asarray(tau, dtype=np.float64)
            if t.ndim == 2 and t.shape[0] == self.n_envs:
                out["tau"] = np.ascontiguousarray(t[lo:hi])
            elif t.size == self.n_envs * self.n_v and t.ndim == 1 and self.n_envs > 1:
                out["tau"] = np.ascontiguousarray(t.reshape(self.n_envs, self.n_v)[lo:hi])
        return out

    def step(self, dt, **kw):
        for s, (lo, hi) in zip(self.shards, self.ranges):  # asynchronous per device
            s.step(dt, **self._slice_kw(kw, lo, hi))

    def synchronize(self):
        for s in self.shards:
            s.synchronize()

    def state(self):
        import numpy as np
        parts = self._each(lambda s, lo, hi: s.state())
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])

    def simulate(self, final_time, dt, q, v, **kw):
        """simulate() through host buffers on every device at once; q / v (float64, C-contiguous,
        [n_envs, .]) are updated in place. Returns the number of steps."""
        import numpy as np
        if not (isinstance(q, np.ndarray) and q.flags.c_contiguous and q.dtype == np.float64):
            raise ValueError("q must be a C-contiguous float64 array (it is updated in place)")
        if not (isinstance(v, np.ndarray) and v.flags.c_contiguous and v.dtype == np.float64):
            raise ValueError("v must be a C-contiguous float64 array (it is updated in place)")
        q2, v2 = q.reshape(self.n_envs, self.n_q), v.reshape(self.n_envs, self.n_v)
        done = self._each(lambda s, lo, hi: s.simulate(final_time, dt, q2[lo:hi], v2[lo:hi], **self._slice_kw(kw, lo, hi))[0])
        return done[0]

    def status(self):
        import numpy as np
        return np.concatenate(self._each(lambda s, lo, hi: s.status()))

    def energy_sums(self):
        """(sum KE, sum PE, sum spring energy) over every device: the end-of-rollout diagnostic; in one
        process the "reduction" is a host-side sum of one triple per device."""
        import numpy as np
        parts = self._each(lambda s, lo, hi: tuple(float(np.sum(e)) for e in s.energies()))
        return tuple(sum(p[k] for p in parts) for k in range(3))
