#!/usr/bin/env python
"""Headline benchmark: env-steps/s of the batched articulated-dynamics stepper (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm on the host cores

A bench "step" is one launch of the step kernel: `--inner` fused time steps (semi-implicit Euler)
of every environment of the workload. Workload (default): SO-101 6-DOF arm with ground contact,
262144 environments per GPU (the configuration BASELINE.json's target is
quoted on). Environments are independent, so ranks shard them with no collective on the step path
(weak scaling: per-GPU work fixed); one optional all-reduce of energy sums happens after the timed
region. Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"

# FP64 work per environment time step of the kernel that runs each workload: 2*DFMA + DADD + DMUL
# thread-level instructions per env-step as counted by ncu
# (smsp__sass_thread_inst_executed_op_d{fma,add,mul}_pred_on.sum summed over the 40 timed launches of
# this file's default command, divided by n_envs * inner steps * 40: contact workloads execute more
# once their bodies lie on the ground, tools/ncu_flops_over_bench.py). profiles/flop_counts.json holds
# the numbers with the capture they come from; DESIGN.md §4 explains why this (executed, not
# reference-formulation) count is used.
def load_flop_counts():
    try:
        return json.loads((ROOT / "profiles" / "flop_counts.json").read_text())["flop_per_env_step"]
    except Exception:
        return {}


WORKLOADS = {
    # name: (n_envs per GPU, dt, randomize kwargs)
    "so101_contact": (262144, 1.0 / 6000.0, dict(q_range=(-1.0, 1.0), v_range=(-1.0, 1.0))),
    "so101": (262144, 1.0 / 6000.0, dict(q_range=(-1.0, 1.0), v_range=(-1.0, 1.0))),
    "double_pendulum": (1048576, 1e-3, dict(q_range=(-math.pi, math.pi), v_range=(-1.0, 1.0))),
    "cart_pole": (1048576, 1e-3, dict(q_range=(-math.pi, math.pi), v_range=(-1.0, 1.0))),
    "rimless_wheel": (262144, 1.0 / 600.0, dict(base_t=(0.0, 0.0, -10.5), t_jitter=(0.0, 0.0, 0.5), rpy_jitter=0.3,
                                                base_v=(0, 0, 0, 1.0, 0, 0), v_jitter=0.2)),
    "hopper_1d": (262144, 1.0 / 500.0, dict(q_range=(0.0, 0.0), v_range=(0.0, 0.0), base_t=(0.0, 0.0, 2.5),
                                            t_jitter=(0.0, 0.0, 2.5))),
    "quadruped": (65536, 1.0 / 3000.0, dict(q_range=(-0.2, 0.2), v_range=(0.0, 0.0), base_t=(0.0, 0.0, 0.8),
                                            t_jitter=(0.01, 0.01, 0.01), rpy_jitter=0.1)),
    "navbot_contact": (65536, 1.0 / 6000.0, dict(q_range=(-0.2, 0.2), v_range=(0.0, 0.0), base_t=(0.0, 0.0, 0.075),
                                                 t_jitter=(0.01, 0.01, 0.01), rpy_jitter=0.1)),
}


def load_reference_flops(workload):
    """flop/env-step of the REFERENCE's formulation (oracle with a counting scalar type,
    tools/count_reference_flops.py -> profiles/roofline.json); information only: `achieved` counts what
    the kernel executes, which is 2.4-5.6x less"""
    try:
        return json.loads((ROOT / "profiles" / "roofline.json").read_text())["workloads"][workload][
            "reference_formulation_flop_per_env_step"]
    except Exception:
        return None


def load_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of one step-kernel launch (ncu --set full), or None"""
    try:
        return json.loads((ROOT / "profiles" / "flop_counts.json").read_text())["dram_bytes_per_launch"].get(workload)
    except Exception:
        return None


def make_mechanism(name):
    from gorilla_physics_b200 import Mechanism
    from tests import models
    if name == "so101_contact":
        return models.so101_with_contact()
    if name == "rimless_wheel":
        return models.rimless_wheel_on_slope()
    if name == "hopper_1d":
        return models.hopper1d_on_ground()
    if name == "quadruped":
        return models.quadruped_on_ground()
    if name == "navbot_contact":
        return models.navbot_with_contact()
    return Mechanism.from_model(name)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self._stop_evt = threading.Event()

    def run(self):
        # NVML in-process (sub-millisecond per sample); nvidia-smi subprocesses as the fallback
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = [(nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else nv.nvmlClocksThrottleReasonHwSlowdown),
                    (nv.nvmlClocksEventReasonHwThermalSlowdown if hasattr(nv, "nvmlClocksEventReasonHwThermalSlowdown") else nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                    (nv.nvmlClocksEventReasonSwThermalSlowdown if hasattr(nv, "nvmlClocksEventReasonSwThermalSlowdown") else nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                    (nv.nvmlClocksEventReasonSwPowerCap if hasattr(nv, "nvmlClocksEventReasonSwPowerCap") else nv.nvmlClocksThrottleReasonSwPowerCap)]
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._stop_evt.is_set():
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = get_reasons(h)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = 0.0
                self.samples.append([str(sm), str(mx), str(pw)] + ["Active" if (r & b) else "Not Active" for b in bits])
                self._stop_evt.wait(0.005)
            return
        except Exception:
            pass
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.gpu < len(ids) and ids[self.gpu].isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, flag in zip(names, s[3:7]) if flag.lower().startswith("active")})
        pw = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "reasons": reasons, "samples": len(self.samples)}


def cpu_reference_rate(workload, seconds_target=12.0, threads=None):
    """The reference's CPU algorithm (oracle port, reference operation order) on a bounded sample."""
    from oracle.binding import OracleMechanism
    from tests.test_parity_gpu import random_states  # same seeded state generator as the parity tests
    n_envs, dt, _ = WORKLOADS[workload]
    mech = make_mechanism(workload)
    desc = mech.desc()
    orc = OracleMechanism(desc)
    threads = threads or os.cpu_count() or 1
    n = 64 * threads
    kw = {}
    if workload in ("rimless_wheel",):
        kw = dict(base_t=(0, 0, -10.5), t_jitter=0.5, rpy_jitter=0.3)
    elif workload in ("quadruped",):
        kw = dict(base_t=(0, 0, 0.8), t_jitter=0.01, rpy_jitter=0.1, q_range=0.2)
    elif workload in ("navbot_contact",):
        kw = dict(base_t=(0, 0, 0.075), t_jitter=0.01, rpy_jitter=0.1, q_range=0.2)
    elif workload in ("hopper_1d",):
        kw = dict(base_t=(0, 0, 2.5), t_jitter=1.0, rpy_jitter=0.0, q_range=0.0)
    q, v = random_states(desc, n, seed=1, **kw)
    # calibrate, then run ~seconds_target
    orc.batch_rollout(q, v, dt, 20, n_threads=threads)  # first call: library load, thread start-up
    t0 = time.perf_counter()
    orc.batch_rollout(q, v, dt, 400, n_threads=threads)
    t_cal = time.perf_counter() - t0
    steps = max(20, int(400 * seconds_target / max(t_cal, 1e-6)))
    t0 = time.perf_counter()
    orc.batch_rollout(q, v, dt, steps, n_threads=threads)
    el = time.perf_counter() - t0
    return {"value": n * steps / el, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} envs x {steps} steps of {workload} (oracle, reference operation order, "
                      f"g++ -O2, std::thread x {threads}), {el:.1f} s"}, n, steps, el


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    inner = args.inner
    # each "step" is a bounded sample: enough work for ~1 s per step
    per_step = max(1.0, 100.0 / max(1, args.steps + args.warmup))
    base, n, steps, el = cpu_reference_rate(args.workload, seconds_target=min(3.0, per_step))
    rates = []
    for i in range(args.warmup + args.steps):
        r, n, steps, el = cpu_reference_rate(args.workload, seconds_target=min(3.0, per_step))
        if i >= args.warmup:
            rates.append((n * steps, el))
    total = sum(a for a, _ in rates)
    t = sum(b for _, b in rates)
    value = total / t
    cb = dict(base)
    cb["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / max(1, len(rates)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "inner_steps_per_launch": inner, "integrator": "SemiImplicitEuler",
                   "note": "reference CPU algorithm (C++ restatement; Rust toolchain absent), bounded sample per step"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="so101_contact", choices=list(WORKLOADS))
    ap.add_argument("--inner", type=int, default=128, help="fused time steps per kernel launch")
    ap.add_argument("--envs", type=int, default=0, help="environments per GPU (0 = workload default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from gorilla_physics_b200 import Integrator, MechanismState, measure_fp64_peak

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n_envs, dt, rnd = WORKLOADS[args.workload]
    if args.envs:
        n_envs = args.envs
    inner = args.inner
    mech = make_mechanism(args.workload)
    st = MechanismState(mech, n_envs, device=local_rank)
    st.randomize(0x60121114 + rank, **rnd)
    st.synchronize()
    stream = torch.cuda.ExternalStream(st.stream, device=local_rank)

    # L2 flush buffer (> 126 MB) written between timed launches, outside the event pairs
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{local_rank}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        st.step(dt, integrator=Integrator.SemiImplicitEuler, n_steps=inner)

    for _ in range(max(3, args.warmup)):
        one_step()
    barrier()

    sampler = ClockSampler(local_rank)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = st.launch_count
    barrier()
    sampler.start()
    for i in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(i & 0xFF)          # untimed L2 flush, ordered before the launch on the same stream
            starts[i].record(stream)
            one_step()
            stops[i].record(stream)
    barrier()
    launches = st.launch_count - launches0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in zip(starts, stops)]
    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * n_envs * inner * args.steps / (total_ms_max * 1e-3)

    # ---- end to end through the public API with HOST buffers: pinned H2D of q,v + rollout + D2H
    nq, nv = st.n_q, st.n_v
    q_host = torch.empty((n_envs, nq), dtype=torch.float64).pin_memory()
    v_host = torch.empty((n_envs, nv), dtype=torch.float64).pin_memory()
    q0, v0 = st.state()
    q_host.copy_(torch.from_numpy(q0))
    v_host.copy_(torch.from_numpy(v0))
    final_time = (inner - 0.5) * dt
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        st.simulate(final_time, dt, q_host.data_ptr(), v_host.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        n_done, _, _ = st.simulate(final_time, dt, q_host.data_ptr(), v_host.data_ptr())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n_envs * n_done * e2e_steps / float(te.item())
    io_bytes = n_envs * (nq + nv) * 8

    # ---- optional end-of-rollout diagnostic reduction (the only collective; outside the timed region)
    sums = torch.zeros(4, dtype=torch.float64, device=f"cuda:{local_rank}")
    st.energy_sums_device(sums.data_ptr())
    st.synchronize()
    if world > 1:
        dist.all_reduce(sums)
    flagged = int(sums[3].item())

    # ---- roofline of the step kernel
    fp64_peak = measure_fp64_peak(local_rank, 1.5) if rank == 0 else 0.0
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    per_gpu_rate = n_envs * inner * args.steps / (total_ms * 1e-3)
    flops = float(load_flop_counts().get(args.workload, 0.0))
    achieved_tf = per_gpu_rate * flops / 1e12
    alg_bytes_per_launch = n_envs * (nq + nv) * 8 * 2  # read + write q,v once per launch
    avg_launch_s = total_ms * 1e-3 / args.steps
    roofline = {
        "bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": achieved_tf / fp64_peak if fp64_peak else None,
        "traffic": load_traffic(args.workload),
        "peak_source": "gp_measure_fp64_peak: DFMA chain on this GPU in this run (MEASURED_PEAKS.json has no FP64 figure)",
        "flop_per_env_step": flops,
        "reference_formulation_flop_per_env_step": load_reference_flops(args.workload),
        "hbm": {"achieved": alg_bytes_per_launch / avg_launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes_per_launch / avg_launch_s / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
    }

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n_envs_per_gpu": n_envs, "inner_steps_per_launch": inner,
                       "dt": dt, "integrator": "SemiImplicitEuler", "kernel": mech.kernel_variant,
                       "parallelism": f"env-sharded x{world}, no collective on the step path",
                       "l2": "192 MB flush written before every timed launch (outside the event pair)",
                       "flagged_envs": flagged},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes,
                    "steps": e2e_steps, "api": "gp_batch_simulate (pinned host q,v in/out)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb, *_ = cpu_reference_rate(args.workload, seconds_target=12.0)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
