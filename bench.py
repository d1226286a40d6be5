#!/usr/bin/env python
"""Headline benchmark: env-steps/s of the batched articulated-dynamics stepper (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm on the host cores

A bench "step" is one launch of the step kernel: `--inner` fused time steps (semi-implicit Euler) of every
environment of the workload. Workload (default): SO-101 6-DOF arm with ground contact, 262144 environments per
GPU (the configuration BASELINE.json's target is quoted on); the other
configurations are `--workload` names of gorilla_physics_b200/workloads.py. Environments are independent, so
ranks shard them with no collective on the step path (weak scaling: per-GPU work fixed; `--envs-total` fixes
the total instead: strong scaling); one optional all-reduce of energy sums happens after the timed region.
Prints ONE JSON line on rank 0.

What the line holds beyond the contract's keys:
  config.contact_active   fraction of contact points / environments in contact at the start and the end of
                          the timed region (how contact-rich the timed trajectory is)
  sustained               the K-step block repeated from the same states until >= --sustain seconds of kernel
                          time: rate of every block, clocks over the whole stretch
  roofline.peak_*         DFMA-chain probe: burst (first 50 ms) and sustained (last quarter of 1.5 s) with the
                          SM clock sampled during the probe, and 148 SM x 128 flop/clk at that clock;
                          `peak` is the LARGEST of them (the most conservative denominator)
  cpu_baseline.parity_sample  the first 4096 environments of the bench's own device-randomised states against
                          the oracle: vdot at the initial state, and the state after the first launch (`inner`
                          steps through the very launch configuration that is timed)
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
PARITY_SAMPLE = 4096


# FP64 work per environment time step of the kernel that runs each workload: 2*DFMA + DADD + DMUL
# thread-level instructions per env-step as counted by ncu
# (smsp__sass_thread_inst_executed_op_d{fma,add,mul}_pred_on.sum summed over the 40 timed launches of
# this file's default command, divided by n_envs * inner steps * 40: contact workloads execute more
# once their bodies lie on the ground, tools/ncu_flops_over_bench.py). profiles/flop_counts.json holds
# the numbers with the capture they come from; DESIGN.md section 4 explains why this (executed, not
# reference-formulation) count is used.
def load_flop_counts():
    try:
        return json.loads((ROOT / "profiles" / "flop_counts.json").read_text())["flop_per_env_step"]
    except Exception:
        return {}


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


# ---- workload table without the product library (reference arm / cpu_baseline leg) -------------------
def frozen_desc(workload):
    """the workload's mechanism description from tests/golden/workload_descs.json (tools/make_workload_descs.py;
    SO-101 and navbot entries are checked against the literals extracted from the reference sources): the CPU
    arm does not take its mechanism from the library it is compared with"""
    from gorilla_physics_b200.desc import MechanismDesc
    table = json.loads((ROOT / "tests" / "golden" / "workload_descs.json").read_text())
    entry = table[workload]
    if "same_as" in entry:
        entry = table[entry["same_as"]]
    return MechanismDesc.from_arrays(**entry)


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


# ---- the reference's CPU algorithm ---------------------------------------------------------------------
def cpu_states(w, desc, orc, n, threads, seed=1):
    """n initial states of the workload for the CPU arm: host_states of the workload's distribution; "resting"
    workloads are settled with the oracle itself (a bounded set of distinct environments, tiled)"""
    from gorilla_physics_b200.workloads import host_states
    if w.settle_steps:
        m = min(n, 64 * threads)
        q, v = host_states(desc, w.randomize, m, seed)
        q, v = orc.batch_rollout(q, v, w.dt, w.settle_steps, controller=int(w.controller), params=tuple(w.ctrl_params),
                                 n_threads=threads)
        reps = -(-n // m)
        return np.tile(q, (reps, 1))[:n], np.tile(v, (reps, 1))[:n]
    return host_states(desc, w.randomize, n, seed)


def cpu_reference_rate(workload, seconds_target=12.0, threads=None, n_envs=None, inner=128):
    """The reference's CPU algorithm (oracle port, reference operation order) on a bounded sample: `n_envs`
    environments (default 64 per thread) advanced by as many time steps as fit into `seconds_target`."""
    from gorilla_physics_b200.workloads import WORKLOADS
    from oracle.binding import OracleMechanism
    w = WORKLOADS[workload]
    desc = frozen_desc(workload)
    orc = OracleMechanism(desc)
    threads = threads or os.cpu_count() or 1
    n = n_envs or 64 * threads
    q, v = cpu_states(w, desc, orc, n, threads)
    roll = dict(controller=int(w.controller), params=tuple(w.ctrl_params), n_threads=threads)
    # calibrate on a small set, then run about seconds_target
    nc = min(n, 64 * threads)
    orc.batch_rollout(q[:nc], v[:nc], w.dt, 20, **roll)  # first call: library load, thread start-up
    t0 = time.perf_counter()
    orc.batch_rollout(q[:nc], v[:nc], w.dt, 200, **roll)
    per_env_step = (time.perf_counter() - t0) / (nc * 200)
    steps = int(seconds_target / max(per_env_step * n, 1e-12))
    steps = max(1, min(steps, inner)) if n_envs else max(20, steps)
    t0 = time.perf_counter()
    orc.batch_rollout(q, v, w.dt, steps, **roll)
    el = time.perf_counter() - t0
    return {"value": n * steps / el, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} envs x {steps} steps of {workload} (oracle: C++ restatement of the reference in its own "
                      f"operation order, g++ -O2, std::thread x {threads}), {el:.1f} s"}, n, steps, el


def workload_config(w, n_envs, inner, world, extra=None):
    cfg = {"workload": w.name, "baseline_config": w.config, "n_envs_per_gpu": n_envs, "inner_steps_per_launch": inner,
           "dt": w.dt, "integrator": "SemiImplicitEuler", "controller": w.controller.name,
           "settle_steps_before_timing": w.settle_steps}
    if extra:
        cfg.update(extra)
    return cfg


def run_reference(args):
    """--impl reference: every "step" advances ALL n_envs environments of the workload by `inner` time steps on
    the host cores (when that fits the time budget: otherwise by fewer time steps, stated in `sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gorilla_physics_b200.workloads import WORKLOADS
    w = WORKLOADS[args.workload]
    n_envs = args.envs or w.n_envs
    inner = args.inner
    budget = 150.0 / max(1, args.steps + args.warmup)  # seconds per step so that the whole run stays under ~3 min
    rates, last = [], None
    for i in range(args.warmup + args.steps):
        cb, n, steps, el = cpu_reference_rate(args.workload, seconds_target=budget, n_envs=n_envs, inner=inner)
        last = cb
        if i >= args.warmup:
            rates.append((n * steps, el))
    total = sum(a for a, _ in rates)
    t = sum(b for _, b in rates)
    value = total / t
    cb = dict(last)
    cb["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / max(1, len(rates)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(w, n_envs, inner, 1, {
            "note": "reference CPU algorithm (C++ restatement; Rust toolchain absent); mechanism from "
                    "tests/golden/workload_descs.json, states from numpy (the product library is not loaded)"}),
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def parity_sample(workload, inner, q0, v0, vdot0, q1, v1):
    """The bench's own states against the oracle (cpu_baseline leg, outside every timed region): vdot at the
    initial states (north_star: 1e-10 relative) and the state after the first launch of `inner` fused steps."""
    from gorilla_physics_b200.workloads import WORKLOADS
    from oracle.binding import OracleMechanism
    w = WORKLOADS[workload]
    orc = OracleMechanism(frozen_desc(workload))
    out = {"n_envs": int(len(q0)), "states": "gp_batch_randomize on the device (the timed batch's own first environments)"}
    if w.controller.name == "NONE":
        ref, _ = orc.batch_dynamics(q0, v0)
        scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1e-9)
        out["vdot_rel_err_max"] = float((np.abs(vdot0 - ref) / scale).max())
        comp = np.abs(vdot0 - ref) / np.maximum(np.abs(ref), 1e-3 * scale)
        out["vdot_componentwise_rel_err_max"] = float(comp.max())
    # one time step (controller included) from the same states, on a small batch of its own: north_star's per-step bar
    from gorilla_physics_b200 import Integrator, MechanismState
    one = MechanismState(w.mechanism(), len(q0))
    one.update(q0, v0)
    one.step(w.dt, integrator=Integrator.SemiImplicitEuler, n_steps=1, controller=w.controller, ctrl_params=tuple(w.ctrl_params))
    qg, vg = one.state()
    q1r, v1r = orc.batch_rollout(q0, v0, w.dt, 1, controller=int(w.controller), params=tuple(w.ctrl_params))

    def rel(a, b):
        return float((np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-9)).max())
    out["one_step_rel_err_max"] = {"q": rel(qg, q1r), "v": rel(vg, v1r)}
    del one
    qr, vr = orc.batch_rollout(q0, v0, w.dt, inner, controller=int(w.controller), params=tuple(w.ctrl_params))
    qp, vp = orc.batch_rollout(q0 * (1.0 + 1e-15), v0 * (1.0 - 1e-15), w.dt, inner, controller=int(w.controller),
                               params=tuple(w.ctrl_params))

    def errs(a, b):
        return np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-6)
    e = np.maximum(errs(q1, qr), errs(v1, vr))
    s = np.maximum(errs(qp, qr), errs(vp, vr))
    ok = np.isfinite(e) & np.isfinite(s)
    out.update({"rollout_steps": inner,
                "rollout_rel_err": {"median": float(np.median(e[ok])), "p90": float(np.quantile(e[ok], 0.9)),
                                    "max": float(e[ok].max())},
                "oracle_sensitivity_to_1e-15_perturbation": {"median": float(np.median(s[ok])),
                                                            "p90": float(np.quantile(s[ok], 0.9)), "max": float(s[ok].max())},
                "non_finite_envs": int((~ok).sum())})
    return out


def fp64_peak_report(device):
    """DFMA-chain probe with the SM clock sampled while it runs"""
    from gorilla_physics_b200 import measure_fp64_peak_trace
    sampler = ClockSampler(device)
    sampler.start()
    t0 = time.perf_counter()
    t, f = measure_fp64_peak_trace(device, 1.5)
    wall = time.perf_counter() - t0
    clk = sampler.stop()
    # clock samples are wall-clock stamped at 5 ms: attribute them to the probe by position in the run
    sm = [float(s[0]) for s in sampler.samples if s[0].replace(".", "").isdigit()]
    k = max(1, len(sm) // 4)
    burst_n = max(1, int(np.searchsorted(t, 0.05)))
    tail = f[len(f) - max(1, len(f) // 4):]
    sm_count = 148
    nominal = lambda mhz: sm_count * 128 * mhz * 1e6 / 1e12  # 64 DFMA/clk/SM
    rep = {"burst": float(np.max(f[:burst_n])), "sustained": float(np.median(tail)), "best": float(f.max()),
           "seconds": float(t[-1]), "wall_s": wall, "launches": int(len(f)),
           "sm_mhz_first_quarter": float(np.median(sm[:k])) if sm else None,
           "sm_mhz_last_quarter": float(np.median(sm[-k:])) if sm else None,
           "sm_mhz_min": clk["sm_min_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "power_w_max": clk["power_w_max"],
           "reasons": clk["reasons"]}
    if sm:
        rep["nominal_at_probe_clock"] = nominal(float(np.median(sm[-k:])))
    if clk["sm_max_mhz"]:
        rep["nominal_at_max_clock"] = nominal(clk["sm_max_mhz"])
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="so101_contact")
    ap.add_argument("--inner", type=int, default=128, help="fused time steps per kernel launch")
    ap.add_argument("--envs", type=int, default=0, help="environments per GPU (0 = workload default)")
    ap.add_argument("--envs-total", type=int, default=0,
                    help="environments over ALL GPUs, split evenly: strong scaling (overrides --envs)")
    ap.add_argument("--sustain", type=float, default=2.0,
                    help="repeat the timed block until this many seconds of kernel time (0: only the K steps)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "jit", "generic", "jit_twin", "generic_twin"],
                    help="jit: run the mechanism on a kernel compiled at run time for its tree (NVRTC); *_twin: the "
                         "same physics with a massless fixed leaf appended, so that no shipped specialisation matches")
    ap.add_argument("--per-step-control", action="store_true",
                    help="also measure the per-step torque paths (launch per step; streamed torque sequence)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    from gorilla_physics_b200.workloads import WORKLOADS
    if args.workload not in WORKLOADS:
        raise SystemExit(f"unknown workload {args.workload}; one of {', '.join(WORKLOADS)}")
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from gorilla_physics_b200 import Integrator, KernelMode, MechanismState

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w = WORKLOADS[args.workload]
    n_envs = args.envs or w.n_envs
    scaling = "weak"
    if args.envs_total:
        from gorilla_physics_b200 import shard_range
        lo, hi = shard_range(args.envs_total, rank, world)
        n_envs = hi - lo
        scaling = "strong"
    inner = args.inner
    mech = w.mechanism()
    if args.kernel in ("jit_twin", "generic_twin"):
        from gorilla_physics_b200 import FIXED, Mechanism
        d = mech.desc()
        d.add_body(d.n_bodies, FIXED, moment=np.zeros((3, 3)), mass=0.0)
        mech = Mechanism.from_desc(d, kernel=KernelMode.JIT if args.kernel == "jit_twin" else KernelMode.GENERIC)
    elif args.kernel == "jit":
        mech.set_kernel_mode(KernelMode.JIT)
    elif args.kernel == "generic":
        mech.set_kernel_mode(KernelMode.GENERIC)
    st = MechanismState(mech, n_envs, device=local_rank)
    ctrl = dict(controller=w.controller, ctrl_params=tuple(w.ctrl_params))
    seed = 0x60121114 + rank

    def reset_states():
        st.randomize(seed, **w.randomize)
        if w.settle_steps:
            st.step(w.dt, integrator=Integrator.SemiImplicitEuler, n_steps=w.settle_steps, **ctrl)
        st.clear_status()

    reset_states()
    st.synchronize()
    stream = torch.cuda.ExternalStream(st.stream, device=local_rank)

    # L2 flush buffer (> 126 MB) written between timed launches, outside the event pairs
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{local_rank}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        st.step(w.dt, integrator=Integrator.SemiImplicitEuler, n_steps=inner, **ctrl)

    def contact_active():
        if mech.n_contact_points == 0 or mech.n_halfspaces == 0:
            return None
        _, cf = st.dynamics(tau=None, contact_forces=True)
        hit = np.abs(cf).reshape(n_envs, -1, 3).max(axis=2) > 0.0
        return {"points": float(hit.mean()), "envs": float(hit.any(axis=1).mean())}

    # parity sample: the batch's own first environments before and after the first (warm-up) launch
    ns = min(PARITY_SAMPLE, n_envs)
    sample = None
    take_sample = rank == 0 and world == 1 and not args.no_cpu_baseline
    if take_sample:
        q_all, v_all = st.state()
        vdot_all = st.dynamics(tau=None) if w.controller.name == "NONE" else None
        sample = [q_all[:ns].copy(), v_all[:ns].copy(), None if vdot_all is None else vdot_all[:ns].copy()]
        del q_all, v_all, vdot_all
    contact_start = contact_active() if rank == 0 else None

    n_warm = max(3, args.warmup)
    one_step()
    if take_sample:
        q_all, v_all = st.state()
        sample += [q_all[:ns].copy(), v_all[:ns].copy()]
        del q_all, v_all
    for _ in range(n_warm - 1):
        one_step()
    barrier()

    def timed_block():
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        for i in range(args.steps):
            with torch.cuda.stream(stream):
                flush.fill_(i & 0xFF)          # untimed L2 flush, ordered before the launch on the same stream
                starts[i].record(stream)
                one_step()
                stops[i].record(stream)
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in zip(starts, stops)]

    sampler = ClockSampler(local_rank)
    launches0 = st.launch_count
    barrier()
    sampler.start()
    step_ms = timed_block()
    barrier()
    launches = st.launch_count - launches0
    clocks = sampler.stop()
    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    n_total = torch.tensor([float(n_envs)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(n_total)
    envs_all_ranks = int(n_total.item())
    value = envs_all_ranks * inner * args.steps / (total_ms_max * 1e-3)
    contact_end = contact_active() if rank == 0 else None

    # ---- sustained: the same block again and again from the same initial states (identical work)
    sustained = None
    if args.sustain > 0:
        s2 = ClockSampler(local_rank)
        s2.start()
        block_rates, kernel_s, wall0 = [], 0.0, time.perf_counter()
        while kernel_s < args.sustain and len(block_rates) < 2000:
            reset_states()
            for _ in range(n_warm):
                one_step()
            ms = timed_block()
            kernel_s += sum(ms) * 1e-3 * (1.0 + n_warm / len(ms))
            block_rates.append(n_envs * inner * args.steps / (sum(ms) * 1e-3))
        c2 = s2.stop()
        sustained = {"seconds_of_step_kernels": kernel_s, "wall_s": time.perf_counter() - wall0, "blocks": len(block_rates),
                     "per_gpu_value_first": block_rates[0], "per_gpu_value_last": block_rates[-1],
                     "per_gpu_value_min": min(block_rates), "per_gpu_value_median": float(np.median(block_rates)),
                     "clocks": c2,
                     "how": "rank 0: randomize (same seed) + warm-up + the K timed launches, repeated; L2 flush "
                            "before every timed launch"}
        barrier()

    # ---- end to end through the public API with HOST buffers: pinned H2D of q,v + rollout + D2H
    reset_states()
    nq, nv = st.n_q, st.n_v
    q_host = torch.empty((n_envs, nq), dtype=torch.float64).pin_memory()
    v_host = torch.empty((n_envs, nv), dtype=torch.float64).pin_memory()
    q0, v0 = st.state()
    q_host.copy_(torch.from_numpy(q0))
    v_host.copy_(torch.from_numpy(v0))
    final_time = (inner - 0.5) * w.dt
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        st.simulate(final_time, w.dt, q_host.data_ptr(), v_host.data_ptr(), **ctrl)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        n_done, _, _ = st.simulate(final_time, w.dt, q_host.data_ptr(), v_host.data_ptr(), **ctrl)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = envs_all_ranks * n_done * e2e_steps / float(te.item())
    io_bytes = n_envs * (nq + nv) * 8

    # ---- per-step control (the reference's closure-per-step contract, simulate.rs:87-112) with host torques
    per_step = None
    if args.per_step_control and rank == 0 and w.controller.name == "NONE":
        per_step = {}
        K = 128
        tau_seq = torch.zeros((K, n_envs, nv), dtype=torch.float64).pin_memory()
        tau_np = tau_seq.numpy()
        tau_np[:] = np.random.default_rng(3).uniform(-0.1, 0.1, size=(K, 1, nv))
        # (a) launch-bound floor: set_tau (H2D + layout change) + one step launch per time step
        st.update_from_host_ptr(q_host.data_ptr(), v_host.data_ptr())
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for s in range(K):
                st.step(w.dt, tau=tau_np[s], n_steps=1)
            st.synchronize()
            el = time.perf_counter() - t0
        per_step["launch_per_step"] = {"value": n_envs * K / el, "unit": UNIT, "us_per_time_step": 1e6 * el / K,
                                       "h2d_bytes_per_time_step": n_envs * nv * 8,
                                       "api": "gp_batch_set_tau + gp_batch_step(n_steps=1) per time step"}
        # (b) a horizon of torques streamed in one call
        for Kb in (16, 128):
            for rep in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                st.step_tau_sequence(w.dt, tau_seq.data_ptr(), n_steps=Kb)
                el = time.perf_counter() - t0
            per_step[f"tau_sequence_{Kb}"] = {"value": n_envs * Kb / el, "unit": UNIT, "us_per_time_step": 1e6 * el / Kb,
                                              "h2d_bytes_per_time_step": n_envs * nv * 8,
                                              "api": "gp_batch_step_tau_sequence (pinned host [K][n_envs][n_v])"}
        # (c) the torques already on the device as planes
        dev = torch.zeros((K, nv, st.ld), dtype=torch.float64, device=f"cuda:{local_rank}")
        torch.cuda.synchronize()
        for rep in range(2):
            t0 = time.perf_counter()
            st.step_tau_sequence_device(w.dt, dev.data_ptr(), K)
            st.synchronize()
            el = time.perf_counter() - t0
        per_step["tau_sequence_device_128"] = {"value": n_envs * K / el, "unit": UNIT, "us_per_time_step": 1e6 * el / K,
                                               "api": "gp_batch_step_tau_sequence_device ([K][n_v][ld] planes in HBM)"}
        del tau_seq, dev

    # ---- optional end-of-rollout diagnostic reduction (the only collective; outside the timed region)
    #      inside the library: gp_batch_reduce_diagnostics = four sums on the device + ncclAllReduce on the batch's
    #      stream; torch.distributed only carries the communicator's 128-byte id to the other ranks
    comm = None
    if world > 1:
        from gorilla_physics_b200 import Communicator
        ident = [Communicator.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        comm = Communicator(rank, world, ident[0], local_rank)
    sums = st.reduce_diagnostics(comm)
    if comm is not None:
        comm.close()
    flagged = int(sums[3])

    # ---- roofline of the step kernel
    peak = fp64_peak_report(local_rank) if rank == 0 else {}
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    per_gpu_rate = n_envs * inner * args.steps / (total_ms * 1e-3)
    flops = float(load_flop_counts().get(w.flops_key or w.name, 0.0))
    achieved_tf = per_gpu_rate * flops / 1e12
    alg_bytes_per_launch = n_envs * (nq + nv) * 8 * 2  # read + write q,v once per launch
    avg_launch_s = total_ms * 1e-3 / args.steps
    roofline = None
    if rank == 0:
        candidates = [peak.get("burst"), peak.get("sustained"), peak.get("nominal_at_probe_clock")]
        fp64_peak = max(c for c in candidates if c)
        roofline = {
            "bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved_tf / fp64_peak if (fp64_peak and flops) else None,
            "traffic": load_traffic(w.flops_key or w.name),
            "peak_source": "largest of: DFMA-chain probe on this GPU in this run (burst, sustained) and 148 SM x 128 "
                           "flop/clk at the SM clock sampled during the probe (MEASURED_PEAKS.json has no FP64 figure)",
            "peak_probe": peak,
            "frac_of_nominal_at_max_clock": (achieved_tf / peak["nominal_at_max_clock"]) if (peak.get("nominal_at_max_clock") and flops) else None,
            "flop_per_env_step": flops or None,
            "reference_formulation_flop_per_env_step": load_reference_flops(w.flops_key or w.name),
            "hbm": {"achieved": alg_bytes_per_launch / avg_launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes_per_launch / avg_launch_s / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        }

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(w, n_envs, inner, world, {
                "envs_all_ranks": envs_all_ranks, "kernel": mech.kernel_variant,
                "mapping": "warp pairs (two warps per 32 environments, half the tree each)" if st.step_lanes == 2 else "thread per environment",
                "parallelism": f"env-sharded x{world}, no collective on the step path",
                "l2": "192 MB flush written before every timed launch (outside the event pair)",
                "contact_active": {"start": contact_start, "end": contact_end},
                "flagged_envs": flagged,
                "diagnostic_sums_all_ranks": {"kinetic": float(sums[0]), "potential": float(sums[1]), "spring": float(sums[2])}}),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes,
                    "steps": e2e_steps, "api": "gp_batch_simulate (pinned host q,v in/out)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if sustained:
            line["sustained"] = sustained
        if per_step:
            line["per_step_control"] = per_step
        if world == 1 and not args.no_cpu_baseline:
            cb, *_ = cpu_reference_rate(args.workload, seconds_target=12.0)
            if sample is not None:
                cb["parity_sample"] = parity_sample(args.workload, inner, *sample)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
