#!/usr/bin/env python
"""What the library is for: many independent copies of one mechanism advanced together. The headline configuration of
BASELINE.json - the SO-101 arm (builders/mod.rs:252-341) above ground contact, 262 144 environments, dt = 1/6000 -
through the Python mirror of the C ABI: randomise on the device, run fused steps (128 per launch, the state stays in
registers in between), read diagnostics. Needs a B200; without a GPU the first device call raises GP_ERR_NO_DEVICE.

    python examples/batched_so101.py [n_envs] [launches]
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np  # noqa: E402

import gorilla_physics_b200 as gp  # noqa: E402


def main():
    n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    launches = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    w = gp.WORKLOADS["so101_contact"]
    mech = w.mechanism()  # Mechanism.from_model("so101") + five contact points + the ground halfspace
    print(f"{mech.n_bodies} bodies, {mech.n_v} dof, {mech.n_contact_points} contact points, kernel {mech.kernel_variant}")
    state = gp.MechanismState(mech, n_envs, device=0)
    state.randomize(seed=1, **w.randomize)
    inner = 128
    state.step(w.dt, n_steps=inner)  # warm-up launch
    state.synchronize()
    t0 = time.perf_counter()
    for _ in range(launches):
        state.step(w.dt, integrator=gp.Integrator.SemiImplicitEuler, n_steps=inner)  # asynchronous: only enqueues
    state.synchronize()
    dt_wall = time.perf_counter() - t0
    print(f"{n_envs * inner * launches / dt_wall:.3e} env-steps/s (wall clock around {launches} launches of {inner} fused steps)")
    ke, pe, _ = state.energies()
    flagged = int(np.count_nonzero(state.status()))
    q, v = state.state()  # [n_envs, n_q], [n_envs, n_v], the reference's flat joint order
    print(f"mean kinetic energy {ke.mean():.4g} J, mean potential energy {pe.mean():.4g} J, {flagged} flagged environments, "
          f"largest joint speed {np.abs(v).max():.3g} rad/s")


if __name__ == "__main__":
    try:
        main()
    except gp._abi.GorillaError as e:  # no CPU path: say so and stop
        print(f"GorillaError {e.code}: {e}", file=sys.stderr)
        sys.exit(2)
