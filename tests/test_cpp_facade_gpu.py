"""Runs tests/cpp/test_reference_suite (built by __graft_entry__.build()): the reference's own
known-answer tests restated against the C++ facade include/gorilla_b200.hpp, executed on the GPU
through the C ABI with n_envs = 1."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "tests" / "cpp" / "test_reference_suite"


@pytest.mark.gpu
def test_cpp_reference_suite():
    deps = [ROOT / "include" / "gorilla_b200.h", ROOT / "include" / "gorilla_b200.hpp",
            ROOT / "tests" / "cpp" / "test_reference_suite.cpp"]
    if not BIN.exists() or any(d.stat().st_mtime > BIN.stat().st_mtime for d in deps):
        import __graft_entry__
        __graft_entry__.build()
    out = subprocess.run([str(BIN)], capture_output=True, text=True, timeout=600)
    print(out.stdout)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-1000:]
    assert "0 failed" in out.stdout


@pytest.mark.gpu
def test_c_host_sharded_batches_and_torque_sequence():
    """tests/cpp/test_sharded.c (plain C, gcc): gp_sharded_* over three shards bitwise equal to one batch,
    gp_sharded_simulate, diagnostic sums, gp_batch_step_tau_sequence == set_tau + step per time step.
    On a box with several GPUs the shards go to distinct devices."""
    import torch
    binary = ROOT / "tests" / "cpp" / "test_sharded"
    src = ROOT / "tests" / "cpp" / "test_sharded.c"
    if not binary.exists() or src.stat().st_mtime > binary.stat().st_mtime:
        import __graft_entry__
        __graft_entry__.build()
    n = torch.cuda.device_count()
    devices = [str(d) for d in range(n)] if n > 1 else ["0", "0", "0"]
    out = subprocess.run([str(binary)] + devices, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "sharded-ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
