"""Runs tests/test_zz_round1_unverified_gpu.py (code written after the round-1 GPU budget was spent) in a CHILD process, so that a
faulting, never-yet-executed kernel cannot poison the CUDA context of the verified suite.  The wrapper is xfail(strict=False): XPASS
when every staged case passes on the GPU, XFAIL (with the child's report in the captured output) otherwise."""
import time

import pytest

import conftest
import test_zz_round1_unverified_gpu as staged

BUDGET_S = 480      # the staged child runs are extras: never let them push a GPU test session towards a driver-side time limit


def _within_budget():
    if time.time() - conftest.SESSION_START > BUDGET_S:
        pytest.skip("GPU test session already ran %d s: staged child run skipped (run tools/round2_bringup.sh instead)" % BUDGET_S)


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="staged code: never executed on a B200 before the end of round 1")
def test_staged_suite_in_subprocess():
    _within_budget()
    r = staged._run_staged_child()
    print(r.stdout[-12000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0, "staged GPU cases failed (see captured output)"


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="fp16-operand build: written after the round-1 GPU budget was spent, never executed on a B200")
def test_fp16_operand_build_in_subprocess():
    """The verified GPU suites (per-kernel parity, model parity, train step) re-run against libtvts_b200_fp16.so (TVTS_OPERAND=fp16:
    IEEE-half operands + static loss scale; the torch restatements follow the operand dtype).  The loss-trajectory runs of that build
    are part of tools/round2_bringup.sh (kept out of here to bound the suite's run time)."""
    _within_budget()
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = ["tests/test_kernels_gpu.py", "tests/test_trainstep_gpu.py", "tests/test_model_gpu.py"]
    r = subprocess.run([sys.executable, "-m", "pytest", *files, "-q", "-m", "gpu", "--tb=line", "-p", "no:cacheprovider", "-rA"],
                       env=dict(os.environ, TVTS_OPERAND="fp16"), capture_output=True, text=True, timeout=600, cwd=root)
    print(r.stdout[-12000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0, "GPU suites failed against the fp16-operand build (see captured output)"
