"""Runs tests/test_zz_round1_unverified_gpu.py (code written after the round-1 GPU budget was spent) in a CHILD process, so that a
faulting, never-yet-executed kernel cannot poison the CUDA context of the verified suite.  The wrapper is xfail(strict=False): XPASS
when every staged case passes on the GPU, XFAIL (with the child's report in the captured output) otherwise."""
import pytest

import test_zz_round1_unverified_gpu as staged


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="staged code: never executed on a B200 before the end of round 1")
def test_staged_suite_in_subprocess():
    r = staged._run_staged_child()
    print(r.stdout[-12000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0, "staged GPU cases failed (see captured output)"
