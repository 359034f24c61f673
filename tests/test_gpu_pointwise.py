"""`-m gpu`: the fused pointwise-MF step (csrc/pointwise_step.cu) against the oracle and the
reference's golden vectors.  The kernel was written after this round's GPU budget was spent: it compiles for sm_100a and
its oracle is pinned on the CPU, but it has NOT run on hardware yet - hence the separate process and the non-strict xfail
(an XPASS in the log is its first device validation)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.xfail(strict=False, reason="csrc/pointwise_step.cu has not run on hardware yet (written with no GPU minutes left)")
def test_pointwise_step_matches_oracle_and_reference_golden():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "pointwise_worker.py")], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0 and "POINTWISE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
