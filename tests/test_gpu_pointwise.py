"""`-m gpu`: the fused pointwise-MF step (csrc/pointwise_step.cu) against the oracle and the
reference's golden vectors (first validated on a B200 by the round-1 driver run; runs in its own process)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pointwise_step_matches_oracle_and_reference_golden():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "pointwise_worker.py")], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0 and "POINTWISE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
