"""`-m gpu`: optional code paths that are OFF by default (validated on a B200 by the round-1 driver run); each check runs
in its own process so a fault cannot take the suite's CUDA context with it."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_staged_rerank_equals_default():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "rerank_staged_worker.py")], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0 and "RERANK_STAGED_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
