import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    gdir = os.path.join(ROOT, "tests", "golden")
    return {n[:-4]: np.load(os.path.join(gdir, n)) for n in os.listdir(gdir) if n.endswith(".npz")}


@pytest.fixture(scope="session")
def built_lib():
    """libb200rec.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from recsys_pytorch_b200.csrc import build as cbuild
    cbuild.build()
    from recsys_pytorch_b200 import _lib
    return _lib.lib()


@pytest.fixture(scope="session")
def oracle_c():
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True)
    from tests.util import OracleC
    return OracleC(os.path.join(ROOT, "oracle", "liboracle.so"))


@pytest.fixture(scope="session")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
