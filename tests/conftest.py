import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcudart.so.12")
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False
    n = ctypes.c_int(0)
    return cuda.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_names():
    names = sorted(f[:-7] for f in os.listdir(GOLDEN) if f.endswith("_in.npz"))
    assert names, "golden fixtures missing: run python tests/golden/make_golden.py"
    return names


def load_golden(name):
    import numpy as np
    from multi_agent_pkgs_b200.scenarios import Batch
    b = Batch.load(os.path.join(GOLDEN, name + "_in.npz"))
    exp = dict(np.load(os.path.join(GOLDEN, name + "_exp.npz")))
    return b, exp
