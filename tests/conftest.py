import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"  # present in the build container only, never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the live reference checkout (build container only)")


def pytest_collection_modifyitems(config, items):
    have_gpu = torch.cuda.is_available()
    have_ref = os.path.isdir(os.path.join(REFERENCE, "networks"))
    for item in items:
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="reference checkout not present"))


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale) and load libdmvs_b200.so.  nvcc cross-compiles without a GPU."""
    from dmvsnet_b200 import build, _native
    build.build()
    return _native.load()


@pytest.fixture(scope="session")
def c_oracle():
    import ctypes
    from oracle import build_oracle
    return ctypes.CDLL(build_oracle.build())


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def rel_linf(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
