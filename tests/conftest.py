import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU parity oracle (oracle/): C restatement + the reference's own .so when oracle/_ref is present."""
    from oracle import LIB_PATH, get_oracle
    if not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle"), "port"])
    o = get_oracle()
    o.open_ref()
    return o


@pytest.fixture(scope="session")
def f16():
    """The product: libf16_b200.so through its Python binding, initialised on cuda:0."""
    import f16_mpc_oop_py_b200 as f
    f.init()
    return f


def load_golden(tag):
    return np.load(os.path.join(GOLDEN, f"env_{tag}.npz"))


@pytest.fixture(params=["xcg25", "xcg35", "lofi_xcg25"])
def golden(request):
    return load_golden(request.param)


def scaled_err(a, ref, axis=-1):
    """max |a-ref| / max(|ref|, rms of the reference along `axis`): the parity metric of SURVEY.md fact 8."""
    ref = np.asarray(ref, dtype=np.float64)
    a = np.asarray(a, dtype=np.float64)
    scale = np.sqrt(np.nanmean(ref * ref, axis=axis, keepdims=True))
    den = np.maximum(np.abs(ref), scale)
    den = np.where(den == 0, 1.0, den)
    return float(np.nanmax(np.abs(a - ref) / den))
