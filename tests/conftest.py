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
    """max |a-ref| / max(|ref|, rms of the reference along `axis`): the parity metric of SURVEY.md fact 8.
    Non-finite entries must sit in the same places in both arrays (a stray NaN in the product is a failure, not a
    dropped sample); the metric is taken over the entries finite in both."""
    ref = np.asarray(ref, dtype=np.float64)
    a = np.asarray(a, dtype=np.float64)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    fin_a, fin_r = np.isfinite(a), np.isfinite(ref)
    assert np.array_equal(fin_a, fin_r), f"{int((fin_a != fin_r).sum())} entries are finite in one array and not in the other"
    if not fin_r.any():
        return 0.0
    r0 = np.where(fin_r, ref, 0.0)
    cnt = np.maximum(fin_r.sum(axis=axis, keepdims=True), 1)
    scale = np.sqrt((r0 * r0).sum(axis=axis, keepdims=True) / cnt)
    den = np.maximum(np.abs(r0), scale)
    den = np.where(den == 0, 1.0, den)
    err = np.where(fin_r, np.abs(np.where(fin_a, a, 0.0) - r0) / den, 0.0)
    return float(err.max())
