"""CPU checks of the DEVICE arithmetic (csrc/f16_model.cuh compiled for the host by tests/hostemu) against the
oracle: same cell search, same node-interleaved gathers, same operation order as the kernels, minus CUDA's libm.
With glibc's sin/cos/tan/pow on both sides the results must be bit-identical, which is what proves that the
restructured lookup (one search per axis, 48 distinct gathers) is the reference's arithmetic."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from _inputs import random_envelope_xu
from conftest import REPO, load_golden
from oracle import BLOB, HIFI_NAMES, PORT, LqrLaw, make_lqr

dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def emu():
    d = os.path.join(REPO, "tests", "hostemu")
    subprocess.check_call(["make", "-s", "-C", d])
    E = ctypes.CDLL(os.path.join(d, "libf16_hostemu.so"))
    E.emu_nlplant.argtypes = [dp, dp, ctypes.c_int, ctypes.c_double]
    E.emu_calc_xdot.argtypes = [dp, dp, dp, ctypes.c_int, ctypes.c_double]
    E.emu_step.argtypes = [dp, dp, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.POINTER(LqrLaw),
                           ctypes.POINTER(ctypes.c_int)]
    E.emu_hifi_probe.argtypes = [ctypes.c_double] * 3 + [dp, ctypes.POINTER(ctypes.c_int)]
    assert E.emu_init(BLOB.encode(), 0) == 0
    return E


def _p(a):
    return a.ctypes.data_as(dp)


@pytest.mark.parametrize("fi", [1, 0])
@pytest.mark.parametrize("xcg", [0.25, 0.35])
def test_nlplant_bit_equal(emu, oracle, fi, xcg):
    xu = random_envelope_xu(4000, seed=11, hifi=bool(fi))
    ref, _ = oracle.nlplant_batch(xu, fi, xcg, PORT)
    out = np.empty_like(ref)
    for i in range(xu.shape[1]):
        x = np.ascontiguousarray(xu[:, i])
        xd = np.zeros(18)
        assert emu.emu_nlplant(_p(x), _p(xd), fi, xcg) == 0
        out[:, i] = xd
    assert np.array_equal(out, ref)


def test_lookup_bit_equal_on_grid_edges(emu, oracle):
    r = np.random.default_rng(5)
    A1 = [-20.0 + 5 * i for i in range(14)]
    B1 = [-30., -25, -20, -15, -10, -8, -6, -4, -2, 0, 2, 4, 6, 8, 10, 15, 20, 25, 30]
    D1 = [-25., -10, 0, 10, 25]
    pts = [(r.uniform(-20, 45), r.uniform(-30, 30), r.uniform(-25, 25)) for _ in range(3000)]

    def around(v, lo, hi):
        out = [v]
        if v < hi:
            out.append(np.nextafter(v, np.inf))
        if v > lo:
            out.append(np.nextafter(v, -np.inf))
        return out

    for a in A1:
        for aa in around(a, -20, 45):
            pts.append((aa, float(r.choice(B1)), float(r.choice(D1))))
            pts.append((aa, r.uniform(-30, 30), r.uniform(-25, 25)))
    for b in B1:
        for bb in around(b, -30, 30):
            pts.append((r.uniform(-20, 45), bb, r.uniform(-25, 25)))
    for d in D1:
        for dd in around(d, -25, 25):
            pts.append((r.uniform(-20, 45), r.uniform(-30, 30), dd))
    for a, b, e in pts:
        ref = oracle.hifi(a, b, e)
        out = np.zeros(44)
        cl = (ctypes.c_int * 8)()
        assert emu.emu_hifi_probe(a, b, e, _p(out), cl) == 0
        bad = [HIFI_NAMES[i] for i in range(44) if ref[i] != out[i]]
        assert not bad, (a, b, e, bad)
        cells = []
        for ax, v in (("ALPHA1", a), ("BETA1", b), ("DH1", e), ("DH2", e)):
            _, lo, hi = oracle.cell(ax, v)
            cells += [lo, hi]
        assert list(cl) == cells, (a, b, e)


def test_calc_xdot_and_step_bit_equal(emu, oracle, golden):
    fi, xcg = int(golden["fi"]), float(golden["xcg"])
    for x, u, xd_ref in zip(golden["xs"], golden["us"], golden["xdots"]):
        xd = np.zeros(18)
        assert emu.emu_calc_xdot(_p(np.ascontiguousarray(x)), _p(np.ascontiguousarray(u)), _p(xd), fi, xcg) == 0
        assert np.array_equal(xd, xd_ref)
    x = golden["x_trim"].copy()
    u = golden["u_trim"].copy()
    for i in range(1, 5):
        done = ctypes.c_int()
        assert emu.emu_step(_p(x), _p(u), 500, 0.001, fi, xcg, None, ctypes.byref(done)) == 0 and done.value == 500
        assert np.array_equal(x, golden["traj_x"][i])


def test_closed_loop_step_bit_equal(emu, oracle):
    g = load_golden("xcg35")
    r = np.random.default_rng(9)
    K = 0.05 * r.normal(size=(3, 9))
    sel = list(g["mpc_x_idx"])
    law = make_lqr(K, sel, g["x_trim"][sel], g["u_trim"], rows=[1, 2, 3])
    x0 = g["xs"][3].copy()
    ref, st = oracle.step_batch(x0[:, None].copy(), g["u_trim"][:, None].copy(), 300, 0.001, 1, 0.35, law)
    x = x0.copy()
    assert emu.emu_step(_p(x), _p(g["u_trim"].copy()), 300, 0.001, 1, 0.35, ctypes.byref(law), None) == int(st[0])
    assert np.array_equal(x, ref[:, 0])


def test_envelope_status_equal(emu, oracle):
    xu = random_envelope_xu(64, seed=2)
    xu[7, :16] = np.deg2rad(np.linspace(45.0001, 90, 16))
    xu[7, 16:24] = np.deg2rad(-20.5)
    xu[8, 24:32] = np.deg2rad(30.01)
    xu[13, 32:40] = 25.5
    xu[3, 40] = np.nan
    xu[7, 41] = np.nan
    for fi in (1, 0):
        ref, st = oracle.nlplant_batch(xu, fi, 0.25, PORT)
        for i in range(64):
            xd = np.zeros(18)
            s = emu.emu_nlplant(_p(np.ascontiguousarray(xu[:, i])), _p(xd), fi, 0.25)
            assert s == int(st[i])
            assert np.array_equal(xd, ref[:, i], equal_nan=True)
