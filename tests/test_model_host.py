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


# ---------------------------------------------------------------------------------------------------------------
# F16_MATH_FAST arithmetic (csrc/f16_fast.cuh): re-associated, so compared with a tolerance instead of bit-for-bit
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emu_fast(emu):
    emu.emu_calc_xdot_fast.argtypes = [dp, dp, dp, ctypes.c_double]
    emu.emu_step_fast.argtypes = [dp, dp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.POINTER(LqrLaw),
                                  ctypes.POINTER(ctypes.c_int)]
    emu.emu_fastmath_probe.argtypes = [ctypes.c_double, ctypes.c_double, dp]
    emu.emu_step_fast_lofi.argtypes = emu.emu_step_fast.argtypes
    return emu


def test_fast_elementary_functions(emu_fast):
    import mpmath as mp
    mp.mp.prec = 200
    r = np.random.default_rng(3)
    out = np.zeros(5)
    worst = np.zeros(3)
    xs = np.concatenate([r.uniform(-np.pi / 4, np.pi / 4, 400), [0.0, np.pi / 4, -np.pi / 4, 1e-300, 1e-9]])
    for x in xs:                                    # sincos_quarter on its domain
        emu_fast.emu_fastmath_probe(float(x), 0.9, _p(out))
        worst[0] = max(worst[0], abs(out[2] - float(mp.sin(x))) / max(abs(float(mp.sin(x))), 1e-300) if x else abs(out[2]),
                       abs(out[3] - float(mp.cos(x))))
    xs = np.concatenate([r.uniform(-700, 700, 600), r.uniform(-7, 7, 400), np.arange(-8, 9) * np.pi / 2, [9.9e4, -9.9e4, 2e5]])
    for x in xs:                                    # sincos_any: absolute error (values are <= 1)
        emu_fast.emu_fastmath_probe(float(x), 0.9, _p(out))
        worst[1] = max(worst[1], abs(out[0] - float(mp.sin(mp.mpf(float(x))))), abs(out[1] - float(mp.cos(mp.mpf(float(x))))))
    alts = np.concatenate([r.uniform(0, 100000, 500), [0.0, 100000.0, 35000.0]])
    for alt in alts:                                # half_rho = 0.5 * 2.377e-3 * tfac^4.14 (nlplant.c:478)
        tfac = 1 - .703e-5 * alt
        emu_fast.emu_fastmath_probe(0.1, float(tfac), _p(out))
        ref = float(mp.mpf(0.5) * mp.mpf(2.377e-3) * mp.mpf(float(tfac)) ** mp.mpf(4.14))
        worst[2] = max(worst[2], abs(out[4] - ref) / ref)
    assert worst[0] < 3e-16 and worst[1] < 3e-16 and worst[2] < 6e-16, worst


@pytest.mark.parametrize("xcg", [0.25, 0.35])
def test_fast_calc_xdot_within_tolerance(emu_fast, oracle, xcg):
    """<= 1e-12 of max(|ref|, rms of that derivative) on random in-envelope states and on perturbed-trim states."""
    from _inputs import X_TRIM_XCG25, perturbed_trim
    from conftest import scaled_err
    xu = random_envelope_xu(3000, seed=21, hifi=True)
    r = np.random.default_rng(4)
    x = np.vstack([xu, r.uniform(-30, 30, (1, xu.shape[1]))])          # lf1
    x[2] = r.uniform(0, 60000, xu.shape[1])                            # altitude both sides of 35000 ft
    u = np.stack([r.uniform(500, 20000, 3000), r.uniform(-30, 30, 3000), r.uniform(-25, 25, 3000), r.uniform(-35, 35, 3000)])
    xp, up = perturbed_trim(3000, X_TRIM_XCG25, seed=8)
    for X, U in ((x, u), (xp, up)):
        ref, st = oracle.calc_xdot_batch(X, U, 1, xcg, PORT)
        out = np.empty_like(ref)
        for i in range(X.shape[1]):
            xd = np.zeros(18)
            s = emu_fast.emu_calc_xdot_fast(_p(np.ascontiguousarray(X[:, i])), _p(np.ascontiguousarray(U[:, i])), _p(xd), xcg)
            assert s == int(st[i])
            out[:, i] = xd
        assert (st == 0).all()
        assert scaled_err(out, ref) < 1e-12


def test_fast_step_trajectory_and_status(emu_fast, oracle):
    """10 s of Euler from perturbed trim: <= 1e-9 scaled; aircraft that leave the envelope stop at the same step with
    the same status word (or, within rounding of a threshold, one step apart)."""
    from _inputs import X_TRIM_XCG25, X_TRIM_XCG35, perturbed_trim
    from conftest import scaled_err
    for xt, xcg, n, K in ((X_TRIM_XCG25, 0.25, 6, 10000), (X_TRIM_XCG35, 0.35, 12, 6000)):
        x0, u0 = perturbed_trim(n, xt, seed=13)
        ref, st = oracle.step_batch(x0, u0, K, 0.001, 1, xcg, None, PORT)
        out = np.empty_like(ref)
        sts = np.zeros(n, dtype=np.int64)
        for i in range(n):
            x = np.ascontiguousarray(x0[:, i])
            sts[i] = emu_fast.emu_step_fast(_p(x), _p(np.ascontiguousarray(u0[:, i])), K, 0.001, xcg, None, None)
            out[:, i] = x
        assert np.array_equal(sts, st)
        alive = st == 0
        assert alive.any()
        assert scaled_err(out[:, alive], ref[:, alive]) < 1e-9
        if (~alive).any():      # frozen aircraft: stopped in the same state up to the accumulated rounding
            assert scaled_err(out[:, ~alive], ref[:, ~alive]) < 1e-6


def test_fast_closed_loop_step(emu_fast, oracle):
    from conftest import scaled_err
    g = load_golden("xcg35")
    r = np.random.default_rng(9)
    K = 0.05 * r.normal(size=(3, 9))
    sel = list(g["mpc_x_idx"])
    law = make_lqr(K, sel, g["x_trim"][sel], g["u_trim"], rows=[1, 2, 3])
    x0 = g["xs"][3].copy()
    ref, st = oracle.step_batch(x0[:, None].copy(), g["u_trim"][:, None].copy(), 300, 0.001, 1, 0.35, law)
    x = x0.copy()
    assert emu_fast.emu_step_fast(_p(x), _p(g["u_trim"].copy()), 300, 0.001, 0.35, ctypes.byref(law), None) == int(st[0])
    assert scaled_err(x[:, None], ref) < 1e-10


def test_fast_step_status_words(emu_fast, oracle):
    """Every way of stopping reports the oracle's status word and freezes the same state."""
    from _inputs import X_TRIM_XCG25
    cases = []
    for idx, val in ((2, -1.0), (2, 100001.0), (6, 901.0), (8, 31.0), (9, -301.0), (12, 999.0), (13, 25.5), (14, -21.6),
                     (15, 30.5), (16, 25.5), (16, -0.1), (7, np.deg2rad(46.0)), (7, np.deg2rad(-20.5)), (8, np.deg2rad(30.5)),
                     (3, np.nan), (17, np.nan), (6, np.nan), (0, np.nan)):
        x = X_TRIM_XCG25.copy()
        x[idx] = val
        cases.append(x)
    X = np.ascontiguousarray(np.array(cases).T)
    U = np.ascontiguousarray(np.tile(X_TRIM_XCG25[12:16][:, None], (1, X.shape[1])))
    U[1, 3] = np.nan
    ref, st = oracle.step_batch(X, U, 5, 0.001, 1, 0.25, None, PORT)
    for i in range(X.shape[1]):
        x = np.ascontiguousarray(X[:, i])
        done = ctypes.c_int(-1)
        s = emu_fast.emu_step_fast(_p(x), _p(np.ascontiguousarray(U[:, i])), 5, 0.001, 0.25, None, ctypes.byref(done))
        assert s == int(st[i]) and s != 0, (i, s, int(st[i]))
        assert done.value == 0
        assert np.array_equal(x, ref[:, i], equal_nan=True)


def test_fast_step_huge_euler_angle_takes_libm_path(emu_fast, oracle):
    """An Euler angle beyond 2^30 rad is not a stop condition of the reference: the fast loop hands over to libm trig."""
    from _inputs import X_TRIM_XCG25
    from conftest import scaled_err
    x0 = X_TRIM_XCG25.copy()
    x0[5] = 3.0e9 + 0.25
    x0[3] = 0.1
    u0 = X_TRIM_XCG25[12:16].copy()
    ref, st = oracle.step_batch(x0[:, None].copy(), u0[:, None].copy(), 200, 0.001, 1, 0.25, None, PORT)
    x = x0.copy()
    done = ctypes.c_int(-1)
    assert emu_fast.emu_step_fast(_p(x), _p(u0), 200, 0.001, 0.25, None, ctypes.byref(done)) == int(st[0]) == 0
    assert done.value == 200
    assert scaled_err(x[:, None], ref) < 1e-10   # single aircraft: scale = |ref| per element (pure relative)


def test_fast_step_on_the_actuator_limits(emu_fast, oracle):
    """Actuator states that start ON their limits with commands beyond them (the saturated branches of the fast
    arithmetic): status and state must follow the oracle."""
    from _inputs import X_TRIM_XCG25
    from conftest import scaled_err
    cases = []
    for (T, dh, da, dr, lf2), u in (((19000.0, 25.0, 21.5, 30.0, 25.0), (25000.0, 40.0, 30.0, 45.0)),
                                     ((1000.0, -25.0, -21.5, -30.0, 0.0), (0.0, -40.0, -30.0, -45.0)),
                                     ((19000.0, -25.0, 21.5, -30.0, 0.0), (500.0, 30.0, -30.0, 31.0)),
                                     ((18999.999999999996, 24.999999999999996, 0.0, 0.0, 1e-300), (19000.0, 25.0, 21.5, 30.0))):
        x = X_TRIM_XCG25.copy()
        x[12:17] = [T, dh, da, dr, lf2]
        cases.append((x, np.array(u)))
    for x0, u0 in cases:
        ref, st = oracle.step_batch(x0[:, None].copy(), u0[:, None].copy(), 1500, 0.001, 1, 0.25, None, PORT)
        x = x0.copy()
        done = ctypes.c_int(-1)
        s = emu_fast.emu_step_fast(_p(x), _p(u0.copy()), 1500, 0.001, 0.25, None, ctypes.byref(done))
        assert s == int(st[0])
        # a frozen aircraft stops at the same step unless the violated threshold is crossed within rounding
        assert scaled_err(x[:, None], ref) < 1e-7 if s else scaled_err(x[:, None], ref) < 1e-9
        assert np.all(x[12] >= 1000) and np.all(x[12] <= 19000) and abs(x[13]) <= 25 and abs(x[14]) <= 21.5 and abs(x[15]) <= 30
        assert 0 <= x[16] <= 25


@pytest.mark.parametrize("fi", [1, 0])
def test_staged_column_evaluation_bit_equal(emu, fi):
    """linearise_batch reuses the stages a perturbation does not touch: f(x + delta e_col) through the staged path
    must equal the plain _calc_xdot at the same point bit for bit, for every column, both signs, both fidelities."""
    from _inputs import X_TRIM_XCG25, perturbed_trim
    emu.emu_calc_xdot_col.argtypes = [dp, dp, ctypes.c_int, ctypes.c_double, dp, ctypes.c_int, ctypes.c_double]
    x0, u0 = perturbed_trim(24, X_TRIM_XCG25, seed=31, frac=0.04)
    x0[13, 3] = 25.0 - 1e-6       # + eps leaves DH1
    x0[8, 4] = np.deg2rad(30.0)   # + eps leaves BETA1, - eps is inside
    for n in range(x0.shape[1]):
        x, u = np.ascontiguousarray(x0[:, n]), np.ascontiguousarray(u0[:, n])
        for col in range(-1, 22):
            for delta in (1e-5, -1e-5):
                xp, up = x.copy(), u.copy()
                if 0 <= col < 18:
                    xp[col] += delta
                elif col >= 18:
                    up[col - 18] += delta
                ref, out = np.zeros(18), np.zeros(18)
                s_ref = emu.emu_calc_xdot(_p(xp), _p(up), _p(ref), fi, 0.25)
                s_out = emu.emu_calc_xdot_col(_p(x), _p(u), col, delta, _p(out), fi, 0.25)
                assert s_ref == s_out, (n, col, delta)
                assert np.array_equal(ref, out, equal_nan=True), (n, col, delta)


def test_trim_nelder_mead_bit_equal(emu, oracle):
    """the register-resident Nelder-Mead of the kernels (f16_model.cuh::nelder_mead_trim) is the oracle's, bit for bit"""
    emu.emu_trim.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, dp, dp]
    for h, V, fi, xcg in ((10000, 700, 1, 0.25), (10000, 700, 1, 0.35), (10000, 700, 0, 0.25), (5000, 300, 1, 0.35),
                          (40000, 900, 1, 0.25), (20000, 500, 0, 0.35), (36000, 450, 1, 0.25)):
        x, info = np.zeros(18), np.zeros(4)
        st = emu.emu_trim(h, V, fi, xcg, 1e-10, 50000, _p(x), _p(info))
        xr, ir, sr = oracle.trim(h, V, fi, xcg)
        assert st == sr and np.array_equal(x, xr), (h, V, fi, xcg)
        assert (info[0], int(info[1]), int(info[2]), bool(info[3])) == (ir["cost"], ir["iterations"], ir["fcalls"], ir["converged"])
    # iteration cap: same partial answer
    x, info = np.zeros(18), np.zeros(4)
    emu.emu_trim(10000, 700, 1, 0.25, 1e-10, 40, _p(x), _p(info))
    xr, ir, _ = oracle.trim(10000, 700, 1, 0.25, maxiter=40)
    assert np.array_equal(x, xr) and int(info[1]) == 40 == ir["iterations"] and not ir["converged"]


def test_trim_fixed_point_exit_changes_nothing(emu, oracle):
    """Flight conditions of the cfg-4 grid at xcg 0.25 whose search never meets xatol / fatol (the simplex collapses onto
    neighbouring floating-point numbers at a kink of the cost, and scipy spins to maxiter -- env.py:273 asks for 50 000).  The
    device search leaves such a fixed point early; point, cost, iteration and evaluation counts must be what the full loop
    gives: the host compile without the exit, and the oracle's restatement of scipy's loop."""
    emu.emu_trim_fp.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                ctypes.c_int, dp, dp]
    stuck = 0
    hs, vs = np.linspace(5000, 40000, 64), np.linspace(300, 900, 64)   # the cfg-4 grid
    for h, V in ((hs[0], vs[5]), (hs[28], vs[0]), (hs[32], vs[0]), (hs[20], vs[4]), (10000.0, 700.0)):
        xa, ia, xb, ib = np.zeros(18), np.zeros(4), np.zeros(18), np.zeros(4)
        sa = emu.emu_trim_fp(h, V, 1, 0.25, 1e-10, 6000, 1, _p(xa), _p(ia))
        sb = emu.emu_trim_fp(h, V, 1, 0.25, 1e-10, 6000, 0, _p(xb), _p(ib))
        assert sa == sb and np.array_equal(xa, xb) and np.array_equal(ia, ib), (h, V, ia, ib)
        xr, ir, sr = oracle.trim(h, V, 1, 0.25, maxiter=6000)
        assert sa == sr and np.array_equal(xa, xr)
        assert (ia[0], int(ia[1]), int(ia[2]), bool(ia[3])) == (ir["cost"], ir["iterations"], ir["fcalls"], ir["converged"])
        stuck += int(ia[1]) == 6000 and not ia[3] and ia[2] > 5 * ia[1]   # a shrink (7 evaluations) every iteration
    assert stuck == 3   # three points spin at a fixed point, one is still creeping along at the cap, 10000 ft / 700 ft/s converges


def test_fast_calc_xdot_on_the_envelope_corners(emu_fast, oracle):
    """the extreme cells of every table axis and of the tfac^4.14 table: exactly on alpha = -20 / 45 deg, beta = +-30 deg,
    dele = +-25 deg, h = 0 / 100000 ft (the cell search must land in the last cell, not beyond it), and within
    an ulp of interior breakpoints"""
    from _inputs import X_TRIM_XCG25
    from conftest import scaled_err
    d2r = np.pi / 180
    alphas = [-20 * d2r, 45 * d2r, np.nextafter(45 * d2r, 0), 0.0, 5 * d2r, np.nextafter(5 * d2r, 1), np.nextafter(5 * d2r, -1)]
    betas = [-30 * d2r, 30 * d2r, np.nextafter(30 * d2r, 0), -10 * d2r, 10 * d2r, 0.0, np.nextafter(10 * d2r, 1)]
    els = [-25.0, 25.0, -10.0, 10.0, 0.0, np.nextafter(10.0, 0), np.nextafter(-10.0, 0)]
    alts = [0.0, 100000.0, 35000.0, np.nextafter(35000.0, 0), 99999.0]
    cases = []
    for a in alphas:
        for b in betas:
            for e in els:
                x = X_TRIM_XCG25.copy()
                x[7], x[8], x[13] = a, b, e
                x[2] = alts[len(cases) % len(alts)]
                if x[7] * 180 / np.pi > 45 or abs(x[8] * 180 / np.pi) > 30:
                    continue   # rounding of the degree conversion put it outside: covered by the status tests
                cases.append(x)
    X = np.ascontiguousarray(np.array(cases).T)
    U = np.ascontiguousarray(np.tile(X_TRIM_XCG25[12:16][:, None], (1, X.shape[1])))
    ref, st = oracle.calc_xdot_batch(X, U, 1, 0.25, PORT)
    out = np.empty_like(ref)
    for i in range(X.shape[1]):
        xd = np.zeros(18)
        s = emu_fast.emu_calc_xdot_fast(_p(np.ascontiguousarray(X[:, i])), _p(np.ascontiguousarray(U[:, i])), _p(xd), 0.25)
        assert s == int(st[i]), i
        out[:, i] = xd
    ok = st == 0
    assert ok.sum() > 200
    assert scaled_err(out[:, ok], ref[:, ok]) < 1e-12

def test_fast_lofi_step_over_the_lofi_envelope(emu_fast, oracle):
    """fastmath::calc_xdot_lofi (the F16_MATH_FAST lofi step) against the oracle: alpha -20..89 deg incl. the grid nodes and
    the extrapolated ends, beta of either sign to +-29.9 deg and exactly 0, every elevator cell, both xcg, open loop and
    with a feedback law; statuses equal"""
    from _inputs import random_envelope_xu
    from conftest import scaled_err
    n = 300
    xu = random_envelope_xu(n, seed=41, hifi=False)
    r = np.random.default_rng(42)
    x0 = np.vstack([xu, r.uniform(-20, 5, (1, n))])
    x0[2] = r.uniform(1000, 39000, n)
    x0[7, :8] = np.deg2rad([-20.0, -10.0, 0.0, 45.0, 60.0, 89.0, 5.0, -5.0])
    x0[8, 8:14] = np.deg2rad([0.0, 5.0, -5.0, 29.9, -29.9, 15.0])
    x0[13, 14:20] = [-25.0, -24.0, -12.0, 0.0, 12.0, 25.0]
    x0[8, 20] = np.deg2rad(30.5)                       # outside: frozen at once with the oracle's status word
    u0 = np.vstack([r.uniform(1000, 19000, n), r.uniform(-25, 25, n), r.uniform(-21.5, 21.5, n), r.uniform(-30, 30, n)])
    g = load_golden("lofi_xcg25")
    sel = list(g["mpc_x_idx"])
    K = np.zeros((3, 9))
    K[0, [2, 5]] = [-3.0, -0.8]
    K[1, [0, 4]] = [-0.3, -0.15]
    law = make_lqr(K, sel, g["x_trim"][sel], g["u_trim"], rows=[1, 2, 3])
    for xcg in (0.25, 0.35):
        for lw in (None, law):
            ref, st = oracle.step_batch(x0.copy(), u0.copy(), 40, 0.001, 0, xcg, lw, PORT)
            out = np.empty_like(x0)
            sts = np.zeros(n, dtype=np.int64)
            for i in range(n):
                x = np.ascontiguousarray(x0[:, i])
                sts[i] = emu_fast.emu_step_fast_lofi(_p(x), _p(np.ascontiguousarray(u0[:, i])), 40, 0.001, xcg,
                                                     ctypes.byref(lw) if lw is not None else None, None)
                out[:, i] = x
            assert np.mean(sts != st) < 0.01 and sts[20] == st[20] != 0
            alive = (st == 0) & (sts == 0)
            assert alive.mean() > 0.9
            assert scaled_err(out[:, alive], ref[:, alive]) < 1e-10


@pytest.mark.parametrize("xcg", [0.25, 0.35])
def test_fast_calc_xdot_in_the_bands_around_every_breakpoint(emu_fast, oracle, xcg):
    """the fast cell search must pick the reference's cell (mexndinterp.c:97-143) everywhere, also
    within 1e-12 .. 1e-6 of a cell width above or below an interior breakpoint -- a wrong cell there extrapolates its
    neighbour and errs by delta * (slope difference) (9e-9 scaled was measured at alpha = 30 deg + 1.5e-8).  Checked
    against the reference's own .so when oracle/_ref is present."""
    from _inputs import X_TRIM_XCG25, breakpoint_band_states
    from conftest import scaled_err
    from oracle import REF
    X = breakpoint_band_states(X_TRIM_XCG25)
    U = np.ascontiguousarray(np.tile(X_TRIM_XCG25[12:16][:, None], (1, X.shape[1])))
    be = REF if oracle.open_ref() else PORT
    ref, st = oracle.calc_xdot_batch(X, U, 1, xcg, be)
    assert (st == 0).all()
    out = np.empty_like(ref)
    for i in range(X.shape[1]):
        xd = np.zeros(18)
        assert emu_fast.emu_calc_xdot_fast(_p(np.ascontiguousarray(X[:, i])), _p(np.ascontiguousarray(U[:, i])), _p(xd), xcg) == 0
        out[:, i] = xd
    scale = np.maximum(np.abs(ref), np.sqrt(np.mean(ref * ref, axis=1, keepdims=True)))
    err = np.abs(out - ref) / np.where(scale == 0, 1.0, scale)
    assert np.isfinite(out).all()
    worst = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() < 1e-12, (err.max(), worst, X[[7, 8, 13], worst[1]] * [180 / np.pi, 180 / np.pi, 1])


def test_fast_image_and_cell_search_against_the_reference_accessors(emu_fast, oracle):
    """f16_fast_probe on the host compile: the cells of locate_hifi against getHyperCube and every table of the (f, d)
    image against the reference aggregators, on random points, every breakpoint, +-1 ulp and the 1e-12 .. 1e-6 bands."""
    from _probe import check_fast_probe, probe_points
    emu_fast.emu_fast_probe.argtypes = [ctypes.c_double] * 3 + [dp, ctypes.POINTER(ctypes.c_int), dp]
    pts = probe_points()
    n = pts.shape[1]
    coef, cells, lam = np.zeros((44, n)), np.zeros((4, n), dtype=np.int32), np.zeros((4, n))
    o, c, l = np.zeros(44), (ctypes.c_int * 4)(), np.zeros(4)
    for i in range(n):
        assert emu_fast.emu_fast_probe(pts[0, i], pts[1, i], pts[2, i], _p(o), c, _p(l)) == 0
        coef[:, i], cells[:, i], lam[:, i] = o, list(c), l
    s = check_fast_probe(oracle, pts, coef, cells, lam)
    assert s["worst_coef_err"] < 1e-13, s
    # s["on_node_other_cell"]: queries on or an ulp from a breakpoint that the one-FMA map to cell units places 1e-16 to
    # its other side (check_fast_probe bounds their weight to 4 ulp of 0 / 1): the same value to rounding, never a band
    print(s)


def test_fast_integer_comparisons_on_their_thresholds(emu_fast, oracle):
    """the comparisons that f16_fast.cuh asks on the integer pipe (airspeed floor, tropopause, the five actuator rate limits) ON,
    one ulp either side of and well either side of their thresholds: same derivative as the reference to 1e-12"""
    from _inputs import X_TRIM_XCG25, comparison_threshold_cases
    from conftest import scaled_err
    X, U = comparison_threshold_cases(X_TRIM_XCG25, oracle.atmos)
    ref, st = oracle.calc_xdot_batch(X, U, 1, 0.25, PORT)
    assert not st.any() and X.shape[1] > 60
    out = np.empty_like(ref)
    for i in range(X.shape[1]):
        xd = np.zeros(18)
        assert emu_fast.emu_calc_xdot_fast(_p(np.ascontiguousarray(X[:, i])), _p(np.ascontiguousarray(U[:, i])), _p(xd), 0.25) == 0
        out[:, i] = xd
    # every rate limit is met from both sides somewhere in the set
    for row, lim in ((12, 10000.0), (13, 60.0), (14, 80.0), (15, 120.0), (16, 25.0)):
        assert (np.abs(ref[row]) == lim).any() and (np.abs(ref[row]) < lim).any(), row
    assert scaled_err(out, ref) < 1e-12
