"""Query points and the checker of f16_fast_probe (the F16_MATH_FAST table image + cell search), shared by the host-compile
test (tests/test_model_host.py) and the device test (tests/test_gpu_parity.py).  Reference: getHyperCube
(mexndinterp.c:97-143: open interval -> (j, j+1), exact hit -> (j, j)) and the accessors / aggregators of
hifi_F16_AeroData.c, through the oracle."""
import numpy as np

from _inputs import ALPHA1_BP, BAND_OFFSETS, BETA1_BP, DH1_BP, DH2_BP

AXES = (("ALPHA1", ALPHA1_BP, 0), ("BETA1", BETA1_BP, 1), ("DH1", DH1_BP, 2), ("DH2", DH2_BP, 2))
NOT_IN_FAST_IMAGE = (24, 43)   # delta_CZq_lef (nlplant.c:339 never uses it), delta_Cm_ds (constant 0)


def probe_points(seed=5, n_random=3000):
    """(alpha, beta, el)[3][n] in degrees: random points, every breakpoint of every axis exactly, +-1 ulp, and
    +-{1e-12 .. 1e-6} of a cell width on both sides (the band where a search by rounding could pick the neighbour)."""
    r = np.random.default_rng(seed)
    pts = [(r.uniform(-20, 45), r.uniform(-30, 30), r.uniform(-25, 25)) for _ in range(n_random)]
    lims = ((-20.0, 45.0), (-30.0, 30.0), (-25.0, 25.0))
    for _, bps, col in AXES:
        lo, hi = lims[col]
        for j, bp in enumerate(bps):
            w = (bps[j + 1] - bp) if j + 1 < len(bps) else (bp - bps[j - 1])
            vals = [bp, np.nextafter(bp, np.inf), np.nextafter(bp, -np.inf)]
            for o in BAND_OFFSETS:
                vals += [bp + o * w, bp - o * w]
            for v in vals:
                if not lo <= v <= hi:
                    continue
                for rep in range(3):
                    q = [r.uniform(-20, 45), r.uniform(-30, 30), r.uniform(-25, 25)]
                    if rep == 2:   # the other axes on nodes too
                        q = [float(r.choice(ALPHA1_BP)), float(r.choice(BETA1_BP)), float(r.choice(DH1_BP))]
                    q[col] = v
                    pts.append(tuple(q))
    return np.ascontiguousarray(np.array(pts).T)


def check_fast_probe(oracle, pts, coef, cells, lam):
    """cells/lam against getHyperCube, coefficients against the reference aggregators.  Returns a summary dict."""
    a, b, e = pts
    n = a.size
    ulp = 2.0 ** -52
    on_node_other_cell = 0
    for ax_i, (ax, bps, col) in enumerate(AXES):
        X = np.array(bps)
        for i in range(n):
            v = pts[col, i]
            _, lo, hi = oracle.cell(ax, v)
            k, l = int(cells[ax_i, i]), float(lam[ax_i, i])
            assert 0 <= k <= len(bps) - 2, (ax, v, k)
            ref_pos = lo if lo == hi else lo + (v - X[lo]) / (X[hi] - X[lo])
            tol = 4 * ulp * max(1.0, ref_pos)
            assert abs((k + l) - ref_pos) <= tol, (ax, v, k, l, ref_pos)          # position on the axis, to rounding
            if lo != hi and k == lo:
                continue                                                          # the reference's cell
            if lo == hi and ((k == lo and l == 0.0) or (k == lo - 1 and l == 1.0)):
                continue                                                          # exact hit: the node value
            # anything else must be a query within rounding of a breakpoint, placed on it: weight 0 or 1 to 4 ulp
            assert min(abs(l), abs(l - 1.0)) <= tol, (ax, v, k, l, lo, hi)
            assert abs(k - lo) <= 1, (ax, v, k, lo, hi)
            on_node_other_cell += 1
    ref = np.array([oracle.hifi(a[i], b[i], e[i]) for i in range(n)]).T
    scale = np.maximum(np.abs(ref), np.sqrt(np.mean(ref * ref, axis=1, keepdims=True)))
    scale = np.where(scale == 0, 1.0, scale)
    err = np.abs(coef - ref) / scale
    for s in NOT_IN_FAST_IMAGE:
        err[s] = 0.0
        assert not coef[s].any()
    assert np.isfinite(coef).all()
    return {"n": n, "worst_coef_err": float(err.max()), "worst_slot": int(np.argmax(err.max(axis=1))),
            "on_node_other_cell": on_node_other_cell}
