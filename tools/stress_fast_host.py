#!/usr/bin/env python3
"""Judge-style probe of the F16_MATH_FAST arithmetic WITHOUT a GPU: the host compile of csrc/f16_fast.cuh (tests/hostemu, tests
only) against the reference's own Nlplant (oracle/_ref) on N random in-envelope states plus, for every interior breakpoint of
ALPHA1 / BETA1 / DH1 / DH2, queries at relative offsets 1e-16 .. 1e-5 of a cell width on both sides.  Prints the worst scaled
derivative error (bar: 1e-12).  Usage: python tools/stress_fast_host.py [N]"""
import ctypes
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from _inputs import ALPHA1_BP, BETA1_BP, DH1_BP, DH2_BP, X_TRIM_XCG25, random_envelope_xu  # noqa: E402
from conftest import scaled_err  # noqa: E402
from oracle import REF, get_oracle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
dp = ctypes.POINTER(ctypes.c_double)
E = ctypes.CDLL(os.path.join(REPO, "tests", "hostemu", "libf16_hostemu.so"))
assert E.emu_init(os.path.join(REPO, "f16_mpc_oop_py_b200", "data", "f16_aero_v1.bin").encode(), 0) == 0
E.emu_calc_xdot_fast.argtypes = [dp, dp, dp, ctypes.c_double]
o = get_oracle()
o.open_ref()
r = np.random.default_rng(2026)
xu = random_envelope_xu(n, seed=99, hifi=True)
x = np.vstack([xu, r.uniform(-30, 30, (1, n))])
x[2] = r.uniform(0, 100000, n)
d2r = np.pi / 180
extra = []
for bps, idx, k in ((ALPHA1_BP, 7, d2r), (BETA1_BP, 8, d2r), (sorted(set(DH1_BP + DH2_BP)), 13, 1.0)):
    for j in range(1, len(bps) - 1):
        w = bps[j + 1] - bps[j]
        for e in range(-16, -4):
            for s in (1.0, -1.0):
                for _ in range(3):
                    c = X_TRIM_XCG25.copy()
                    c[7], c[8], c[13] = r.uniform(-19, 44) * d2r, r.uniform(-29, 29) * d2r, r.uniform(-24, 24)
                    c[idx] = (bps[j] + s * w * 10.0 ** e * r.uniform(1, 10)) * k
                    c[6] = r.uniform(200, 900)
                    extra.append(c)
x = np.ascontiguousarray(np.hstack([x, np.array(extra).T]))
n = x.shape[1]
u = np.ascontiguousarray(np.stack([r.uniform(500, 20000, n), r.uniform(-30, 30, n), r.uniform(-25, 25, n), r.uniform(-35, 35, n)]))
for xcg in (0.25, 0.35):
    ref, st = o.calc_xdot_batch(x, u, 1, xcg, REF)
    out = np.full_like(ref, np.nan)
    xd = np.zeros(18)
    for i in range(n):
        if st[i]:
            continue
        s = E.emu_calc_xdot_fast(np.ascontiguousarray(x[:, i]).ctypes.data_as(dp), np.ascontiguousarray(u[:, i]).ctypes.data_as(dp),
                                 xd.ctypes.data_as(dp), xcg)
        assert s == 0, (i, s)
        out[:, i] = xd
    ok = st == 0
    print(f"xcg {xcg}: {int(ok.sum())} of {n} states inside the envelope, worst scaled derivative error fast-vs-reference "
          f"{scaled_err(out[:, ok], ref[:, ok]):.3e}")
