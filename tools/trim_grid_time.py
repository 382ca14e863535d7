#!/usr/bin/env python3
"""Time trim_batch on the BASELINE cfg-4 grid (64 x 64 altitude x airspeed, env.py::trim settings) in both math modes."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import f16_mpc_oop_py_b200 as f16  # noqa: E402

f16.init()
hh, vv = np.meshgrid(np.linspace(5000, 40000, 64), np.linspace(300, 900, 64), indexing="ij")
for mode, name in ((f16.MATH_STRICT, "strict"), (f16.MATH_FAST, "fast")):
    f16.lib.f16_set_math_mode(mode)
    for fi in (1, 0):
        for xcg in (0.25, 0.35):
            f16.trim(hh.ravel(), vv.ravel(), fi=fi, xcg=xcg)
            t0 = time.perf_counter()
            x, opt = f16.trim(hh.ravel(), vv.ravel(), fi=fi, xcg=xcg)
            dt = time.perf_counter() - t0
            it = opt["nit"]
            print(f"{name:6s} fi {fi} xcg {xcg}: {1e3 * dt:8.1f} ms  converged {np.mean(opt['success']):.3f}  iterations median "
                  f"{np.median(it):.0f} / p99 {np.percentile(it, 99):.0f} / max {it.max():.0f}  evaluations {opt['nfev'].sum():.3g}")
