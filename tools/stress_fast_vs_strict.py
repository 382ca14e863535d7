#!/usr/bin/env python3
"""Stress check (GPU): 2^20 random in-envelope states per fidelity, 25 Euler steps in F16_MATH_STRICT and F16_MATH_FAST --
status words and steps taken must agree everywhere, surviving states to rounding (measured: 100 %, 2.5e-15 scaled)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import f16_mpc_oop_py_b200 as f16
from _inputs import random_envelope_xu
f16.init()
n = 1 << 20
for fi in (0, 1):
    xu = random_envelope_xu(n, seed=100 + fi, hifi=bool(fi))
    r = np.random.default_rng(7 + fi)
    x = np.vstack([xu, r.uniform(-20, 5, (1, n))]); x[2] = r.uniform(500, 39000, n)
    u = np.vstack([r.uniform(1000, 19000, n), r.uniform(-25, 25, n), r.uniform(-21.5, 21.5, n), r.uniform(-30, 30, n)])
    out = {}
    for mode, name in ((f16.MATH_STRICT, "strict"), (f16.MATH_FAST, "fast")):
        f16.lib.f16_set_math_mode(mode)
        b = f16.F16Batch(x, u, fi_flag=fi, xcg=0.3)
        b.step(K=25)
        out[name] = (b.x.copy(), b.status.copy(), b.steps_done.copy())
    xs, ss, ks = out["strict"]; xf, sf, kf = out["fast"]
    same = (ss == sf) & (ks == kf)
    alive = same & (ss == 0)
    scale = np.maximum(np.abs(xs), np.sqrt(np.mean(xs[:, alive] ** 2, axis=1, keepdims=True)) + 1e-300)
    err = np.abs(xf - xs) / scale
    print("fi", fi, "status/steps agree", same.mean(), "alive", alive.mean(), "max scaled err over alive", err[:, alive].max(), "argmax state", np.unravel_index(err[:, alive].argmax(), err[:, alive].shape)[0])
    d = ~same
    if d.any():
        i = np.flatnonzero(d)[:5]
        print("  first disagreements: strict", ss[i], ks[i], "fast", sf[i], kf[i])
