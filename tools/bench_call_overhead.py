#!/usr/bin/env python3
"""Per-call cost of the small-batch paths (run under gpurun): the legacy Nlplant / atmos symbols exactly as env.py calls them (one
aircraft, synchronous), and step_batch_dev / calc_xdot_batch_dev with K = 1 on small resident batches -- what a per-step control
loop pays per call.  Wall clock over many calls."""
import ctypes
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import perturbed_trim, trim_state  # noqa: E402
import f16_mpc_oop_py_b200 as f16  # noqa: E402

L = f16.lib
f16.init()
x_trim, u_trim, _ = trim_state("xcg25")
for math in ("strict", "fast"):
    L.f16_set_math_mode(f16.MATH_FAST if math == "fast" else f16.MATH_STRICT)
    xu = np.ascontiguousarray(x_trim[:17]); xd = np.zeros(18); co = np.zeros(3)
    pxu, pxd, pco = ctypes.c_void_p(xu.ctypes.data), ctypes.c_void_p(xd.ctypes.data), ctypes.c_void_p(co.ctypes.data)
    for name, fn in (("Nlplant (legacy symbol, 1 aircraft)", lambda: L.Nlplant(pxu, pxd, 1)),
                     ("atmos (legacy symbol)", lambda: L.atmos(ctypes.c_double(1e4), ctypes.c_double(700.0), pco))):
        for _ in range(200):
            fn()
        t0 = time.perf_counter()
        for _ in range(5000):
            fn()
        dt = (time.perf_counter() - t0) / 5000
        print(json.dumps({"math": math, "call": name, "us_per_call": 1e6 * dt}), flush=True)
    for n in (1, 256, 4096, 65536):
        x, u = perturbed_trim(n, x_trim, u_trim, seed=1)
        d_x, d_u, d_o, d_st = L.f16_dev_alloc(x.nbytes), L.f16_dev_alloc(u.nbytes), L.f16_dev_alloc(x.nbytes), L.f16_dev_alloc(4 * n)
        L.f16_memcpy_h2d(d_x, x.ctypes.data, x.nbytes); L.f16_memcpy_h2d(d_u, u.ctypes.data, u.nbytes)
        for name, fn in (("step_batch_dev K=1", lambda: L.step_batch_dev(d_x, n, d_u, n, n, 1, 0.001, None, None, 1, None, 0.25, d_st, None)),
                         ("calc_xdot_batch_dev", lambda: L.calc_xdot_batch_dev(d_x, n, d_u, n, d_o, n, None, 1, None, 0.25, n, d_st))):
            for _ in range(100):
                fn()
            L.f16_sync()
            t0 = time.perf_counter()
            for _ in range(2000):
                fn()
            L.f16_sync()
            dt = (time.perf_counter() - t0) / 2000
            print(json.dumps({"math": math, "call": name, "aircraft": n, "us_per_call": 1e6 * dt}), flush=True)
        for p in (d_x, d_u, d_o, d_st):
            L.f16_dev_free(p)
