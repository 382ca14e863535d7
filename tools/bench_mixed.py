#!/usr/bin/env python3
"""Throughput of a mixed-fidelity batch with interleaved flags (BASELINE cfg 3 as a step workload): step_batch_dev with a
per-aircraft fi array (partitioned by fidelity inside the call) against the same aircraft run as two uniform batches."""
import ctypes
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import perturbed_trim, trim_state  # noqa: E402
import f16_mpc_oop_py_b200 as f16  # noqa: E402

L = f16.lib
f16.init()
L.f16_set_math_mode(f16.MATH_FAST)
n, K = 1 << 20, 200
x_trim, u_trim, _ = trim_state("xcg25")
x, u = perturbed_trim(n, x_trim, u_trim, seed=3, frac=0.03)
fi = (np.arange(n) % 2).astype(np.uint8)
d_x0, d_x, d_u, d_st, d_fi = (L.f16_dev_alloc(x.nbytes), L.f16_dev_alloc(x.nbytes), L.f16_dev_alloc(u.nbytes), L.f16_dev_alloc(4 * n),
                              L.f16_dev_alloc(n))
L.f16_memcpy_h2d(d_x0, x.ctypes.data, x.nbytes)
L.f16_memcpy_h2d(d_u, u.ctypes.data, u.nbytes)
L.f16_memcpy_h2d(d_fi, fi.ctypes.data, n)


def timed(fn, reps=3):
    best = None
    for _ in range(reps + 1):
        L.f16_memcpy_d2d(d_x, d_x0, x.nbytes)
        L.f16_timer_start()
        fn()
        ms = ctypes.c_float(0)
        L.f16_timer_stop(ctypes.byref(ms))
        best = ms.value if best is None else min(best, ms.value)
    return best


t_mixed = timed(lambda: L.step_batch_dev(d_x, n, d_u, n, n, K, 0.001, None, d_fi, 1, None, 0.25, d_st, None))
h = n // 2
t_split = timed(lambda: (L.step_batch_dev(d_x, n, d_u, n, h, K, 0.001, None, None, 1, None, 0.25, d_st, None),
                         L.step_batch_dev(d_x + h * 8, n, d_u + h * 8, n, h, K, 0.001, None, None, 0, None, 0.25, d_st + h * 4, None)))
print(f"mixed interleaved batch, {n} aircraft x {K} steps: {t_mixed:.3f} ms = {n * K / t_mixed / 1e-3:.3e} aircraft-steps/s; "
      f"the same as two uniform batches: {t_split:.3f} ms = {n * K / t_split / 1e-3:.3e}")
