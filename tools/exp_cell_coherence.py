#!/usr/bin/env python3
"""Experiment: how much of the step kernel's time is the table gather's address divergence?  The same 2^20-aircraft launch
(K = 500, fast hifi) with the sideslip of the batch (a) as in cfg 2 (+-0.05 rad: four beta cells, straddling the breakpoint at
0), (b) inside ONE beta cell (0.5 .. 1.5 deg), (c) alpha, beta and elevator all cell-uniform and lateral rates zero.  Run under gpurun."""
import ctypes
import json
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
n, K = 1 << 20, int(sys.argv[1]) if len(sys.argv) > 1 else 500
x_trim, u_trim, _ = trim_state("xcg25")
x, u = perturbed_trim(n, x_trim, u_trim, seed=0xF16)
r = np.random.default_rng(1)


def ck(rc):
    assert rc == 0, L.f16_last_error().decode()


def run(tag, xx):
    d_x0, d_x, d_u, d_st = (L.f16_dev_alloc(xx.nbytes), L.f16_dev_alloc(xx.nbytes), L.f16_dev_alloc(u.nbytes), L.f16_dev_alloc(4 * n))
    ck(L.f16_memcpy_h2d(d_x0, xx.ctypes.data, xx.nbytes)); ck(L.f16_memcpy_h2d(d_u, u.ctypes.data, u.nbytes))
    best = 1e9
    for i in range(4):
        ck(L.f16_memcpy_d2d(d_x, d_x0, xx.nbytes)); ck(L.f16_sync()); ck(L.f16_timer_start())
        ck(L.step_batch_dev(d_x, n, d_u, n, n, K, 0.001, None, None, 1, None, 0.25, d_st, None))
        ms = ctypes.c_float(0); ck(L.f16_timer_stop(ctypes.byref(ms)))
        if i: best = min(best, ms.value)
    st = np.zeros(n, np.int32); ck(L.f16_memcpy_d2h(st.ctypes.data, d_st, 4 * n))
    print(json.dumps({"case": tag, "K": K, "ms": best, "steps_per_s": n * K / best * 1e3, "alive": float((st == 0).mean())}), flush=True)
    for p in (d_x0, d_x, d_u, d_st): L.f16_dev_free(p)


run("cfg2 (+-0.05 rad sideslip)", x)
xb = x.copy(); xb[8] = np.deg2rad(r.uniform(0.5, 1.5, n)); run("beta in one cell", xb)
xc = xb.copy(); xc[9] = 0; xc[11] = 0; xc[3] = 0; run("beta in one cell, p = r = phi = 0", xc)
xd = x.copy(); xd[8] = np.deg2rad(r.uniform(-1.9, 1.9, n)); run("beta in two cells (+-1.9 deg)", xd)
