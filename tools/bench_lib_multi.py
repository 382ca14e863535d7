#!/usr/bin/env python3
"""One process, every GPU of the box through the C ABI alone (f16_init_devices): e2e aircraft-steps/s of step_batch on host
arrays -- what a plain-C caller of libf16_b200.so gets without torchrun -- and the copy-bound calls (K = 1 step, one derivative)
with the chunk pipeline on and off.  Usage (under gpurun [--gpus N]): python tools/bench_lib_multi.py [--aircraft-per-gpu 1048576]"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import perturbed_trim, trim_state  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--aircraft-per-gpu", type=int, default=1 << 20)
ap.add_argument("--euler-steps", type=int, default=10000)
ap.add_argument("--devices", type=str, default="", help="comma-separated CUDA ordinals (default: all visible)")
ap.add_argument("--samples", type=int, default=3)
args = ap.parse_args()

import f16_mpc_oop_py_b200 as f16  # noqa: E402
L = f16.lib
nd = f16.init_devices([int(d) for d in args.devices.split(",")] if args.devices else None)
L.f16_set_math_mode(f16.MATH_FAST)
n = args.aircraft_per_gpu * nd
x_trim, u_trim, _ = trim_state("xcg25")
x, u = perturbed_trim(n, x_trim, u_trim, seed=0xF16)


def ck(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: {L.f16_last_error().decode()}")


def pinned(shape, dtype):
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = L.f16_host_alloc_pinned(nbytes)
    return p, np.frombuffer((ctypes.c_char * nbytes).from_address(p), dtype=dtype).reshape(shape)


hx, px = pinned((18, n), np.float64)
hu, pu = pinned((4, n), np.float64)
ho, po = pinned((18, n), np.float64)
hs, ps = pinned((n,), np.int32)


def timed(fn, reps, reset=None):
    fn()   # grows the library's scratch
    best = []
    for _ in range(reps):
        if reset:
            reset()
        else:
            px[:] = x
            pu[:] = u
        t0 = time.perf_counter()
        fn()
        best.append(time.perf_counter() - t0)
    return best


def line(**kw):
    print(json.dumps(kw), flush=True)


for K in (args.euler_steps, 1):
    for pipe in (1, 0):
        L.f16_set_host_pipeline(pipe)
        t = timed(lambda: ck(L.step_batch(hx, hu, n, K, 0.001, None, None, 1, None, 0.25, hs, None), "step_batch"), args.samples)
        line(entry="step_batch", devices=nd, aircraft=n, K=K, pipeline=pipe, host="pinned", wall_ms=[1e3 * v for v in t],
             aircraft_steps_per_s=n * K / min(t), alive=float((ps == 0).mean()),
             pcie_gbs=(x.nbytes * 2 + u.nbytes + 4 * n) / min(t) / 1e9 if K == 1 else None)
for pipe in (1, 0):
    L.f16_set_host_pipeline(pipe)
    t = timed(lambda: ck(L.calc_xdot_batch(hx, hu, ho, None, 1, None, 0.25, n, hs), "calc_xdot_batch"), args.samples)
    line(entry="calc_xdot_batch", devices=nd, aircraft=n, pipeline=pipe, host="pinned", wall_ms=[1e3 * v for v in t],
         evals_per_s=n / min(t), pcie_gbs=(x.nbytes * 2 + u.nbytes + 4 * n) / min(t) / 1e9)
L.f16_set_host_pipeline(1)
gx, gu, gs = x.copy(), u.copy(), np.zeros(n, dtype=np.int32)


def pageable():
    ck(L.step_batch(gx.ctypes.data, gu.ctypes.data, n, args.euler_steps, 0.001, None, None, 1, None, 0.25, gs.ctypes.data, None), "step_batch")


def reset_pageable():
    gx[:] = x


t = timed(pageable, args.samples, reset_pageable)
line(entry="step_batch", devices=nd, aircraft=n, K=args.euler_steps, pipeline=1, host="pageable", wall_ms=[1e3 * v for v in t],
     aircraft_steps_per_s=n * args.euler_steps / min(t))
