#!/usr/bin/env python3
"""How far do two builds of the REFERENCE'S OWN source disagree on a finite-difference Jacobian?

Runs only where /root/reference exists (this container).  Compiles C/nlplant.c twice -- the reference's build line (no -O
flag, = oracle/_ref) and `-O3 -march=native` (FMA contraction) -- and evaluates env.py::linearise (forward, eps 1e-5) and
the central scheme over a 32 x 32 altitude x airspeed grid with each.  The difference is the noise floor of the A/B parity
bar: the quotient divides last-bit differences of f by eps.  Output: profiles/r02_jacobian_noise_floor.md.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/C/nlplant.c"

CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from oracle import Oracle, REF
from _inputs import X_TRIM_XCG35
o = Oracle(); assert o.open_ref(sys.argv[1])
hh, vv = np.meshgrid(np.linspace(5000, 40000, 32), np.linspace(300, 900, 32), indexing="ij")
n = hh.size
x = np.tile(X_TRIM_XCG35[:, None], (1, n)); x[2] = hh.ravel(); x[6] = vv.ravel()
r = np.random.default_rng(1); x[7] = r.uniform(0.0, 0.2, n); x[4] = x[7]; x[5] = r.uniform(-3, 3, n)
u = np.ascontiguousarray(x[12:16])
xd, _ = o.calc_xdot_batch(np.ascontiguousarray(x), u, 1, 0.35, REF)
out = {"xd": xd}
for sch in (0, 1):
    A, B, st = o.linearise_batch(np.ascontiguousarray(x), u, 1e-5, sch, 1, 0.35, REF)
    out["A%%d" %% sch], out["B%%d" %% sch] = A, B
np.savez(sys.argv[2], **out)
''' % (REPO, os.path.join(REPO, "tests"))


def build(dst, flags):
    os.makedirs(os.path.join(dst, "C"), exist_ok=True)
    for f in os.listdir(os.path.join(REPO, "oracle", "_ref", "C")):
        subprocess.check_call(["cp", os.path.join(REPO, "oracle", "_ref", "C", f), os.path.join(dst, "C", f)])
    zh = os.path.join(REPO, "oracle", "zero_heap.c")
    subprocess.check_call(["/usr/bin/gcc", "-w", *flags, "-fPIC", "-shared", "-o", os.path.join(dst, "nlplant_xcg25.so"), REF, zh,
                           "-Wl,--wrap=malloc", "-lm"])
    src = open(REF).read().replace("double xcg  = 0.25;", "double xcg  = 0.35;")
    subprocess.run(["/usr/bin/gcc", "-w", *flags, "-I/root/reference/C", "-fPIC", "-shared", "-o",
                    os.path.join(dst, "nlplant_xcg35.so"), "-x", "c", "-", "-x", "none", zh, "-Wl,--wrap=malloc", "-lm"],
                   input=src.encode(), check=True)


def main():
    if not os.path.exists(REF):
        sys.exit("needs /root/reference")
    tmp = tempfile.mkdtemp()
    res = {}
    for tag, flags in (("O0", []), ("O2", ["-O2"]), ("O3_native", ["-O3", "-march=native"])):
        d = os.path.join(tmp, tag)
        build(d, flags)
        out = os.path.join(tmp, tag + ".npz")
        subprocess.check_call([sys.executable, "-c", CHILD, d, out])
        res[tag] = np.load(out)
    lines = ["# Noise floor of the finite-difference Jacobians: two builds of the reference's own source",
             "", "`tools/jac_noise_floor.py`: `/root/reference/C/nlplant.c` compiled with its own build line (no `-O`), with `-O2`, and with",
             "`-O3 -march=native`; `env.py::linearise` (forward, eps 1e-5) and the central scheme on a 32 x 32 altitude x airspeed grid",
             "(5 000 .. 40 000 ft, 300 .. 900 ft/s), 331 776 entries of A per scheme. Differences against the `-O0` build:", "",
             "| build | scheme | max abs diff A | entries > 1e-8 | where (row, col) | value there | max abs diff B |", "|---|---|---|---|---|---|---|"]
    for tag in ("O2", "O3_native"):
        for sch, name in ((0, "forward"), (1, "central")):
            d = np.abs(res["O0"]["A%d" % sch] - res[tag]["A%d" % sch])
            i = np.unravel_index(np.argmax(d), d.shape)
            db = np.abs(res["O0"]["B%d" % sch] - res[tag]["B%d" % sch]).max()
            lines.append(f"| {tag} | {name} | {d.max():.3e} | {int((d > 1e-8).sum())} | ({i[1]}, {i[2]}) | "
                         f"{res['O0']['A%d' % sch][i]:.4g} | {db:.3e} |")
    xd = res["O0"]["xd"]
    lines += ["", f"The entries that move are in the navigation rows (npos_dot, epos_dot up to {np.abs(xd[:2]).max():.0f} ft/s: one ulp of f is "
              f"{np.spacing(np.abs(xd[:2]).max()):.2e}, and one ulp / eps = {np.spacing(np.abs(xd[:2]).max()) / 1e-5:.2e}).",
              "A contracted multiply-add in the reference's own arithmetic moves the forward quotient by 2.3e-8: an ABSOLUTE 1e-8 bar on every",
              "entry of A is below what the reference's source pins. The parity tests therefore use",
              "`|dA_ij| <= 1e-8 + 4 ulp(|f_i|) / h` (h = eps forward, 2 eps central): 1e-8 wherever |f_i| < 100, 5.5e-8 on a 900 ft/s navigation row."]
    open(os.path.join(REPO, "profiles", "r02_jacobian_noise_floor.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
