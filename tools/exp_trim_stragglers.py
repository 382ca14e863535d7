import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import f16_mpc_oop_py_b200 as f16
f16.init()
f16.lib.f16_set_math_mode(f16.MATH_FAST)
hh, vv = np.meshgrid(np.linspace(5000, 40000, 64), np.linspace(300, 900, 64), indexing="ij")
h, v = hh.ravel(), vv.ravel()
res = {}
for mi in (3000, 6000, 12000):
    t0 = time.perf_counter(); x, opt = f16.trim(h, v, fi=1, xcg=0.25, maxiter=mi); dt = time.perf_counter() - t0
    res[mi] = (x, opt)
    bad = ~opt["success"]
    print(mi, f"{dt*1e3:.1f} ms", "unconverged", bad.sum(), "cost of unconverged: min %.3g median %.3g max %.3g" % (opt["fun"][bad].min(), np.median(opt["fun"][bad]), opt["fun"][bad].max()),
          "nfev/nit", (opt["nfev"][bad] / opt["nit"][bad]).mean(), "status nonzero", (opt["status"][bad] != 0).sum())
xa, oa = res[6000]; xb, ob = res[12000]
bad = ~ob["success"]
same = np.array([np.array_equal(xa[:, i], xb[:, i]) for i in np.flatnonzero(bad)])
print("unconverged at 12000 whose x is bit-identical at 6000:", same.sum(), "of", bad.sum())
d = np.abs(xa[:, bad] - xb[:, bad]).max(axis=0)
print("max |x(6000) - x(12000)| over unconverged: median %.3g max %.3g" % (np.median(d), d.max()))
print("cost change", np.abs(oa["fun"][bad] - ob["fun"][bad]).max())
i = np.flatnonzero(bad)[:5]
print("h", h[i], "V", v[i], "cost", ob["fun"][i], "alpha", xb[7, i])
