"""Child process of tests/test_reference_on_dropin.py: imports the UNMODIFIED reference (baseline/_ref/env.py, utils.py,
parameters.py) with the stub modules of tools/gen_golden.py, from a working directory whose C/ is the drop-in directory
(f16_mpc_oop_py_b200/dropin/C), so that parameters.py:108-114 CDLL-loads OUR nlplant_xcg25.so and every Nlplant / atmos
call of env.py:100,187 / utils.py:291 runs on the B200.  Writes what it computed to an .npz for the parent to compare
with tests/golden.

    python _run_reference_on_dropin.py <repo> <workdir> <tag> <out.npz>
"""
import ctypes
import os
import sys

import numpy as np

repo, workdir, tag, out = sys.argv[1:5]
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(repo, "tools"))
from gen_golden import install_stubs  # noqa: E402

install_stubs()
os.chdir(workdir)                                  # parameters.py resolves os.getcwd() + "/C/nlplant_xcg25.so"
sys.path.insert(0, os.path.join(repo, "baseline", "_ref"))
import parameters as P  # noqa: E402
from env import F16  # noqa: E402

golden = np.load(os.path.join(repo, "tests", "golden", f"env_{tag}.npz"))
fi = int(golden["fi"])
# parameters.py picked xcg25 at import (stab_flag = 0); the library handle is a constructor argument (env.py:31,39)
lib = P.nlplant if "35" not in tag else ctypes.CDLL(os.path.join(workdir, "C", "nlplant_xcg35.so"))
loaded = [l.split()[-1] for l in open("/proc/self/maps") if "nlplant_xcg" in l or "libf16_b200" in l]
sv = P.stateVector(P.states, np.copy(P.x0), P.x_units, P.x_ub, P.x_lb, np.copy(P.x0), P.observed_states, P.mpc_states,
                   P.mpc_inputs, P.mpc_controlled_states)
iv = P.inputVector(P.inputs, np.copy(P.u0), P.u_units, P.u_ub, P.u_lb, P.udot_ub, P.udot_lb, np.copy(P.u0), P.mpc_inputs)
sp = P.simulationParameters(P.dt, P.time_start, P.time_end, 1 if "35" in tag else 0, fi)
ss = P.stateSpace(*[None] * 8)
f16 = F16(sv, iv, sp, ss, lib)                     # trim(10000, 700) + two linearisations through the drop-in
x_trim = np.copy(f16.x.initial_condition)
u_trim = np.copy(f16.u.initial_condition)
# from here on at the GOLDEN trim point, so that every comparison is of the same function at the same argument
gx, gu = golden["x_trim"].copy(), golden["u_trim"].copy()
xdot_trim = f16._calc_xdot(gx, gu)
Ac, Bc, _, _ = f16.linearise(gx, gu)
xdots = np.array([f16._calc_xdot(x, u) for x, u in zip(golden["xs"], golden["us"])])
f16.x.initial_condition, f16.u.initial_condition = gx.copy(), gu.copy()
f16.reset()
traj = [np.copy(f16.x.values)]
for k in range(2000):
    f16.step(f16.u.values)
    if (k + 1) % 500 == 0:
        traj.append(np.copy(f16.x.values))
f16.reset()
K_lqr = np.asarray(f16._calc_LQR_gain())
na_xdots = np.array([f16._calc_xdot_na(x, u) for x, u in zip(golden["na_x"], golden["na_u"])])
np.savez(out, x_trim=x_trim, u_trim=u_trim, xdot_trim=xdot_trim, Ac=Ac, Bc=Bc, xdots=xdots, traj_x=np.array(traj),
         K_lqr=K_lqr, na_xdots=na_xdots, loaded=np.array(loaded))
