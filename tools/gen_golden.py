#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (env.py, utils.py, parameters.py and its shipped
C/nlplant_xcg25.so / nlplant_xcg35.so) in this container.  Needs /root/reference; the fixtures it writes are
committed so that the tests never need it.

The reference imports packages that are absent here (gym, ursina, osqp, progressbar, control, matplotlib) and
uses np.infty (removed in NumPy 2): empty stub modules and `np.infty = np.inf` are injected, nothing in the
reference tree is modified (it is mounted read-only).

Fixtures
  env_xcg25.npz / env_xcg35.npz (hifi), env_lofi_xcg25.npz:
    x_trim            F16.trim(10000, 700)                         env.py:198-292
    Ac, Bc            F16.linearise at trim (forward FD, 1e-5)     env.py:294-342
    xdot_trim         F16._calc_xdot(x_trim, u_trim)               env.py:65-103
    xs, us, xdots     64 perturbed states/inputs and their _calc_xdot
    traj_x            state after 0,500,...,2000 F16.step calls from trim, open loop   env.py:105-130
    nl_xu, nl_xdot    parameters.py x0 (known-answer vector of SURVEY 8c) through Nlplant, hifi and lofi
    Ad, Bd            cont2discrete(Ac, Bc, dt) (ZOH)                                   env.py:46
    na_Ac, na_Bc, na_Ad, na_Bd   the reduced 9-state / 3-input model of env.py:49-60 (ssr)
    na_x, na_u, na_xdots         16 reduced states/inputs and their F16._calc_xdot_na   env.py:152-193
    K_lqr             F16._calc_LQR_gain(): -dlqr(Ad, Bd, C'C, I) on the reduced model  env.py:344-358, utils.py:219-245
"""
import ctypes
import os
import sys
import types

import numpy as np

REF = os.environ.get("F16_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class Env:
        def __init__(self, *a, **k):
            pass

    mod("gym", Env=Env, spaces=mod("gym.spaces"))
    mod("ursina")
    mod("osqp")
    mod("progressbar")
    control = mod("control")
    control.matlab = mod("control.matlab", ctrb=None, obsv=None, lqr=None)
    mpl = mod("matplotlib")
    mpl.pyplot = mod("matplotlib.pyplot")
    np.infty = np.inf


def main():
    sys.dont_write_bytecode = True
    install_stubs()
    ctypes.CDLL("libm.so.6", mode=ctypes.RTLD_GLOBAL)  # the shipped .so leave libm unresolved
    os.chdir(REF)
    sys.path.insert(0, REF)
    import parameters as P
    from env import F16

    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(0xF16)
    rng_na = np.random.default_rng(0xF17)   # separate stream: the fixtures above keep their values
    for tag, so, fi in (("xcg25", "nlplant_xcg25.so", 1), ("xcg35", "nlplant_xcg35.so", 1),
                        ("lofi_xcg25", "nlplant_xcg25.so", 0)):
        lib = ctypes.CDLL(os.path.join(REF, "C", so))
        sv = P.stateVector(P.states, np.copy(P.x0), P.x_units, P.x_ub, P.x_lb, np.copy(P.x0), P.observed_states,
                           P.mpc_states, P.mpc_inputs, P.mpc_controlled_states)
        iv = P.inputVector(P.inputs, np.copy(P.u0), P.u_units, P.u_ub, P.u_lb, P.udot_ub, P.udot_lb, np.copy(P.u0),
                           P.mpc_inputs)
        sp = P.simulationParameters(P.dt, P.time_start, P.time_end, 1 if "35" in tag else 0, fi)
        ss = P.stateSpace(*[None] * 8)
        f16 = F16(sv, iv, sp, ss, lib)
        x_trim = np.copy(f16.x.initial_condition)
        u_trim = np.copy(f16.u.initial_condition)
        xdot_trim = f16._calc_xdot(x_trim, u_trim)
        # perturbed evaluations
        xs, us, xdots = [], [], []
        for _ in range(64):
            x = x_trim * (1 + 0.05 * rng.uniform(-1, 1, 18)) + np.where(x_trim == 0, 0.05 * rng.uniform(-1, 1, 18), 0)
            u = u_trim * (1 + 0.2 * rng.uniform(-1, 1, 4))
            xs.append(x)
            us.append(u)
            xdots.append(f16._calc_xdot(x, u))
        # open-loop trajectory through the real F16.step
        f16.reset()
        traj = [np.copy(f16.x.values)]
        for k in range(2000):
            f16.step(f16.u.values)
            if (k + 1) % 500 == 0:
                traj.append(np.copy(f16.x.values))
        # raw Nlplant on the rough-trim x0 of parameters.py:105
        nl_xu = np.copy(P.x0[:17])
        nl = []
        for fid in (1, 0):
            xd = np.zeros(18)
            lib.Nlplant(ctypes.c_void_p(nl_xu.ctypes.data), ctypes.c_void_p(xd.ctypes.data), ctypes.c_int(fid))
            nl.append(xd)
        # reduced (no-actuator) model: what feeds _calc_LQR_gain / _calc_MPC_action
        f16.reset()
        na_x, na_u, na_xdots = [], [], []
        x_mpc0, u_mpc0 = f16.x._get_mpc_x(), f16.u._get_mpc_u()
        for _ in range(16):
            xm = x_mpc0 * (1 + 0.05 * rng_na.uniform(-1, 1, 9)) + np.where(x_mpc0 == 0, 0.05 * rng_na.uniform(-1, 1, 9), 0)
            um = u_mpc0 * (1 + 0.2 * rng_na.uniform(-1, 1, 3))
            na_x.append(xm)
            na_u.append(um)
            na_xdots.append(f16._calc_xdot_na(xm, um))
        K_lqr = f16._calc_LQR_gain()
        np.savez(os.path.join(OUT, f"env_{tag}.npz"), Ad=f16.ss.Ad, Bd=f16.ss.Bd, na_Ac=f16.ssr.Ac, na_Bc=f16.ssr.Bc,
                 na_Ad=f16.ssr.Ad, na_Bd=f16.ssr.Bd, na_x=np.array(na_x), na_u=np.array(na_u), na_xdots=np.array(na_xdots),
                 K_lqr=np.asarray(K_lqr), mpc_u_in_x_idx=np.array(f16.x._mpc_u_in_x_idx), x_trim=x_trim, u_trim=u_trim, Ac=f16.ss.Ac, Bc=f16.ss.Bc,
                 xdot_trim=xdot_trim, xs=np.array(xs), us=np.array(us), xdots=np.array(xdots), traj_x=np.array(traj),
                 nl_xu=nl_xu, nl_xdot=np.array(nl), fi=fi, xcg=0.35 if "35" in tag else 0.25,
                 mpc_x_idx=np.array(f16.x._mpc_x_idx), mpc_u_idx=np.array(f16.u._mpc_u_idx),
                 obs_x_idx=np.array(f16.x._obs_x_idx))
        print(tag, "trim alpha", x_trim[7], "T", x_trim[12], "xdot_trim[6:12]", xdot_trim[6:12])


if __name__ == "__main__":
    main()
