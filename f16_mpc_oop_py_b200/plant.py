"""Host-side mirror of the reference's plant interface (env.py::F16) over the batched C ABI.

`F16Batch` keeps the reference's method names and argument meaning -- `_calc_xdot(x, u)`, `step(action)`,
`reset()`, `get_obs(x, u)`, `linearise(x, u)` -- for N aircraft at once, with states as SoA arrays [18][N].
Every method is a call into libf16_b200.so (CUDA); nothing is computed in Python.

Differences from env.py that a user sees:
  * leaving the envelope does not `exit()` the process (env.py:117-124): the aircraft is flagged in `status`
    and frozen at the violating state;
  * `linearise` offers scheme='forward' (env.py:319-340) and 'central';
  * xcg and fidelity are arguments (scalars or per-aircraft arrays) instead of a choice of .so file
    (parameters.py:108-114) and a dataclass field.
"""
import ctypes

import numpy as np

from . import _lib
from . import parameters as P
from ._lib import LqrLaw, check, lib


def _c(a, dtype=np.float64):
    return np.ascontiguousarray(a, dtype=dtype)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _sel(fi, xcg, n):
    """(fi array or None, default fi, xcg array or None, default xcg) for the C ABI."""
    fi_arr = xcg_arr = None
    fi_def, xcg_def = 1, 0.25
    if np.ndim(fi) == 0:
        fi_def = int(fi)
    else:
        fi_arr = _c(fi, np.uint8)
        assert fi_arr.shape == (n,)
    if np.ndim(xcg) == 0:
        xcg_def = float(xcg)
    else:
        xcg_arr = _c(xcg)
        assert xcg_arr.shape == (n,)
    return fi_arr, fi_def, xcg_arr, xcg_def


def make_lqr(K, sel, x_ref, u0, rows):
    """f16_lqr_t for u[rows] = u0[rows] - K (x[sel] - x_ref); K is [len(rows)][len(sel)]."""
    law = LqrLaw()
    K = np.atleast_2d(np.asarray(K, dtype=np.float64))
    assert K.shape == (len(rows), len(sel)) and len(sel) <= 18
    law.n_sel = len(sel)
    law.row_mask = 0
    for j, s in enumerate(sel):
        law.sel[j] = int(s)
        law.x_ref[j] = float(x_ref[j])
    for i, r in enumerate(rows):
        law.row_mask |= 1 << int(r)
        for j in range(len(sel)):
            law.K[r][j] = float(K[i, j])
    for r in range(4):
        law.u0[r] = float(u0[r])
    return law


def nlplant(xu, fi=1, xcg=0.25):
    """Nlplant_batch: xu [17][N] -> (xdot [18][N], status [N])  (C/nlplant.c:23-457 for N aircraft)."""
    xu = _c(xu)
    assert xu.ndim == 2 and xu.shape[0] == 17
    n = xu.shape[1]
    fa, fd, xa, xd = _sel(fi, xcg, n)
    out = np.empty((18, n))
    st = np.zeros(n, dtype=np.int32)
    check(lib.Nlplant_batch(_p(xu), _p(out), None if fa is None else fa.ctypes.data_as(_lib.c_ubp), fd,
                            None if xa is None else xa.ctypes.data_as(_lib.c_dp), xd, n, _p(st)), "Nlplant_batch")
    return out, st


def atmos(alt, vt):
    """atmos_batch: -> [3][N] = mach, qbar, ps (C/nlplant.c:467-490)."""
    alt, vt = _c(np.atleast_1d(alt)), _c(np.atleast_1d(vt))
    out = np.empty((3, alt.size))
    check(lib.atmos_batch(_p(alt), _p(vt), alt.size, _p(out)), "atmos_batch")
    return out


def trim(h_t, v_t, fi=1, xcg=0.25, tol=1e-10, maxiter=50000, ux0=None):
    """trim_batch: env.py::trim(h_t, v_t) for arrays of altitudes [ft] and airspeeds [ft/s] (env.py:198-292).
    -> (x_trim [18][N], opt) with opt = dict(fun, nit, nfev, success, status), the fields of scipy's OptimizeResult the
    reference returns as `opt`."""
    h, v = _c(np.atleast_1d(h_t)), _c(np.atleast_1d(v_t))
    assert h.shape == v.shape and h.ndim == 1
    n = h.size
    fa, fd, xa, xd = _sel(fi, xcg, n)
    x = np.empty((18, n))
    info = np.empty((4, n))
    st = np.zeros(n, dtype=np.int32)
    g = None if ux0 is None else _c(ux0)
    check(lib.trim_batch(_p(h), _p(v), n, float(tol), int(maxiter), _p(g), _p(x), _p(info),
                         None if fa is None else fa.ctypes.data_as(_lib.c_ubp), fd,
                         None if xa is None else xa.ctypes.data_as(_lib.c_dp), xd, _p(st)), "trim_batch")
    opt = {"fun": info[0], "nit": info[1].astype(np.int64), "nfev": info[2].astype(np.int64),
           "success": info[3] != 0, "status": st}
    return x, opt


def state_summary(x, status=None):
    """state_summary_batch: the end-of-run statistics of a batch, reduced on the device.  x [18][N], status [N] int32 or
    None -> the 74-double row [N, alive, min[18], max[18], mean[18], M2[18]] over the aircraft with status 0 (the layout of
    shard.summarise / shard.merge_summaries: what a rank contributes to the one all-gather of a multi-GPU run)."""
    x = _c(x)
    assert x.ndim == 2 and x.shape[0] == 18
    st = None if status is None else _c(status, np.int32)
    row = np.empty(74)
    check(lib.state_summary_batch(_p(x), x.shape[1], _p(st), _p(row)), "state_summary_batch")
    return row


def state_summary_dev(d_x, ld, n, d_status=None):
    """state_summary_batch_dev on device pointers (plane stride ld): no copy of the state to the host."""
    row = np.empty(74)
    check(lib.state_summary_batch_dev(d_x, int(ld), int(n), d_status, _p(row)), "state_summary_batch_dev")
    return row


def reduce_jacobian(A):
    """A [N][18][18] -> (A_na [N][9][9], B_na [N][9][3]): the reduced model of env.py:49,152-193 (an exact gather)."""
    A = _c(A).reshape(-1, 18, 18)
    n = A.shape[0]
    Ana, Bna = np.empty((n, 9, 9)), np.empty((n, 9, 3))
    check(lib.reduce_jacobian_batch(_p(A), n, _p(Ana), _p(Bna)), "reduce_jacobian_batch")
    return Ana, Bna


def discretise(A, B, dt=P.dt):
    """scipy.signal.cont2discrete((A, B, ., .), dt)[0:2] (zero-order hold, env.py:46,50) for stacks [N][n][n], [N][n][m]."""
    A, B = _c(A), _c(B)
    A = A.reshape((-1,) + A.shape[-2:])
    B = B.reshape((-1,) + B.shape[-2:])
    n, m = A.shape[1], B.shape[2]
    Ad, Bd = np.empty_like(A), np.empty_like(B)
    check(lib.discretise_batch(_p(A), _p(B), n, m, A.shape[0], float(dt), _p(Ad), _p(Bd)), "discretise_batch")
    return Ad, Bd


def dlqr(A, B, Q, R):
    """utils.py:219-245 for stacks of discrete systems: K [N][m][n] (and P) with u = -K x; Q, R shared."""
    A, B = _c(A), _c(B)
    A = A.reshape((-1,) + A.shape[-2:])
    B = B.reshape((-1,) + B.shape[-2:])
    n, m = A.shape[1], B.shape[2]
    Q, R = _c(Q).reshape(n, n), _c(R).reshape(m, m)
    K, Pm = np.empty((A.shape[0], m, n)), np.empty_like(A)
    info = np.zeros((A.shape[0], 2), dtype=np.int32)
    check(lib.dlqr_batch(_p(A), _p(B), _p(Q), _p(R), n, m, A.shape[0], _p(K), _p(Pm), _p(info)), "dlqr_batch")
    return K, Pm, info


class F16Batch:
    """N independent F-16s behind the reference's F16 interface (env.py:29-342)."""

    def __init__(self, x0, u0, fi_flag=P.fi_flag, xcg=0.25, dt=P.dt):
        self.initial_x = _c(x0).reshape(18, -1).copy()
        self.initial_u = _c(u0).reshape(4, -1).copy()
        self.n = self.initial_x.shape[1]
        assert self.initial_u.shape[1] == self.n
        self.fi_flag, self.xcg, self.dt = fi_flag, xcg, dt
        self._selargs = _sel(fi_flag, xcg, self.n)
        self.reset()

    def _sel_c(self):
        fa, fd, xa, xd = self._selargs
        return (None if fa is None else fa.ctypes.data_as(_lib.c_ubp), fd,
                None if xa is None else xa.ctypes.data_as(_lib.c_dp), xd)

    # env.py:132-135
    def reset(self):
        self.x = self.initial_x.copy()
        self.u = self.initial_u.copy()
        self.status = np.zeros(self.n, dtype=np.int32)
        self.steps_done = np.zeros(self.n, dtype=np.int32)
        return self.get_obs(self.x, self.u)

    # env.py:137-150
    def get_obs(self, x, u):
        return np.asarray(x)[P.obs_x_idx]

    # env.py:65-103
    def _calc_xdot(self, x, u):
        x, u = _c(x).reshape(18, -1), _c(u).reshape(4, -1)
        n = x.shape[1]
        out = np.empty((18, n))
        st = np.zeros(n, dtype=np.int32)
        check(lib.calc_xdot_batch(_p(x), _p(u), _p(out), *self._sel_c(), n, _p(st)), "calc_xdot_batch")
        self.last_status = st
        return out

    # env.py:105-130: `action` = [T, dh, da, dr] demands, [4] or [4][N]; K fused Euler steps per call
    def step(self, action=None, K=1, lqr=None):
        if action is not None:
            a = _c(action)
            self.u = np.ascontiguousarray(np.broadcast_to(a.reshape(4, -1), (4, self.n)))
        law = ctypes.byref(lqr) if lqr is not None else None
        check(lib.step_batch(_p(self.x), _p(self.u), self.n, int(K), float(self.dt), law, *self._sel_c(),
                             _p(self.status), _p(self.steps_done)), "step_batch")
        reward, isdone = 1, self.status != 0
        info = {'fidelity': 'high' if np.all(np.asarray(self.fi_flag) == 1) else 'mixed/low', 'status': self.status}
        return self.get_obs(self.x, self.u), reward, isdone, info

    # env.py:198-292
    def trim(self, h_t, v_t, **kw):
        return trim(h_t, v_t, fi=self.fi_flag if np.ndim(self.fi_flag) == 0 else 1,
                    xcg=self.xcg if np.ndim(self.xcg) == 0 else 0.25, **kw)

    # the driver loops of the reference (test_env.py:452-462, test_env_mk2.py:70-85): K steps, state stored every snap_every
    def summary(self):
        """statistics of the current states over the aircraft still flying (status 0), reduced on the device"""
        return state_summary(self.x, self.status)

    def rollout_stats(self, K, snap_every, lqr=None):
        """Monte-Carlo rollout that keeps statistics instead of trajectories: -> rows [K // snap_every][74] (state_summary
        layout) of the batch after every snap_every steps; self.x, self.status end as after step(K=K)"""
        rows = np.empty((int(K) // int(snap_every), 74))
        law = ctypes.byref(lqr) if lqr is not None else None
        check(lib.step_batch_stats(_p(self.x), _p(self.u), self.n, int(K), int(snap_every), float(self.dt), law, *self._sel_c(),
                                   _p(rows), _p(self.status)), "step_batch_stats")
        return rows

    def rollout(self, K, snap_every, lqr=None):
        """-> traj [K // snap_every][18][N]; self.x, self.status end as after step(K=K)"""
        ns = int(K) // int(snap_every)
        traj = np.empty((ns, 18, self.n))
        law = ctypes.byref(lqr) if lqr is not None else None
        check(lib.step_batch_traj(_p(self.x), _p(self.u), self.n, int(K), int(snap_every), float(self.dt), law, *self._sel_c(),
                                  _p(traj), _p(self.status)), "step_batch_traj")
        return traj

    # env.py:344-358
    def _calc_LQR_gain(self, x=None, u=None):
        """K [N][3][9] = -dlqr(Ad, Bd, C'C, I) of the reduced model at (x, u) (default: the current state)."""
        x = self.x if x is None else _c(x).reshape(18, -1)
        u = self.u if u is None else _c(u).reshape(4, -1)
        n = x.shape[1]
        K = np.empty((n, 3, 9))
        st = np.zeros(n, dtype=np.int32)
        check(lib.lqr_gain_batch(_p(_c(x)), _p(_c(u)), n, float(self.dt), _p(K), *self._sel_c(), _p(st)), "lqr_gain_batch")
        self.last_status = st
        return K

    # env.py:294-342
    def linearise(self, x, u, scheme='forward', eps=1e-5):
        x, u = _c(x).reshape(18, -1), _c(u).reshape(4, -1)
        n = x.shape[1]
        A = np.empty((n, 18, 18))
        B = np.empty((n, 18, 4))
        st = np.zeros(n, dtype=np.int32)
        sch = {'forward': _lib.FD_FORWARD, 'central': _lib.FD_CENTRAL}[scheme]
        check(lib.linearise_batch(_p(x), _p(u), n, float(eps), sch, _p(A), _p(B), *self._sel_c(), _p(st)),
              "linearise_batch")
        self.last_status = st
        # C, D of env.py:311-340 are selector matrices of the observed states
        C = np.zeros((len(P.obs_x_idx), 18))
        C[np.arange(len(P.obs_x_idx)), P.obs_x_idx] = 1.0
        D = np.zeros((len(P.obs_x_idx), 4))
        return A, B, C, D
