"""ctypes front-end of the parity oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (f16_mpc_oop_py_b200) never does.

Two backends behind the same calls (`backend=` argument):
  PORT (0)  oracle/f16_oracle.c, the C restatement of the reference algorithm;
  REF  (1)  the reference's own Nlplant/atmos from oracle/_ref/nlplant_xcg{25,35}.so (compiled by
            oracle/Makefile from /root/reference/C/nlplant.c), wrapped in the restated env.py logic.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
BLOB = os.path.join(REPO, "f16_mpc_oop_py_b200", "data", "f16_aero_v1.bin")
LIB_PATH = os.path.join(HERE, "libf16_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

PORT, REF = 0, 1

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int)

# names of the 44 outputs of orc_hifi, in order (f16_oracle.c enum H_*)
HIFI_NAMES = [
    "Cx", "Cz", "Cm", "Cy", "Cn", "Cl",
    "Cxq", "Cyr", "Cyp", "Czq", "Clr", "Clp", "Cmq", "Cnr", "Cnp",
    "dCx_lef", "dCz_lef", "dCm_lef", "dCy_lef", "dCn_lef", "dCl_lef",
    "dCxq_lef", "dCyr_lef", "dCyp_lef", "dCzq_lef", "dClr_lef", "dClp_lef", "dCmq_lef", "dCnr_lef", "dCnp_lef",
    "dCy_r30", "dCn_r30", "dCl_r30",
    "dCy_a20", "dCy_a20_lef", "dCn_a20", "dCn_a20_lef", "dCl_a20", "dCl_a20_lef",
    "dCnbeta", "dClbeta", "dCm", "eta_el", "dCm_ds",
]

# canonical table ids (f16_oracle.c enum T_*) -> (reference accessor symbol, number of arguments)
TABLES = [
    ("_Cx", 3), ("_Cz", 3), ("_Cm", 3), ("_Cn", 3), ("_Cl", 3),
    ("_Cy", 2), ("_Cy_r30", 2), ("_Cn_r30", 2), ("_Cl_r30", 2), ("_Cy_a20", 2), ("_Cn_a20", 2), ("_Cl_a20", 2),
    ("_Cx_lef", 2), ("_Cz_lef", 2), ("_Cm_lef", 2), ("_Cy_lef", 2), ("_Cn_lef", 2), ("_Cl_lef", 2),
    ("_Cy_a20_lef", 2), ("_Cn_a20_lef", 2), ("_Cl_a20_lef", 2),
    ("_CXq", 1), ("_CZq", 1), ("_CMq", 1), ("_CYp", 1), ("_CYr", 1), ("_CNr", 1), ("_CNp", 1), ("_CLp", 1),
    ("_CLr", 1), ("_delta_CNbeta", 1), ("_delta_CLbeta", 1), ("_delta_Cm", 1),
    ("_delta_CXq_lef", 1), ("_delta_CYr_lef", 1), ("_delta_CYp_lef", 1), ("_delta_CZq_lef", 1),
    ("_delta_CLr_lef", 1), ("_delta_CLp_lef", 1), ("_delta_CMq_lef", 1), ("_delta_CNr_lef", 1),
    ("_delta_CNp_lef", 1),
    ("_eta_el", 1),
]
AXES = {"ALPHA1": 0, "ALPHA2": 1, "BETA1": 2, "DH1": 3, "DH2": 4}


class LqrLaw(ctypes.Structure):
    """Mirror of f16_lqr_t (include/f16_b200.h) / orc_lqr_t."""
    _fields_ = [
        ("n_sel", ctypes.c_int),
        ("row_mask", ctypes.c_int),
        ("sel", ctypes.c_int * 18),
        ("K", (ctypes.c_double * 18) * 4),
        ("x_ref", ctypes.c_double * 18),
        ("u0", ctypes.c_double * 4),
    ]


def make_lqr(K, sel, x_ref, u0, rows):
    """K: [len(rows)][len(sel)] gain; rows: input rows (0..3) it drives; x_ref: reference for the selected states."""
    law = LqrLaw()
    K = np.atleast_2d(np.asarray(K, dtype=np.float64))
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


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _soa(a, rows):
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.ndim == 2 and a.shape[0] == rows, a.shape
    return a


class Oracle:
    def __init__(self, lib_path=LIB_PATH, blob=BLOB):
        if not os.path.exists(lib_path):
            raise RuntimeError(f"{lib_path} missing: run `make -C oracle port` (or __graft_entry__.build())")
        self.lib = L = ctypes.CDLL(lib_path)
        L.orc_init.argtypes = [ctypes.c_char_p]
        L.orc_cell.argtypes = [ctypes.c_int, ctypes.c_double, c_ip, c_ip]
        L.orc_interp.argtypes = [ctypes.c_int] + [ctypes.c_double] * 3
        L.orc_interp.restype = ctypes.c_double
        L.orc_hifi.argtypes = [ctypes.c_double] * 3 + [c_dp]
        L.orc_atmos.argtypes = [ctypes.c_double, ctypes.c_double, c_dp]
        L.orc_lofi_damping.argtypes = [ctypes.c_double, c_dp]
        L.orc_lofi_dmomdcon.argtypes = [ctypes.c_double, ctypes.c_double, c_dp]
        L.orc_lofi_clcn.argtypes = [ctypes.c_double, ctypes.c_double, c_dp]
        L.orc_lofi_cxcm.argtypes = [ctypes.c_double, ctypes.c_double, c_dp]
        L.orc_lofi_cz.argtypes = [ctypes.c_double] * 3
        L.orc_lofi_cz.restype = ctypes.c_double
        L.orc_ref_open.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_double]
        L.orc_ref_symbol.argtypes = [ctypes.c_double, ctypes.c_char_p]
        L.orc_ref_symbol.restype = ctypes.c_void_p
        L.orc_nlplant_batch.argtypes = [ctypes.c_int, c_dp, c_dp, ctypes.c_longlong, ctypes.c_int, ctypes.c_double, c_ip]
        L.orc_calc_xdot_batch.argtypes = [ctypes.c_int, c_dp, c_dp, c_dp, ctypes.c_longlong, ctypes.c_int,
                                          ctypes.c_double, c_ip]
        L.orc_step_batch.argtypes = [ctypes.c_int, c_dp, c_dp, ctypes.c_longlong, ctypes.c_int, ctypes.c_double,
                                     ctypes.c_int, ctypes.c_double, ctypes.POINTER(LqrLaw), c_ip]
        L.orc_linearise_batch.argtypes = [ctypes.c_int, c_dp, c_dp, ctypes.c_longlong, ctypes.c_double, ctypes.c_int,
                                          c_dp, c_dp, ctypes.c_int, ctypes.c_double, c_ip]
        L.orc_trim.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_double, c_dp, ctypes.c_double,
                               ctypes.c_int, c_dp, c_dp]
        L.orc_trim_cost.argtypes = [ctypes.c_int, c_dp, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_double, c_dp]
        L.orc_trim_batch.argtypes = [ctypes.c_int, c_dp, c_dp, ctypes.c_longlong, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_int, c_dp, c_dp, c_ip]
        L.orc_max_threads.restype = ctypes.c_int
        rc = L.orc_init(blob.encode())
        if rc != 0:
            raise RuntimeError(f"orc_init({blob}) failed: {rc}")
        self.have_ref = False

    # -- reference backend -------------------------------------------------------------------
    def open_ref(self, ref_dir=REF_DIR):
        """dlopen the reference's own shared objects (oracle/_ref). Returns False if they are absent."""
        if self.have_ref:
            return True
        for xcg, name in ((0.25, "nlplant_xcg25.so"), (0.35, "nlplant_xcg35.so")):
            so = os.path.join(ref_dir, name)
            if not os.path.exists(so) or not os.path.isdir(os.path.join(ref_dir, "C")):
                return False
            rc = self.lib.orc_ref_open(so.encode(), ref_dir.encode(), xcg)
            if rc != 0:
                raise RuntimeError(f"orc_ref_open({so}) failed: {rc}")
        self.have_ref = True
        return True

    def ref_function(self, xcg, name, restype, argtypes):
        addr = self.lib.orc_ref_symbol(xcg, name.encode())
        if not addr:
            raise RuntimeError(f"reference symbol {name} not found")
        return ctypes.CFUNCTYPE(restype, *argtypes)(addr)

    def set_clr_mode(self, from_file):
        """0: CLr table = 0 as the reference binaries compute it; 1: CL1320 data as intended (see f16_oracle.c)."""
        self.lib.orc_set_clr_mode(int(bool(from_file)))

    def trim(self, h, V, fi=1, xcg=0.25, tol=1e-10, maxiter=50000, backend=PORT, ux0=None):
        """env.py::trim(h, V): -> (x_trim [18], {cost, iterations, fcalls, converged}, status)"""
        x = np.zeros(18)
        info = np.zeros(4)
        u0 = None if ux0 is None else np.ascontiguousarray(ux0, dtype=np.float64).ctypes.data_as(c_dp)
        st = self.lib.orc_trim(backend, float(h), float(V), int(fi), float(xcg), u0, float(tol), int(maxiter),
                               x.ctypes.data_as(c_dp), info.ctypes.data_as(c_dp))
        return x, dict(cost=info[0], iterations=int(info[1]), fcalls=int(info[2]), converged=bool(info[3])), int(st)

    def trim_cost(self, ux, h, V, fi=1, xcg=0.25, backend=PORT):
        ux = np.ascontiguousarray(ux, dtype=np.float64)
        c = ctypes.c_double()
        st = self.lib.orc_trim_cost(backend, ux.ctypes.data_as(c_dp), float(h), float(V), int(fi), float(xcg), ctypes.byref(c))
        return c.value, int(st)

    def trim_batch(self, h, V, fi=1, xcg=0.25, tol=1e-10, maxiter=50000, backend=PORT):
        h, V = np.ascontiguousarray(h, dtype=np.float64), np.ascontiguousarray(V, dtype=np.float64)
        n = h.size
        x = np.empty((18, n))
        info = np.empty((4, n))
        st = np.zeros(n, dtype=np.int32)
        self.lib.orc_trim_batch(backend, h.ctypes.data_as(c_dp), V.ctypes.data_as(c_dp), n, int(fi), float(xcg), float(tol),
                                int(maxiter), x.ctypes.data_as(c_dp), info.ctypes.data_as(c_dp), st.ctypes.data_as(c_ip))
        return x, info, st

    def max_threads(self):
        return int(self.lib.orc_max_threads())

    def set_threads(self, n):
        """OpenMP team size of the batch loops (torchrun exports OMP_NUM_THREADS=1)."""
        self.lib.orc_set_threads(int(n))

    # -- scalar pieces --------------------------------------------------------------------------
    def cell(self, axis, x):
        lo, hi = ctypes.c_int(), ctypes.c_int()
        out = self.lib.orc_cell(AXES[axis] if isinstance(axis, str) else axis, x, ctypes.byref(lo), ctypes.byref(hi))
        return out, lo.value, hi.value

    def interp(self, table_id, a, b=0.0, d=0.0):
        return self.lib.orc_interp(table_id, a, b, d)

    def hifi(self, alpha, beta, el):
        out = np.zeros(44)
        self.lib.orc_hifi(alpha, beta, el, _dp(out))
        return out

    def atmos(self, alt, vt):
        out = np.zeros(3)
        self.lib.orc_atmos(alt, vt, _dp(out))
        return out

    # -- batches (SoA: [component][aircraft]) ---------------------------------------------------
    def nlplant_batch(self, xu, fi=1, xcg=0.25, backend=PORT):
        xu = _soa(xu, 17)
        n = xu.shape[1]
        xdot = np.empty((18, n))
        st = np.zeros(n, dtype=np.int32)
        self.lib.orc_nlplant_batch(backend, _dp(xu), _dp(xdot), n, fi, xcg, st.ctypes.data_as(c_ip))
        return xdot, st

    def calc_xdot_batch(self, x, u, fi=1, xcg=0.25, backend=PORT):
        x, u = _soa(x, 18), _soa(u, 4)
        n = x.shape[1]
        xdot = np.empty((18, n))
        st = np.zeros(n, dtype=np.int32)
        self.lib.orc_calc_xdot_batch(backend, _dp(x), _dp(u), _dp(xdot), n, fi, xcg, st.ctypes.data_as(c_ip))
        return xdot, st

    def step_batch(self, x, u, K, dt=0.001, fi=1, xcg=0.25, lqr=None, backend=PORT):
        x = _soa(x, 18).copy()
        u = _soa(u, 4)
        n = x.shape[1]
        st = np.zeros(n, dtype=np.int32)
        law = ctypes.byref(lqr) if lqr is not None else None
        self.lib.orc_step_batch(backend, _dp(x), _dp(u), n, K, dt, fi, xcg, law, st.ctypes.data_as(c_ip))
        return x, st

    def linearise_batch(self, x, u, eps=1e-5, scheme=0, fi=1, xcg=0.25, backend=PORT):
        x, u = _soa(x, 18), _soa(u, 4)
        n = x.shape[1]
        A = np.empty((n, 18, 18))
        B = np.empty((n, 18, 4))
        st = np.zeros(n, dtype=np.int32)
        self.lib.orc_linearise_batch(backend, _dp(x), _dp(u), n, eps, scheme, _dp(A), _dp(B), fi, xcg,
                                     st.ctypes.data_as(c_ip))
        return A, B, st


_ORACLE = None


def get_oracle():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE


# ---- between linearise and the control law: the reference does these with numpy / scipy (env.py:46-60,344-358, utils.py:219-245)
NA_ROWS = [3, 4, 7, 8, 9, 10, 11, 16, 17]      # _calc_xdot_na output rows in full-state numbering (lf dots swapped, env.py:184,189)
NA_COLS = [3, 4, 7, 8, 9, 10, 11, 17, 16]      # parameters.py:135 mpc_states
NA_UCOLS = [13, 14, 15]                        # inputs written into the actuator states (env.py:175-177)


def reduce_jacobian(A):
    """A_na, B_na of F16.linearise(..., _calc_xdot_na) as a gather of the full forward-difference A (bit-identical: both
    difference the same Nlplant evaluations; checked against the reference's own ssr in tests/test_oracle.py)."""
    A = np.asarray(A)
    return A[..., NA_ROWS, :][..., :, NA_COLS], A[..., NA_ROWS, :][..., :, NA_UCOLS]


def discretise(A, B, dt):
    """env.py:46,50: cont2discrete((A, B, C, D), dt)[0:2], scipy's own routine (zero-order hold via expm)"""
    from scipy.signal import cont2discrete
    n, m = A.shape[-1], B.shape[-1]
    Ad, Bd = cont2discrete((A, B, np.eye(n), np.zeros((n, m))), dt)[0:2]
    return Ad, Bd


def dlqr(A, B, Q, R):
    """utils.py:219-245 restated: P = solve_discrete_are(A, B, Q, R); K = inv(B'PB + R) (B'PA)"""
    import scipy.linalg
    P = np.array(scipy.linalg.solve_discrete_are(A, B, Q, R))
    K = np.array(scipy.linalg.inv(B.T @ P @ B + R) @ (B.T @ P @ A))
    return K, P

