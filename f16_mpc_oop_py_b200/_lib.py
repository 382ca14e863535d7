"""ctypes binding of libf16_b200.so (include/f16_b200.h).  There is no fallback: if the shared library is missing
the import fails, and if no B200 is usable every call raises F16Error."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libf16_b200.so")
TABLE_BLOB = os.path.join(HERE, "data", "f16_aero_v1.bin")

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int)
c_ubp = ctypes.POINTER(ctypes.c_ubyte)
c_ll = ctypes.c_longlong
c_vp = ctypes.c_void_p

F16_OK, F16_ERR_CUDA, F16_ERR_TABLES, F16_ERR_ARG, F16_ERR_NOINIT, F16_ERR_HOST = 0, -1, -2, -3, -4, -5
MATH_STRICT, MATH_FAST = 0, 1
CLR_AS_BUILT, CLR_FROM_FILE = 0, 1
FD_FORWARD, FD_CENTRAL = 0, 1
ST_ALPHA, ST_BETA, ST_DELE, ST_NAN, ST_FIDELITY = 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22


class F16Error(RuntimeError):
    pass


class LqrLaw(ctypes.Structure):
    """f16_lqr_t"""
    _fields_ = [
        ("n_sel", ctypes.c_int),
        ("row_mask", ctypes.c_int),
        ("sel", ctypes.c_int * 18),
        ("K", (ctypes.c_double * 18) * 4),
        ("x_ref", ctypes.c_double * 18),
        ("u0", ctypes.c_double * 4),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C f16_mpc_oop_py_b200/csrc` (or __graft_entry__.build()). "
            "f16_mpc_oop_py_b200 has no CPU implementation to fall back to.")
    L = ctypes.CDLL(LIB_PATH)
    sel = [c_ubp, ctypes.c_int, c_dp, ctypes.c_double]
    L.Nlplant.argtypes = [c_vp, c_vp, ctypes.c_int]
    L.Nlplant.restype = None
    L.atmos.argtypes = [ctypes.c_double, ctypes.c_double, c_vp]
    L.atmos.restype = None
    L.f16_nlplant_xcg.argtypes = [c_vp, c_vp, ctypes.c_int, ctypes.c_double]
    L.f16_nlplant_xcg.restype = None
    L.f16_atmos.argtypes = [ctypes.c_double, ctypes.c_double, c_vp]
    L.f16_atmos.restype = None
    L.f16_init.argtypes = [ctypes.c_char_p, ctypes.c_int]
    L.f16_init_devices.argtypes = [ctypes.c_char_p, c_ip, ctypes.c_int]
    L.f16_use_device.argtypes = [ctypes.c_int]
    L.f16_set_host_pipeline.argtypes = [ctypes.c_int]
    L.f16_plan_slices.argtypes = [c_ll, c_ll, ctypes.c_int, ctypes.POINTER(c_ll)]
    L.f16_plan_chunks.argtypes = [c_ll, c_ll, ctypes.c_int, ctypes.POINTER(c_ll), c_ip]
    L.f16_set_step_compaction.argtypes = [ctypes.c_int]
    L.f16_set_trim_fixed_point_exit.argtypes = [ctypes.c_int]
    L.f16_shutdown.restype = None
    L.f16_last_error.restype = ctypes.c_char_p
    L.f16_set_default_xcg.argtypes = [ctypes.c_double]
    L.f16_set_default_xcg.restype = None
    L.f16_tables_sha256.argtypes = [ctypes.c_char_p]
    L.Nlplant_batch.argtypes = [c_vp, c_vp] + sel + [c_ll, c_vp]
    L.calc_xdot_batch.argtypes = [c_vp, c_vp, c_vp] + sel + [c_ll, c_vp]
    L.step_batch.argtypes = [c_vp, c_vp, c_ll, ctypes.c_int, ctypes.c_double, ctypes.POINTER(LqrLaw)] + sel + [c_vp, c_vp]
    L.step_batch_traj.argtypes = [c_vp, c_vp, c_ll, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.POINTER(LqrLaw)] + sel + [c_vp, c_vp]
    L.linearise_batch.argtypes = [c_vp, c_vp, c_ll, ctypes.c_double, ctypes.c_int, c_vp, c_vp] + sel + [c_vp]
    L.Nlplant_batch_dev.argtypes = [c_vp, c_ll, c_vp, c_ll] + [c_vp, ctypes.c_int, c_vp, ctypes.c_double] + [c_ll, c_vp]
    L.calc_xdot_batch_dev.argtypes = [c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_vp, ctypes.c_int, c_vp, ctypes.c_double,
                                      c_ll, c_vp]
    L.step_batch_dev.argtypes = [c_vp, c_ll, c_vp, c_ll, c_ll, ctypes.c_int, ctypes.c_double, ctypes.POINTER(LqrLaw),
                                 c_vp, ctypes.c_int, c_vp, ctypes.c_double, c_vp, c_vp]
    L.linearise_batch_dev.argtypes = [c_vp, c_ll, c_vp, c_ll, c_ll, ctypes.c_double, ctypes.c_int, c_vp, c_vp, c_vp,
                                      ctypes.c_int, c_vp, ctypes.c_double, c_vp]
    L.trim_batch.argtypes = [c_vp, c_vp, c_ll, ctypes.c_double, ctypes.c_int, c_vp, c_vp, c_vp] + sel + [c_vp]
    L.trim_batch_dev.argtypes = [c_vp, c_vp, c_ll, ctypes.c_double, ctypes.c_int, c_vp, c_vp, c_ll, c_vp, c_ll, c_vp, ctypes.c_int,
                                 c_vp, ctypes.c_double, c_vp]
    L.f16_set_linearise_variant.argtypes = [ctypes.c_int]
    L.f16_set_step_chunking.argtypes = [ctypes.c_int]
    L.reduce_jacobian_batch.argtypes = [c_vp, c_ll, c_vp, c_vp]
    L.discretise_batch.argtypes = [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_ll, ctypes.c_double, c_vp, c_vp]
    L.dlqr_batch.argtypes = [c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_ll, c_vp, c_vp, c_vp]
    L.lqr_gain_batch.argtypes = [c_vp, c_vp, c_ll, ctypes.c_double, c_vp] + sel + [c_vp]
    L.state_summary_batch.argtypes = [c_vp, c_ll, c_vp, c_vp]
    L.state_summary_batch_dev.argtypes = [c_vp, c_ll, c_ll, c_vp, c_vp]
    L.step_batch_stats.argtypes = [c_vp, c_vp, c_ll, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.POINTER(LqrLaw)] + sel + [c_vp, c_vp]
    L.step_batch_stats_dev.argtypes = [c_vp, c_ll, c_vp, c_ll, c_ll, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                       ctypes.POINTER(LqrLaw)] + sel + [c_vp, c_vp]
    L.reduce_jacobian_batch_dev.argtypes = [c_vp, c_ll, c_vp, c_vp]
    L.discretise_batch_dev.argtypes = [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_ll, ctypes.c_double, c_vp, c_vp]
    L.dlqr_batch_dev.argtypes = [c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_ll, c_vp, c_vp, c_vp]
    L.f16_hifi_probe.argtypes = [c_vp, c_vp, c_vp, c_ll, c_vp, c_vp, c_vp]
    L.f16_fast_probe.argtypes = [c_vp, c_vp, c_vp, c_ll, c_vp, c_vp, c_vp, c_vp]
    L.f16_lofi_probe.argtypes = [c_vp] * 5 + [c_ll, c_vp]
    L.atmos_batch.argtypes = [c_vp, c_vp, c_ll, c_vp]
    L.f16_div_probe.argtypes = [c_vp, c_vp, c_ll, c_vp]
    L.f16_dev_alloc.argtypes = [ctypes.c_ulonglong]
    L.f16_dev_alloc.restype = c_vp
    L.f16_dev_free.argtypes = [c_vp]
    L.f16_dev_free.restype = None
    L.f16_host_alloc_pinned.argtypes = [ctypes.c_ulonglong]
    L.f16_host_alloc_pinned.restype = c_vp
    L.f16_host_free_pinned.argtypes = [c_vp]
    L.f16_host_free_pinned.restype = None
    L.f16_memcpy_h2d.argtypes = [c_vp, c_vp, ctypes.c_ulonglong]
    L.f16_memcpy_d2h.argtypes = [c_vp, c_vp, ctypes.c_ulonglong]
    L.f16_memcpy_d2d.argtypes = [c_vp, c_vp, ctypes.c_ulonglong]
    L.f16_memset_dev.argtypes = [c_vp, ctypes.c_int, ctypes.c_ulonglong]
    L.f16_stream.restype = c_vp
    L.f16_timer_stop.argtypes = [ctypes.POINTER(ctypes.c_float)]
    L.f16_launch_count.restype = ctypes.c_ulonglong
    L.f16_measure_fp64_peak.argtypes = [ctypes.c_double, c_dp]
    return L


lib = _load()


def check(rc, what):
    if rc != F16_OK:
        msg = lib.f16_last_error()
        raise F16Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def init(table_path=None, device=-1):
    """f16_init: the packed blob shipped with the package unless told otherwise."""
    path = table_path if table_path is not None else (os.environ.get("F16_TABLE_PATH") or TABLE_BLOB)
    check(lib.f16_init(path.encode(), device), "f16_init")


def init_devices(devices=None, table_path=None):
    """f16_init_devices: one context per CUDA ordinal in `devices` (None = every visible GPU); the host-buffer batch calls then
    split their aircraft over the contexts.  Returns the number of contexts."""
    path = table_path if table_path is not None else (os.environ.get("F16_TABLE_PATH") or TABLE_BLOB)
    if devices is None:
        check(lib.f16_init_devices(path.encode(), None, 0), "f16_init_devices")
    else:
        arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        check(lib.f16_init_devices(path.encode(), arr, len(devices)), "f16_init_devices")
    return lib.f16_device_count()


def shutdown():
    lib.f16_shutdown()
