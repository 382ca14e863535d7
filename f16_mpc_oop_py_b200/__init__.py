"""f16_mpc_oop_py_b200 -- B200-native batched F-16 plant behind the reference's ctypes C ABI.

Only the hot path of johnviljoen/f16_mpc_oop_py lives here: Nlplant / atmos, the aero-table interpolation, the
env.py Euler step and the finite-difference linearise, as hand-written sm_100a CUDA in libf16_b200.so.  Importing
this package loads that library and fails if it has not been built; there is no CPU fallback.
"""
from . import parameters
from . import shard
from ._lib import (CLR_AS_BUILT, CLR_FROM_FILE, FD_CENTRAL, FD_FORWARD, MATH_FAST, MATH_STRICT, F16Error, LqrLaw, init,
                   init_devices, lib, shutdown)
from .plant import (F16Batch, atmos, discretise, dlqr, make_lqr, nlplant, reduce_jacobian, state_summary, state_summary_dev,
                    trim)

__all__ = ["F16Batch", "F16Error", "LqrLaw", "atmos", "init", "init_devices", "shutdown", "lib", "make_lqr", "nlplant", "trim", "reduce_jacobian", "discretise", "dlqr", "state_summary", "state_summary_dev", "parameters",
           "MATH_STRICT", "MATH_FAST", "CLR_AS_BUILT", "CLR_FROM_FILE", "FD_FORWARD", "FD_CENTRAL"]
