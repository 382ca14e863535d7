"""div_by (csrc/f16_model.cuh): a / y through the correctly rounded reciprocal and one residual correction must give the
bits of the IEEE quotient -- the strict (parity) build uses it for every division by a constant, by a table cell width
and by the finite-difference step of linearise (env.py:330,339).  Checked in exact rational arithmetic on the hardest
numerators (quotients within ~2^-106 relative of a rounding midpoint or of a representable number, built with a modular
inverse of the divisor's significand) and on the host compile of the device function (tests/hostemu)."""
import ctypes
import math
import os
import random
import subprocess
from fractions import Fraction as F

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# every divisor the strict build hands to div_by: nlplant.c constants (m, Jy, Jx Jz - Jxz^2, 21.5, 30, 25), lofi (12, 25,
# 57.3), utils.py:293 (pi), the table cell widths (5, 2, 15, 10, 25) and the finite-difference steps eps and 2 eps
DIVISORS = [636.94, 55814.0, 9496.0 * 63100.0 - 982.0 * 982.0, 21.5, 30.0, 25.0, 12.0, 57.3, 3.141592653589793,
            5.0, 2.0, 15.0, 10.0, 1e-5, 2e-5, 1e-6, 2e-6, 1e-4, 2e-4, 1e-7, 1e-3]


def fl(x):
    return float(x)  # Fraction -> double, correctly rounded


def div_by_exact(a, y, r):
    q = a * r
    rem = fl(F(a) - F(q) * F(y))
    return fl(F(q) + F(rem) * F(r))


def hard_numerators(y, count, rnd):
    m, _ = math.frexp(y)
    Y = int(m * (1 << 53))
    Yo = Y >> ((Y & -Y).bit_length() - 1)
    inv = pow(Yo, -1, 1 << 60)
    out = []
    for _ in range(count):
        delta = rnd.randint(-6, 6)
        n = (delta * inv) % (1 << 54)      # Yo n == delta (mod 2^54): A / Yo sits delta / (Yo 2^54) from n / 2^54
        if n < (1 << 53):
            n |= 1 << 53
        A = (Yo * n) >> 54
        for dA in (0, 1):
            if (A + dA).bit_length() <= 53:
                a = float(A + dA) * 2.0 ** rnd.randint(-30, 30)
                out.append(a if rnd.random() < 0.5 else -a)
    return out


def test_generator_has_teeth():
    rnd = random.Random(5)
    y = 1e-5
    cases = hard_numerators(y, 500, rnd)
    wrong = sum(1 for a in cases if a * (1.0 / y) != fl(F(a) / F(y)))
    assert wrong > len(cases) // 10          # the uncorrected product misrounds a large share of them
    closest = min(abs(abs(F(a) / F(y) - F(fl(F(a) / F(y)))) / F(math.ulp(fl(F(a) / F(y)))) - F(1, 2)) for a in cases[:100])
    assert closest < 1e-12                   # ... because they sit on top of rounding midpoints


@pytest.mark.parametrize("y", DIVISORS)
def test_one_correction_is_the_ieee_quotient_exact_arithmetic(y):
    rnd = random.Random(int(y * 1e7) & 0xFFFF)
    r = 1.0 / y
    cases = hard_numerators(y, 600, rnd) + [rnd.uniform(-1, 1) * 10.0 ** rnd.randint(-12, 8) for _ in range(600)]
    for a in cases:
        assert div_by_exact(a, y, r) == fl(F(a) / F(y)), (a, y)


def test_random_divisors_exact_arithmetic():
    rnd = random.Random(11)
    for _ in range(40):
        y = rnd.uniform(1, 2) * 2.0 ** rnd.randint(-40, 20)
        r = 1.0 / y
        for a in hard_numerators(y, 150, rnd):
            assert div_by_exact(a, y, r) == fl(F(a) / F(y)), (a, y)


def test_device_function_host_compile():
    """the host compile of div_by: IEEE bits on the hard numerators, on ordinary ones from 1e-290 up, on 0, -0, Inf, NaN and
    on denormal quotients up to their last place; below |a| = 2^-969 (the residual underflows) at most one ulp off"""
    d = os.path.join(REPO, "tests", "hostemu")
    subprocess.run(["make", "-s", "-C", d], check=True)
    E = ctypes.CDLL(os.path.join(d, "libf16_hostemu.so"))
    E.emu_div_by.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_double, ctypes.c_void_p]
    E.emu_div_by.restype = None
    rnd = random.Random(3)
    rng = np.random.default_rng(3)

    def run(a, y):
        out = np.empty_like(a)
        E.emu_div_by(a.ctypes.data, a.size, y, out.ctypes.data)
        with np.errstate(all="ignore"):
            return out, a / y

    for y in DIVISORS:
        a = np.array(hard_numerators(y, 20000, rnd) + [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-290, -1e-290, 1e300, 1.7e308 * min(y, 1.0),
                                                        2.0 ** -900, 2.0 ** 900 * min(y, 1.0), y, -y, 3 * y])
        a = np.concatenate([a, rng.uniform(-1, 1, 200000) * 10.0 ** rng.integers(-14, 9, 200000),
                            rng.uniform(-1, 1, 2000) * 10.0 ** rng.integers(-290, 290, 2000).astype(float)])
        out, ref = run(a, y)
        fin = ~np.isnan(ref)
        assert np.array_equal(out[fin].view(np.uint64), ref[fin].view(np.uint64)) and np.isnan(out[~fin]).all(), y
        # the documented edge: numerators below 2^-969 -- never more than one unit in the last place away
        tiny = np.concatenate([rng.uniform(-1, 1, 4000) * 10.0 ** rng.integers(-323, -292, 4000).astype(float), [5e-324, -5e-324, 1e-310]])
        out, ref = run(tiny, y)
        ulp = np.maximum(np.abs(np.spacing(ref)), 5e-324)
        assert np.all(np.abs(out - ref) <= ulp), y
        assert np.mean(out == ref) > 0.9
