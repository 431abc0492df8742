"""
The device forms of the arithmetic helpers of rn_prelude.cuh, compiled for the HOST (RN_HOST_MODEL_DEVICE_ARITH): the check-free
division / square root (CUDA's inline Newton sequences) and the branch-free row functions, with the two MUFU seeds modelled as a
20-bit reciprocal / reciprocal square root of the operand's high word -- coarser than the hardware's.  What this checks without a
GPU: everything above the seed (Newton steps + residual correction give the correctly rounded quotient / root over the stated
domain; argument reduction, polynomials and special-value selects of exp / log / reciprocal).  The bit-for-bit comparison with the
hardware seeds is tests/test_gpu_divsqrt.py (-m gpu).
"""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SHIM = r"""
extern "C" void md_div(const double* a, const double* b, long long n, double* got, double* want) { for (long long i = 0; i < n; i++) { got[i] = rn_div_nc(a[i], b[i]); want[i] = a[i] / b[i]; } }
extern "C" void md_sqrt(const double* a, const double*, long long n, double* got, double* want) { for (long long i = 0; i < n; i++) { got[i] = rn_sqrt_nc(a[i]); want[i] = sqrt(a[i]); } }
extern "C" void md_exp(const double* a, const double*, long long n, double* got, double* want) { for (long long i = 0; i < n; i++) { got[i] = rn_row_exp(a[i]); want[i] = exp(a[i]); } }
extern "C" void md_log(const double* a, const double*, long long n, double* got, double* want) { for (long long i = 0; i < n; i++) { got[i] = rn_row_log(a[i]); want[i] = log(a[i]); } }
extern "C" void md_rcp(const double* a, const double*, long long n, double* got, double* want) { for (long long i = 0; i < n; i++) { got[i] = rn_row_rcp(a[i]); want[i] = 1.0 / a[i]; } }
"""


def _lib():
    src = open(os.path.join(ROOT, "rainier_b200", "csrc", "rn_prelude.cuh")).read() + _SHIM
    d = os.path.join(tempfile.gettempdir(), "rn_emul")
    os.makedirs(d, exist_ok=True)
    key = hashlib.sha1(("model" + src).encode()).hexdigest()[:16]
    so = os.path.join(d, "devarith_" + key + ".so")
    if not os.path.exists(so):
        cpp = os.path.join(d, "devarith_" + key + ".cpp")
        open(cpp, "w").write(src)
        fma = ["-mfma"] if "fma" in open("/proc/cpuinfo").read().split("flags", 1)[-1].split("\n", 1)[0].split() else []
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-DRN_HOST_EMULATION", "-DRN_HOST_MODEL_DEVICE_ARITH", "-w",
                        "-ffp-contract=off"] + fma + [cpp, "-o", so], check=True)
    return C.CDLL(so)


def _run(fn, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b if b is not None else a, dtype=np.float64)
    got, want = np.empty_like(a), np.empty_like(a)
    fn(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_longlong(len(a)), got.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p))
    return got, want


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float64).view(np.int64)


def _rand_exp(rng, n, lo, hi, signed=True):
    v = np.ldexp(rng.uniform(1.0, 2.0, n), rng.integers(lo, hi + 1, n))
    return v * rng.choice([-1.0, 1.0], n) if signed else v


def _assert_identical(got, want, what):
    bad = np.nonzero(_bits(got) != _bits(want))[0]
    assert len(bad) == 0, "%s: %d of %d differ, first: got %r want %r" % (what, len(bad), len(got), got[bad[0]].hex(), want[bad[0]].hex())


def _assert_close(got, want, x, tol_ulp, what):
    nan = np.isnan(want) | np.isnan(got)
    assert np.array_equal(np.isnan(want), np.isnan(got)), what + ": NaN classes differ"
    special = ~nan & (np.isinf(want) | np.isinf(got) | (want == 0) | (got == 0))
    tiny = special & (np.abs(want) < 2.3e-308) & (np.abs(got) < 2.3e-308)  # underflow neighbourhood: one unit of the subnormal grid
    exact = special & ~tiny
    assert np.array_equal(_bits(got[exact]), _bits(want[exact])), what + ": special values differ"
    rest = ~nan & ~exact
    d = np.abs(_bits(got[rest]) - _bits(want[rest]))
    worst = int(d.max()) if len(d) else 0
    assert worst <= tol_ulp, "%s: %d ulp at x = %r" % (what, worst, x[rest][int(d.argmax())].hex())
    return worst


def test_check_free_division_is_correctly_rounded_with_a_20_bit_seed():
    L, rng = _lib(), np.random.default_rng(21)
    n = 3_000_000
    ea, eq = rng.integers(-960, 1000, n), rng.integers(-1000, 1000, n)
    eb = np.clip(ea - eq, -1000, 1000)
    keep = np.abs(ea - eb) <= 1015
    a = (np.ldexp(rng.uniform(1, 2, n), ea) * rng.choice([-1.0, 1.0], n))[keep]
    b = (np.ldexp(rng.uniform(1, 2, n), eb) * rng.choice([-1.0, 1.0], n))[keep]
    _assert_identical(*_run(L.md_div, a, b), "a / b over the stated domain")
    ia, ib = rng.integers(1, 1 << 26, n // 2).astype(np.float64), rng.integers(1, 1 << 26, n // 2).astype(np.float64)
    for x, y, what in ((ia, ib, "small integers"), (ia * ib, ib, "exact quotients"), (ia * ib + 1.0, ib, "next to exact quotients")):
        _assert_identical(*_run(L.md_div, x, y), what)
    s = _rand_exp(rng, n // 2, -104, -1, signed=False)  # the polar method: -2 log(s) / s
    _assert_identical(*_run(L.md_div, -2.0 * np.log(s), s), "polar method")
    f = rng.uniform(-0.2929, 0.4143, n // 2)
    f = np.where(np.abs(f) < 2.0 ** -20, 2.0 ** -20, f)
    _assert_identical(*_run(L.md_div, f, 2.0 + f), "fdlibm log: f / (2 + f)")


def test_check_free_square_root_is_correctly_rounded_with_a_20_bit_seed():
    L, rng = _lib(), np.random.default_rng(22)
    n = 3_000_000
    _assert_identical(*_run(L.md_sqrt, _rand_exp(rng, n, -970, 1023, signed=False)), "sqrt over the stated domain")
    i = rng.integers(1, 1 << 26, n // 2).astype(np.float64)
    for x, what in ((i * i, "exact roots"), (i * i + 1.0, "above exact roots"), (i * i - 1.0, "below exact roots")):
        _assert_identical(*_run(L.md_sqrt, x), what)


SPECIAL = np.array([0.0, -0.0, 1.0, -1.0, 2.0, 0.5, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
                    -1.7976931348623157e308, 709.782712893384, 709.7827128933841, 709.9, 1000.0, -745.1332191019411, -745.1332191019412, -746.0,
                    -1000.0, -708.5, -740.0, 2.0 ** -1000, 2.0 ** 1000, 2.0 ** -1030, 2.0 ** 1023, -(2.0 ** -1030), 3.0, 1 - 2.0 ** -53, 1 + 2.0 ** -52])


def test_row_functions_against_libm():
    L, rng = _lib(), np.random.default_rng(23)
    n = 2_000_000
    x = np.concatenate([SPECIAL, rng.normal(size=n) * 5, rng.uniform(-750, 720, n), _rand_exp(rng, n // 2, -60, 11)])
    _assert_close(*_run(L.md_exp, x), x, 2, "rn_row_exp")
    x = np.concatenate([SPECIAL, _rand_exp(rng, n, -1074, 1023, signed=False), rng.uniform(0, 2, n), 1.0 + rng.uniform(-0.5, 0.5, n // 2) * 2.0 ** -18,
                        -_rand_exp(rng, 1000, -100, 100, signed=False)])
    _assert_close(*_run(L.md_log, x), x, 2, "rn_row_log")
    x = np.concatenate([SPECIAL, _rand_exp(rng, n, -1074, 1023), rng.normal(size=n)])
    _assert_close(*_run(L.md_rcp, x), x, 1, "rn_row_rcp")
    x = _rand_exp(rng, n, -950, 950)
    assert _assert_close(*_run(L.md_rcp, x), x, 0, "rn_row_rcp, no scaling") == 0
