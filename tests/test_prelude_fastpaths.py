"""
rn_prelude.cuh: the straight-line common paths of the fdlibm log / exp / pow (rn_strict_*) against the complete functions
(rn_strict_*_full, the transcription of fdlibm that the oracle shares and that the reference's golden vectors pin), bit for bit:
random arguments over all magnitudes, dense neighbourhoods of every range boundary the functions branch on, special values.
The device source is compiled for the host (RN_HOST_EMULATION, g++ -ffp-contract=off), like tests/host_emulation.py does.
"""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SHIM = r"""
extern "C" void fp_exp(const double* x, long long n, double* fast, double* full) { for (long long i = 0; i < n; i++) { fast[i] = rn_strict_exp(x[i]); full[i] = rn_strict_exp_full(x[i]); } }
extern "C" void fp_log(const double* x, long long n, double* fast, double* full) { for (long long i = 0; i < n; i++) { fast[i] = rn_strict_log(x[i]); full[i] = rn_strict_log_full(x[i]); } }
extern "C" void fp_pow(const double* x, const double* y, long long n, double* fast, double* full) { for (long long i = 0; i < n; i++) { fast[i] = rn_strict_pow(x[i], y[i]); full[i] = rn_strict_pow_full(x[i], y[i]); } }
"""


def _lib():
    src = open(os.path.join(ROOT, "rainier_b200", "csrc", "rn_prelude.cuh")).read() + _SHIM
    d = os.path.join(tempfile.gettempdir(), "rn_emul")
    os.makedirs(d, exist_ok=True)
    key = hashlib.sha1(src.encode()).hexdigest()[:16]
    so = os.path.join(d, "prelude_" + key + ".so")
    if not os.path.exists(so):
        cpp = os.path.join(d, "prelude_" + key + ".cpp")
        open(cpp, "w").write(src)
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-DRN_HOST_EMULATION", "-w", "-ffp-contract=off", cpp, "-o", so], check=True)
    return C.CDLL(so)


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _same(fast, full):
    nan = np.isnan(fast) & np.isnan(full)  # any NaN is the same value to Java
    return bool(np.all((_bits(fast) == _bits(full)) | nan))


def _around(hi_words, rng, width=3):
    """doubles whose high word is within `width` of each given constant, random low words, both signs"""
    out = []
    for h in hi_words:
        for dh in range(-width, width + 1):
            lo = np.concatenate([rng.integers(0, 2 ** 32, 200, dtype=np.uint64), np.array([0, 1, 2 ** 32 - 1], dtype=np.uint64)])
            v = ((np.uint64(h + dh) << np.uint64(32)) | lo).view(np.float64)
            out += [v, -v]
    return np.concatenate(out)


def _wide(rng, n, lo_exp=-1074, hi_exp=1023):
    m = rng.uniform(1.0, 2.0, n)
    e = rng.integers(lo_exp, hi_exp + 1, n)
    return np.ldexp(m, e) * rng.choice([-1.0, 1.0], n)


SPECIAL = np.array([0.0, -0.0, 1.0, -1.0, 2.0, 0.5, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
                    709.782712893384, 709.7827128933841, -745.1332191019411, -745.1332191019412, -708.3964185322641, 0.34657359027997264,
                    1.0397207708399179, 2 ** -28, 2 ** -29, 1 - 2 ** -53, 1 + 2 ** -52, 1 + 2 ** -20, 1 - 2 ** -21])


def _run1(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    a, b = np.empty_like(x), np.empty_like(x)
    fn(x.ctypes.data_as(C.c_void_p), C.c_longlong(len(x)), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    return a, b


def test_exp_fast_path_is_fdlibm_bit_for_bit():
    L, rng = _lib(), np.random.default_rng(1)
    x = np.concatenate([SPECIAL, rng.normal(size=2_000_000) * 3, rng.uniform(-750, 720, 2_000_000), _wide(rng, 500_000, -40, 11),
                        _around([0x3e300000, 0x3fd62e42, 0x3FF0A2B2, 0x40862E42, 0x7ff00000, 0x40874385, 0x40862000], rng)])
    assert _same(*_run1(L.fp_exp, x))


def test_log_fast_path_is_fdlibm_bit_for_bit():
    L, rng = _lib(), np.random.default_rng(2)
    x = np.concatenate([SPECIAL, rng.uniform(0, 1, 2_000_000), rng.uniform(0, 1, 1_000_000) ** 8, np.abs(_wide(rng, 2_000_000)), _wide(rng, 100_000),
                        _around([0x00100000, 0x000fffff, 0x7ff00000, 0x3ff00000, 0x3fefffff, 0x3ff6147a, 0x3ff6b851, 0x3fe6a09e, 0x3ff6a09e, 0x3fe00000], rng),
                        1.0 + rng.uniform(-3e-6, 3e-6, 500_000)])
    assert _same(*_run1(L.fp_log, x))


def test_pow_fast_path_is_fdlibm_bit_for_bit():
    L, rng = _lib(), np.random.default_rng(3)
    n = 1_500_000
    xs = [np.abs(_wide(rng, n, -30, 30)), np.exp(rng.normal(size=n) * 2), np.abs(_wide(rng, n)), _wide(rng, 200_000, -5, 5),
          np.abs(_around([0x00100000, 0x3ff00000, 0x7ff00000, 0x3ff3988e, 0x3ffbb67a, 0x3fe00000], rng))]
    ys = [rng.choice([-2.0, -0.75, 3.0, -1.5, 0.25, 7.0, -0.5, 1.5], n), rng.normal(size=n) * 2, rng.normal(size=n) * 40,
          rng.choice([-2.0, 2.0, 3.0, 0.5, -1.0, 1.0, 0.0, 1e10, -1e10, np.inf, np.nan, 2.0 ** 31, 2.0 ** 32, 1e-30], 200_000), None]
    ys[4] = rng.choice([-2.0, -0.75, 1.7, 1e3, -1e3], len(xs[4]))
    x = np.concatenate(xs + [np.repeat(SPECIAL, len(SPECIAL))])
    y = np.concatenate(ys + [np.tile(SPECIAL, len(SPECIAL))])
    # results near overflow / underflow and subnormal results
    x2 = np.exp(rng.uniform(-5, 5, n))
    y2 = rng.uniform(-1100, 1100, n) * np.log(2) / np.log(x2)
    x, y = np.concatenate([x, x2]), np.concatenate([y, y2])
    a, b = np.empty_like(x), np.empty_like(x)
    L.fp_pow(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_longlong(len(x)), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    assert _same(a, b)
    taken = np.isfinite(b) & (x > 0) & (b != 0)
    assert taken.mean() > 0.5  # the comparison is not vacuous: most arguments have finite non-zero results
