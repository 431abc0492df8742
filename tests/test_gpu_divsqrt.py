"""
Device-side bit-for-bit checks of the arithmetic helpers the parity kernels rest on (round 2, second session):
  * rn_div_nc / rn_sqrt_nc -- CUDA's inline division / square-root sequences without the range test and the conditional
    call of the complete routine -- against the `/` and sqrt() operators, over the whole domain the sequences are stated
    for (rn_prelude.cuh) and, densely, over the operand ranges of every call site;
  * rn_strict_exp / log / pow -- the speculative straight-line common paths with their coefficients in the constant bank
    -- against the complete fdlibm transcriptions (rn_strict_*_full, the functions the golden vectors pin through the
    oracle), the device-side twin of tests/test_prelude_fastpaths.py.
The kernels are scripts/probes/divsqrt_probe.cu, which includes rainier_b200/csrc/rn_prelude.cuh itself.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def build_probe():
    so = os.path.join(ROOT, "build", "libdivsqrt_probe.so")
    src = os.path.join(ROOT, "scripts", "probes", "divsqrt_probe.cu")
    pre = os.path.join(ROOT, "rainier_b200", "csrc", "rn_prelude.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(pre)):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-w",
                        "-shared", "-Xcompiler", "-fPIC", src, "-o", so], check=True)
    return so


@pytest.fixture(scope="module")
def probe():
    import torch
    L = C.CDLL(build_probe())
    L.divsqrt_probe_run.restype = C.c_int
    L.divsqrt_probe_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]

    def run(which, a, b=None):
        a = a.contiguous()
        b = b.contiguous() if b is not None else None
        bad = torch.zeros(1, dtype=torch.int64, device="cuda")
        first = torch.full((1,), 2 ** 62, dtype=torch.int64, device="cuda")
        rc = L.divsqrt_probe_run(which, a.data_ptr(), b.data_ptr() if b is not None else None, a.numel(), bad.data_ptr(), first.data_ptr())
        assert rc == 0
        nbad = int(bad.item())
        if nbad:
            i = int(first.item())
            raise AssertionError("%d of %d results differ; first at %d: a=%r b=%r" % (nbad, a.numel(), i, a[i].item().hex(),
                                                                                     b[i].item().hex() if b is not None else None))
    return run


def _rand_exp(gen, n, lo, hi, signed=True):
    """doubles m * 2^e, m uniform in [1, 2) with random low bits, e uniform integer in [lo, hi]"""
    import torch
    m = 1.0 + torch.rand(n, dtype=torch.float64, device="cuda", generator=gen)
    e = torch.randint(lo, hi + 1, (n,), device="cuda", generator=gen).to(torch.float64)
    v = torch.ldexp(m, e)
    if signed:
        v = v * (torch.randint(0, 2, (n,), device="cuda", generator=gen).to(torch.float64) * 2 - 1)
    return v


def test_div_nc_equals_the_division_operator(probe):
    import torch
    g = torch.Generator(device="cuda").manual_seed(11)
    n = 8_000_000
    # the stated domain: |a| >= 2^-969, |b| < 2^1017, quotient exponent field in [1, 0x7f7]; operands drawn so that it holds
    ea = torch.randint(-960, 1000, (n,), device="cuda", generator=g)
    eq = torch.randint(-1000, 1000, (n,), device="cuda", generator=g)  # target exponent of the quotient
    eb = torch.clamp(ea - eq, -1000, 1000)
    keep = ((ea - eb) >= -1015) & ((ea - eb) <= 1015)
    m = lambda: 1.0 + torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    s = lambda: torch.randint(0, 2, (n,), device="cuda", generator=g).to(torch.float64) * 2 - 1
    a = (torch.ldexp(m(), ea.to(torch.float64)) * s())[keep]
    b = (torch.ldexp(m(), eb.to(torch.float64)) * s())[keep]
    probe(0, a, b)
    # the call sites: exp  r c / (2 - c);  log  f / (2 + f);  pow  1 / (ax + bp), z t1 / (t1 - 2);  polar  -2 log(s) / s
    probe(0, _rand_exp(g, n // 4, -160, -3, signed=False), 2.0 - (torch.rand(n // 4, dtype=torch.float64, device="cuda", generator=g) - 0.5) * 0.72)
    f = (torch.rand(n // 4, dtype=torch.float64, device="cuda", generator=g) * 0.7072 - 0.2929)
    f = torch.where(f.abs() < 2.0 ** -20, torch.full_like(f, 2.0 ** -20), f)
    probe(0, f, 2.0 + f)
    probe(0, torch.ones(n // 4, dtype=torch.float64, device="cuda"), 2.0 + torch.rand(n // 4, dtype=torch.float64, device="cuda", generator=g) * 1.5)
    sv = _rand_exp(g, n // 4, -104, -1, signed=False)
    probe(0, -2.0 * torch.log(sv), sv)
    # exact quotients, ties and near-ties of the last bit: small-integer operands scaled by powers of two
    ia = torch.randint(1, 1 << 26, (n // 4,), device="cuda", generator=g).to(torch.float64)
    ib = torch.randint(1, 1 << 26, (n // 4,), device="cuda", generator=g).to(torch.float64)
    probe(0, ia, ib)
    probe(0, ia * ib, ib)
    probe(0, ia * ib + 1.0, ib)
    # the edges of the stated domain
    edge_a = torch.tensor([2.0 ** -969, 2.0 ** -969 * 1.5, 2.0 ** 1000, 1.0, 3.0, 2.0 ** -960], dtype=torch.float64, device="cuda")
    edge_b = torch.tensor([2.0 ** 40, 2.0 ** 45, 2.0 ** 1016 * 1.99, 2.0 ** 1016, 2.0 ** -1000, 2.0 ** -1020 * 1.0], dtype=torch.float64, device="cuda")
    aa, bb = torch.meshgrid(edge_a, edge_b, indexing="ij")
    q = aa / bb
    ok = (q.abs() >= 2.0 ** -1021) & (q.abs() < 2.0 ** 1016)
    probe(0, aa[ok], bb[ok])


def test_sqrt_nc_equals_the_sqrt_operator(probe):
    import torch
    g = torch.Generator(device="cuda").manual_seed(12)
    n = 8_000_000
    probe(1, _rand_exp(g, n, -970, 1023, signed=False))  # the stated domain: 2^-970 <= x < inf
    probe(1, _rand_exp(g, n // 2, -53, 112, signed=False))  # the polar method's quotients
    i = torch.randint(1, 1 << 26, (n // 4,), device="cuda", generator=g).to(torch.float64)
    probe(1, i * i)  # exact roots
    probe(1, i * i + 1.0)
    probe(1, i * i - 1.0)
    probe(1, torch.tensor([2.0 ** -970, 2.0 ** -969, 1.7976931348623157e308, 1.0, 2.0, 4.0, 0.25], dtype=torch.float64, device="cuda"))


def test_fdlibm_common_paths_on_the_device(probe):
    import torch
    g = torch.Generator(device="cuda").manual_seed(13)
    n = 6_000_000
    special = torch.tensor([0.0, -0.0, 1.0, -1.0, 2.0, 0.5, float("inf"), float("-inf"), float("nan"), 5e-324, -5e-324, 2.2250738585072014e-308,
                            1.7976931348623157e308, 709.782712893384, 709.7827128933841, -745.1332191019411, -745.1332191019412,
                            -708.3964185322641, 0.34657359027997264, 1.0397207708399179, 2.0 ** -28, 2.0 ** -29, 1 - 2.0 ** -53, 1 + 2.0 ** -52,
                            1 + 2.0 ** -20, 1 - 2.0 ** -21, 4.0, 0.25, 16.0, 1024.0], dtype=torch.float64, device="cuda")
    nan_neg = torch.tensor([-1], dtype=torch.int64, device="cuda").view(torch.float64)  # a sign-set NaN (round 2's rn_k_eval bug)
    x_exp = torch.cat([special, nan_neg, torch.randn(n, dtype=torch.float64, device="cuda", generator=g) * 3,
                       torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 1470 - 750, _rand_exp(g, n // 4, -40, 11),
                       # multiples of ln2 to the last bits (the reduced argument's numerator test)
                       torch.arange(-1075, 1025, dtype=torch.float64, device="cuda") * 0.6931471805599453])
    probe(2, x_exp)
    x_log = torch.cat([special, nan_neg, _rand_exp(g, n, -1074, 1023), torch.rand(n, dtype=torch.float64, device="cuda", generator=g),
                       1.0 + (torch.rand(n // 4, dtype=torch.float64, device="cuda", generator=g) - 0.5) * 2.0 ** -18,
                       torch.ldexp(torch.ones(2098, dtype=torch.float64, device="cuda"), torch.arange(-1074, 1024, device="cuda").to(torch.float64))])
    probe(3, x_log)
    xs = torch.cat([special, _rand_exp(g, n, -1074, 1023, signed=False), torch.exp(torch.randn(n, dtype=torch.float64, device="cuda", generator=g) * 2),
                    torch.ldexp(torch.ones(2098, dtype=torch.float64, device="cuda"), torch.arange(-1074, 1024, device="cuda").to(torch.float64))])
    ys = torch.cat([special, torch.randn(xs.numel() - special.numel(), dtype=torch.float64, device="cuda", generator=g) * 4])
    ys[-2098::3] = -2.0  # exact powers of two with integer exponents: y log2(x) is an integer, z = 0 (numerator test)
    ys[-2097::3] = 3.0
    probe(4, xs, ys)
    probe(5, xs)  # the funnel's pow(sigma, -2.0) with the exponent as a literal


def test_row_functions_against_cuda_libm(probe):
    """rn_row_exp / rn_row_log / rn_row_rcp (total, branch-free; the streamed row bodies of the warp-per-chain kernels) against
    CUDA's exp, log and 1.0 / x: the same class for every special input, finite results within 2 units in the last place
    (both sides are < 1 ulp functions; the reciprocal is correctly rounded in the normal range: 0 there)."""
    import torch
    L = C.CDLL(build_probe())
    L.rowlibm_probe_run.restype = C.c_int
    L.rowlibm_probe_run.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]

    def run(which, a, tol):
        a = a.contiguous()
        bad = torch.zeros(1, dtype=torch.int64, device="cuda")
        worst = torch.zeros(1, dtype=torch.int64, device="cuda")
        first = torch.full((1,), 2 ** 62, dtype=torch.int64, device="cuda")
        assert L.rowlibm_probe_run(which, a.data_ptr(), a.numel(), tol, bad.data_ptr(), worst.data_ptr(), first.data_ptr()) == 0
        if int(bad.item()):
            i = int(first.item())
            raise AssertionError("which=%d: %d of %d results out of tolerance; first at %d: x=%s" % (which, int(bad.item()), a.numel(), i, a[i].item().hex()))
        return int(worst.item())

    g = torch.Generator(device="cuda").manual_seed(14)
    n = 6_000_000
    special = torch.tensor([0.0, -0.0, 1.0, -1.0, 2.0, 0.5, float("inf"), float("-inf"), float("nan"), 5e-324, -5e-324, 2.2250738585072014e-308,
                            1.7976931348623157e308, -1.7976931348623157e308, 709.782712893384, 709.7827128933841, 709.9, 1000.0, -745.1332191019411,
                            -745.1332191019412, -746.0, -1000.0, -708.3964185322641, -708.5, -740.0, 2.0 ** -1000, 2.0 ** 1000, 2.0 ** -1030, 2.0 ** 1023,
                            -2.0 ** -1030, 3.0, 1 - 2.0 ** -53, 1 + 2.0 ** -52], dtype=torch.float64, device="cuda")
    x = torch.cat([special, torch.randn(n, dtype=torch.float64, device="cuda", generator=g) * 5,
                   torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 1470 - 750, _rand_exp(g, n // 2, -60, 11)])
    w = run(0, x, 2)
    assert w <= 2, w
    x = torch.cat([special, _rand_exp(g, n, -1074, 1023, signed=False), torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2,
                   1.0 + (torch.rand(n // 2, dtype=torch.float64, device="cuda", generator=g) - 0.5) * 2.0 ** -18, -_rand_exp(g, 1000, -100, 100, signed=False)])
    w = run(1, x, 2)
    assert w <= 2, w
    x = torch.cat([special, _rand_exp(g, n, -1074, 1023), torch.randn(n, dtype=torch.float64, device="cuda", generator=g)])
    run(2, x, 1)
    assert run(2, _rand_exp(g, n, -950, 950), 0) == 0  # correctly rounded wherever no scaling is involved
