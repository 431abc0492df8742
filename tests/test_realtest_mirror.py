"""
Mirror of the reference's own op/gradient test, rainier-test/.../compute/RealTest.scala:8-204 (same expression list, same
evaluation points, same assertWithinEpsilon of ComputeTest.scala:6-17): for every expression f and point n

    constant folding f(Real(n))  ==  Evaluator  ==  compiled IR          (values)
    numeric derivative           ==  Evaluator(f')  ==  compiled f'      (derivatives, Gradient.derive)

where "compiled" is, here, (a) the oracle's interpreter of the RIR the expression freezes to and (b) for a subset the
CUDA source the emitter writes for it, compiled for the host (symbolic gradient: must equal (a) bit for bit; the emitter's
own reverse-mode adjoints: within the test's epsilon).  This pins the IR/op semantics (a4-a8 of SURVEY.md 8) on this box.
"""
import math

import numpy as np
import pytest

from oracle.rainier_py.binding import OracleModel
from oracle.rainier_py.compute import ArithmeticException, Evaluator, Real, Scalar, gradient_derive, lookup_apply, to_real
from oracle.rainier_py.core import Gamma, Model, Normal, Poisson
from rainier_b200 import abi, api

import host_emulation as he

INF, NAN = math.inf, math.nan
POINTS = [1.0, 0.0, -1.0, 2.0, -2.0, 0.5, -0.5, -INF, INF]


def within_epsilon(x, y):  # ComputeTest.assertWithinEpsilon
    if abs(x) > 10e-8 or abs(y) > 10e-8:
        if math.isnan(x) and math.isnan(y):
            return True
        if x == y:
            return True
        if x == 0 or math.isinf(x) or math.isnan(x) or math.isnan(y):
            return False
        return abs((x - y) / x) < 0.001
    return True


_EXPONENTS = list(np.random.default_rng(7).permutation(np.arange(-40, 41)))  # scala.util.Random.shuffle(-40.to(40)), fixed here


def _exponent_sums(x):
    a = x
    for e in _EXPONENTS:
        a = (a + x.pow(int(e))) * x
    return a


def _safe(fn):
    try:
        return fn()
    except (ValueError, OverflowError):
        return NAN


always, finite = (lambda n: True), (lambda n: not math.isinf(n))
# (name, fn, defined, derivable, reference)
CASES = [
    ("plus", lambda x: x + 1, always, always, None),
    ("exp", lambda x: x.exp(), always, always, None),
    ("square", lambda x: x * x, always, always, None),
    ("log", lambda x: x.abs().log(), always, always, None),
    ("sin", lambda x: x.sin(), finite, always, math.sin),
    ("cos", lambda x: x.cos(), finite, always, math.cos),
    ("tan", lambda x: x.tan(), finite, always, math.tan),
    ("asin", lambda x: x.asin(), lambda n: -1 < n < 1, always, math.asin),
    ("acos", lambda x: x.acos(), lambda n: -1 < n < 1, always, math.acos),
    ("atan", lambda x: x.atan(), always, always, math.atan),
    ("sinh", lambda x: x.sinh(), always, always, lambda n: _safe(lambda: math.sinh(n))),
    ("cosh", lambda x: x.cosh(), always, always, lambda n: _safe(lambda: math.cosh(n))),
    ("tanh", lambda x: x.tanh(), finite, always, math.tanh),
    ("tanh at infty", lambda x: x.tanh(), always, always, None),
    ("cos(x^2)", lambda x: (x * x).cos(), finite, always, None),
    ("temp", lambda x: (x * 3) + (x * 3), always, always, None),
    ("abs", lambda x: x.abs(), always, always, None),
    ("max(x, 0)", lambda x: x.max(0), always, lambda n: n != 0, None),
    ("max(x, x)", lambda x: x.max(x), always, always, None),
    ("x > 0 ? x^2 : 1", lambda x: Real.gt(x, 0, x * x, 1), always, lambda n: n != 0, None),
    ("x > 0 ? 1 : x + 1", lambda x: Real.gt(x, 0, 1, x + 1), always, lambda n: n != 0, None),
    ("x > 0 ? x^2 : x + 1", lambda x: Real.gt(x, 0, x * x, x + 1), always, lambda n: n != 0, None),
    ("normal", lambda x: Normal(x, 1).logDensity(to_real(1.0)), lambda n: n != INF, always, None),
    ("normal sum", lambda x: Real.sum([Normal(x, 1).logDensity(to_real(y)) for y in (0.0, 1.0)]), lambda n: n != INF, always, None),
    ("logistic", lambda x: ((Real.one / (Real.one + (x * -1).exp())) * (Real.one - Real.one / (Real.one + (x * -1).exp()))).log(),
     always, always, None),
    ("minimal logistic", lambda x: Real.one / (x.exp() + 1), always, always, lambda n: 1.0 / (_safe(lambda: math.exp(n)) + 1) if n != INF else 0.0),
    ("log x^2", lambda x: x.pow(2).log(), always, lambda n: n != 0, lambda n: math.log(n * n) if n != 0 else -INF),
    ("poisson", lambda x: Real.sum([Poisson(x.abs() + 1).logDensity(y) for y in range(0, 11)]), always, always, None),
    ("4x^3", lambda x: ((((x + x) * x) + (x * x)) * x) + (x * x * x), always, always, lambda n: 4 * n * n * n),
    ("lookup", lambda x: lookup_apply(x.abs() * 2, [to_real(v) for v in (0, 1, 2, 3, 4)]),
     lambda n: abs(n) <= 2 and float(abs(n) * 2).is_integer(), lambda n: False, lambda n: abs(n) * 2),
    ("exponent sums", _exponent_sums, always, lambda n: n != 0, None),
    ("cancelling x^2 then distributing", lambda x: (x.pow(2) * 2) / (x.pow(2)) + x, lambda n: n != 0 and finite(n) and float(n).is_integer(),
     always, None),
    ("pow", lambda x: x.pow(x), lambda n: n >= 0, always, None),
    ("gamma fit", lambda x: Real.sum([Gamma.standard(x.abs()).logDensity(to_real(y)) for y in (1.0, 2.0, 3.0)]), always, always, None),
]
EMITTED = {"tan", "acos", "tanh at infty", "max(x, 0)", "x > 0 ? x^2 : x + 1", "logistic", "poisson", "lookup", "exponent sums", "pow",
           "gamma fit"}


def _eval_at(fn, d):
    try:
        r = fn(to_real(d))
    except ArithmeticException:
        return NAN
    assert isinstance(r, Scalar), "Non-constant value %r" % (r,)
    return r.getDouble()


@pytest.mark.parametrize("name,fn,defined,derivable,reference", CASES, ids=[c[0] for c in CASES])
def test_real_expression(name, fn, defined, derivable, reference):
    holder = {}

    def prior(t):
        holder["x"] = t[0]
        holder["result"] = fn(t[0])
        return holder["result"]

    params = Real.parameters(1, prior)
    x, result = holder["x"], holder["result"]
    deriv = gradient_derive([x], result)[0]
    model = Model.track_(list(params))
    rir, cols = model.compile(True)
    om = OracleModel(rir, cols)
    points = [n for n in POINTS if defined(n)]
    compiled = om.density_batch(np.array([[n] for n in points]))  # [density, d/dx] per point: c and dc of RealTest
    emitted = {}
    if name in EMITTED:
        cm = api.CudaModel(rir, cols, device=-1)
        q = np.array([[n] for n in points])
        sym, err = he.density(cm.emit_source(api.make_config(sampler=api.HMCSampler(1), gradientMode=abi.RN_GRAD_SYMBOLIC)), q, cols, cm, opt="-O0")
        assert err == 0
        same = (sym == compiled) | (np.isnan(sym) & np.isnan(compiled))
        assert np.all(same), "emitted symbolic-gradient code differs from the oracle interpreter"
        prir, pcols = model.compile(False)
        pm = api.CudaModel(prir, pcols, device=-1)
        emitted["adj"], err = he.density(pm.emit_source(api.make_config(sampler=api.HMCSampler(1))), q, pcols, pm, opt="-O0")
        assert err == 0
    for k, n in enumerate(points):
        constant = _eval_at(fn, n)
        if reference is not None:
            assert within_epsilon(constant, reference(n)), "[c/ref, n=%s] %r %r" % (n, constant, reference(n))
        ev = Evaluator({x: n})
        with_var = ev.toDouble(result)
        assert within_epsilon(constant, with_var), "[c/ev, n=%s] %r %r" % (n, constant, with_var)
        assert within_epsilon(with_var, compiled[k, 0]), "[ev/ir, n=%s] %r %r" % (n, with_var, compiled[k, 0])
        if "adj" in emitted:
            assert within_epsilon(with_var, emitted["adj"][k, 0]), "[ev/cuda, n=%s]" % n
        if derivable(n) and not math.isinf(n):
            dx = 10e-6
            num_diff = (_eval_at(fn, n + dx) - _eval_at(fn, n - dx)) / (dx * 2)
            diff_with_var = ev.toDouble(deriv)
            assert within_epsilon(num_diff, diff_with_var), "[numDiff/diffWithVar, n=%s] %r %r" % (n, num_diff, diff_with_var)
            assert within_epsilon(diff_with_var, compiled[k, 1]), "[diffWithVar/diffCompiled, n=%s] %r %r" % (n, diff_with_var, compiled[k, 1])
            if "adj" in emitted:
                assert within_epsilon(diff_with_var, emitted["adj"][k, 1]), "[diffWithVar/cuda adjoint, n=%s] %r %r" % (
                    n, diff_with_var, emitted["adj"][k, 1])
