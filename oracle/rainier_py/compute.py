"""
oracle/rainier_py/compute.py -- TEST INFRASTRUCTURE ONLY (oracle tooling, never imported by the product).

Python restatement of rainier-compute's symbolic front end, used to *construct* frozen DAGs the way the
reference would, so that the CPU oracle (and the RIR blobs fed to the CUDA emitter in tests/bench) carry the
reference's exact floating-point operation order.  Cites are relative to
rainier-compute/src/main/scala/com/stripe/rainier/ (C/ = compute/, IR/ = ir/).

Restated here:
  Bounds                C/Bounds.scala:5-141
  Real node types       C/Real.scala:9-315   (equality semantics: case classes structural, others identity)
  ConstantOps           C/ConstantOps.scala:5-114
  Coefficients          C/Coefficients.scala:5-140
  RealOps               C/RealOps.scala:5-99
  LineOps / LogLineOps  C/LineOps.scala:3-97, C/LogLineOps.scala:6-91
  ToReal                C/ToReal.scala:3-41
  Gradient              C/Gradient.scala:6-153
  PartialEvaluator      C/PartialEvaluator.scala:3-98
  Target / TargetGroup  C/Target.scala:5-208
  Translator -> RIR     C/Translator.scala:5-188  (flattened to include/rainier_rir.h instead of ir.Expr)
  Evaluator             C/Evaluator.scala:3-48
  Vec                   C/Vec.scala:3-175

Where the reference iterates a hash-ordered immutable Set/Map of more than four identity-hashed elements
(C/Target.scala:73-75 priors, :28-30 gradientColumns, :84-86 columns) its own order is JVM-run dependent; this
restatement uses insertion order (what Scala's Set1..Set4/Map1..Map4 give).
"""
import math
import struct
import sys

import numpy as np

sys.setrecursionlimit(100000)

INF = float("inf")
NAN = float("nan")


# ------------------------------------------------------------------------------------------------------
# java.lang.Math on scalars (IEEE results instead of Python exceptions)
# ------------------------------------------------------------------------------------------------------
def jexp(x):
    try:
        return math.exp(x)
    except OverflowError:
        return INF


def jlog(x):
    if x != x:
        return NAN
    if x == 0.0:
        return -INF
    if x < 0.0:
        return NAN
    return math.log(x)


def jpow(x, y):
    if y == 0.0:
        return 1.0
    if y != y or x != x:
        return NAN
    if math.isinf(y) and abs(x) == 1.0:
        return NAN
    with np.errstate(all="ignore"):
        return float(np.power(np.float64(x), np.float64(y)))


def jdiv(x, y):
    if y == 0.0:
        if x != x or x == 0.0:
            return NAN
        neg = (math.copysign(1.0, x) < 0) != (math.copysign(1.0, y) < 0)
        return -INF if neg else INF
    return x / y


def jd2i(v):
    if v != v:
        return 0
    if v >= 2147483647.0:
        return 2147483647
    if v <= -2147483648.0:
        return -2147483648
    return int(v)


def is_whole(d):
    if d != d or math.isinf(d):
        return False
    return float(int(d)) == d


class ArithmeticException(ArithmeticError):
    pass


# ------------------------------------------------------------------------------------------------------
# IR ops and symbols  (IR/Ops.scala:3-37, IR/IR.scala:41-51)
# ------------------------------------------------------------------------------------------------------
ExpOp, LogOp, AbsOp, NoOp, SinOp, CosOp, TanOp, AsinOp, AcosOp, AtanOp = range(10)
AddOp, MultiplyOp, SubtractOp, DivideOp, PowOp, CompareOp = range(6)
_COMMUTATIVE = {AddOp: True, MultiplyOp: True, SubtractOp: False, DivideOp: False, PowOp: False, CompareOp: False}
UNARY_NAMES = ["exp", "log", "abs", "noop", "sin", "cos", "tan", "asin", "acos", "atan"]


class _Sym:
    """IR/IR.scala:41-51: a single global counter shared by Params (parameters AND columns) and VarDefs."""

    counter = 0

    @classmethod
    def fresh(cls):
        v = cls.counter
        cls.counter += 1
        return v


# ------------------------------------------------------------------------------------------------------
# Bounds  (C/Bounds.scala)
# ------------------------------------------------------------------------------------------------------
class Bounds:
    __slots__ = ("lower", "upper")

    def __init__(self, lower, upper):
        self.lower = float(lower)
        self.upper = float(upper)

    @property
    def is_positive(self):
        return self.lower >= 0.0

    @staticmethod
    def or_(seq):
        return Bounds(min(b.lower for b in seq), max(b.upper for b in seq))

    @staticmethod
    def sum(seq):
        lo = 0.0
        hi = 0.0
        for b in seq:
            lo += b.lower
            hi += b.upper
        return Bounds(lo, hi)

    @staticmethod
    def _mul(left, right):  # C/Bounds.scala:31-37
        if math.isinf(left) and right == 0.0:
            return left
        if left == 0.0 and math.isinf(right):
            return right
        return left * right

    @staticmethod
    def multiply(l, r):
        o = [Bounds._mul(l.lower, r.lower), Bounds._mul(l.lower, r.upper), Bounds._mul(l.upper, r.lower),
             Bounds._mul(l.upper, r.upper)]
        return Bounds(_jmin_list(o), _jmax_list(o))

    @staticmethod
    def pow(x, y):  # :39-52
        if y.lower >= 0.0:
            return Bounds._positive_pow(x, y)
        if y.upper <= 0.0:
            return Bounds._negative_pow(x, y)
        return Bounds.or_([Bounds._negative_pow(x, Bounds(y.lower, 0.0)), Bounds._positive_pow(x, Bounds(0.0, y.upper))])

    @staticmethod
    def _positive_pow(x, y):  # :54-67
        if x.lower >= 0.0:
            return Bounds._pp_pow(x, y)
        if x.upper <= 0.0:
            return Bounds._np_pow(x, y)
        return Bounds.or_([Bounds._np_pow(Bounds(x.lower, 0.0), y), Bounds._pp_pow(Bounds(0.0, x.upper), y)])

    @staticmethod
    def _negative_pow(x, y):  # :69-70
        return Bounds.reciprocal(Bounds._positive_pow(x, Bounds(y.lower * -1, y.upper * -1)))

    @staticmethod
    def _pp_pow(x, y):  # :72-80
        o = [jpow(x.lower, y.lower), jpow(x.lower, y.upper), jpow(x.upper, y.lower), jpow(x.upper, y.upper)]
        return Bounds(_jmin_list(o), _jmax_list(o))

    @staticmethod
    def _np_pow(x, y):  # :82-92
        if y.lower == y.upper and is_whole(y.lower) and abs(y.lower) <= 2147483647:
            o = [jpow(x.lower, y.lower), jpow(x.upper, y.lower)]
            return Bounds(_jmin_list(o), _jmax_list(o))
        return Bounds(-INF, INF)

    @staticmethod
    def reciprocal(x):  # :94-98
        if x.lower <= 0.0 and x.upper >= 0.0:
            return Bounds(-INF, INF)
        return Bounds(jdiv(1.0, x.upper), jdiv(1.0, x.lower))

    @staticmethod
    def abs(x):  # :100-106
        if x.lower <= 0.0 and x.upper >= 0.0:
            return Bounds(0.0, max(abs(x.lower), x.upper))
        o = [abs(x.lower), abs(x.upper)]
        return Bounds(min(o), max(o))

    @staticmethod
    def log(x):
        return Bounds(jlog(x.lower), jlog(x.upper))

    @staticmethod
    def exp(x):
        return Bounds(jexp(x.lower), jexp(x.upper))

    # :111-137
    @staticmethod
    def test(value, fn):
        return fn(value.bounds.lower) and fn(value.bounds.upper)

    @staticmethod
    def positive(value, calc):
        if Bounds.test(value, lambda v: v >= 0.0):
            return calc()
        return Real.gte(value, Real.zero, calc(), Real.negInfinity)

    @staticmethod
    def zeroToOne(value, calc):
        if Bounds.test(value, lambda v: v >= 0.0 and v <= 1.0):
            return calc()
        return Real.gte(value, Real.zero, Real.lte(value, Real.one, calc(), Real.negInfinity), Real.negInfinity)

    @staticmethod
    def check(value, description, fn):
        pass  # only logs a warning in the reference (:139-141)


def _jmin_list(o):
    # Scala List[Double].min uses Ordering.Double (compare); NaN handling irrelevant to DAG shape here
    m = o[0]
    for v in o[1:]:
        if v < m or (v != v):
            m = v
    return m


def _jmax_list(o):
    m = o[0]
    for v in o[1:]:
        if v > m or (v != v):
            m = v
    return m


# ------------------------------------------------------------------------------------------------------
# Real  (C/Real.scala)
# ------------------------------------------------------------------------------------------------------
def to_real(v):  # C/ToReal.scala:7-40
    if isinstance(v, Real):
        return v
    if isinstance(v, bool):
        raise TypeError("bool is not a Real")
    if isinstance(v, (int, np.integer)):
        return Scalar(float(v))
    d = float(v)
    if d == -INF:
        return Real.negInfinity
    if d == INF:
        return Real.infinity
    if d != d:
        raise ArithmeticException("Trying to convert NaN to Real")
    return Scalar(d)


class Real:
    __slots__ = ("bounds",)

    # operators, C/Real.scala:12-43
    def __add__(self, other):
        return RealOps.add(self, to_real(other))

    def __radd__(self, other):
        return RealOps.add(to_real(other), self)

    def __mul__(self, other):
        return RealOps.multiply(self, to_real(other))

    def __rmul__(self, other):
        return RealOps.multiply(to_real(other), self)

    def __neg__(self):
        return self * (-1)

    def __sub__(self, other):
        return self + (-to_real(other))

    def __rsub__(self, other):
        return to_real(other) + (-self)

    def __truediv__(self, other):
        return RealOps.divide(self, to_real(other))

    def __rtruediv__(self, other):
        return RealOps.divide(to_real(other), self)

    def min(self, other):
        return RealOps.min(self, to_real(other))

    def max(self, other):
        return RealOps.max(self, to_real(other))

    def pow(self, exponent):
        return RealOps.pow(self, to_real(exponent))

    def exp(self):
        return RealOps.unary(self, ExpOp)

    def log(self):
        return RealOps.unary(self, LogOp)

    def sin(self):
        return RealOps.unary(self, SinOp)

    def cos(self):
        return RealOps.unary(self, CosOp)

    def tan(self):
        return RealOps.unary(self, TanOp)

    def asin(self):
        return RealOps.unary(self, AsinOp)

    def acos(self):
        return RealOps.unary(self, AcosOp)

    def atan(self):
        return RealOps.unary(self, AtanOp)

    def sinh(self):
        return (self.exp() - (-self).exp()) / 2

    def cosh(self):
        return (self.exp() + (-self).exp()) / 2

    def tanh(self):
        return self.sinh() / self.cosh()

    def abs(self):
        return RealOps.unary(self, AbsOp)

    def logit(self):
        return -((Real.one / self - 1).log())

    def logistic(self):
        return Real.one / (Real.one + (-self).exp())

    # companion, C/Real.scala:45-108
    @staticmethod
    def sum(seq):
        acc = Real.zero
        for x in seq:
            acc = acc + x
        return acc

    @staticmethod
    def logSumExp(seq):
        seq = list(seq)
        mx = seq[0]
        for x in seq[1:]:
            mx = mx.max(x)
        shifted = [x - mx for x in seq]
        summed = Real.sum([x.exp() for x in shifted])
        return summed.log() + mx

    @staticmethod
    def parameter(fn=None):
        x = Parameter(Prior(Real.zero))
        if fn is not None:
            x.prior = Prior(fn(x))
        return x

    @staticmethod
    def parameters(size, fn):
        vector = [Parameter(Prior(Real.zero)) for _ in range(size)]
        prior = Prior(fn(vector))
        for x in vector:
            x.prior = prior
        return vector

    @staticmethod
    def doubles(seq):
        return Column(np.asarray(seq, dtype=np.float64))

    @staticmethod
    def eq(left, right, ifTrue, ifFalse):
        return Real._lookupCompare(left, right, ifFalse, ifTrue, ifFalse)

    @staticmethod
    def lt(left, right, ifTrue, ifFalse):
        return Real._lookupCompare(left, right, ifFalse, ifFalse, ifTrue)

    @staticmethod
    def gt(left, right, ifTrue, ifFalse):
        return Real._lookupCompare(left, right, ifTrue, ifFalse, ifFalse)

    @staticmethod
    def lte(left, right, ifTrue, ifFalse):
        return Real._lookupCompare(left, right, ifFalse, ifTrue, ifTrue)

    @staticmethod
    def gte(left, right, ifTrue, ifFalse):
        return Real._lookupCompare(left, right, ifTrue, ifTrue, ifFalse)

    @staticmethod
    def _lookupCompare(left, right, gt, eq, lt):
        left, right, gt, eq, lt = (to_real(v) for v in (left, right, gt, eq, lt))
        return lookup_apply(RealOps.compare(left, right), [lt, eq, gt], -1)


class Constant(Real):  # C/Real.scala:110-131
    __slots__ = ()

    @property
    def isZero(self):
        return self.bounds.lower == 0.0 and self.bounds.upper == 0.0

    @property
    def isOne(self):
        return self.bounds.lower == 1.0 and self.bounds.upper == 1.0

    @property
    def isTwo(self):
        return self.bounds.lower == 2.0 and self.bounds.upper == 2.0

    @property
    def isPosInfinity(self):
        return self.bounds.lower == INF and self.bounds.upper == INF

    @property
    def isNegInfinity(self):
        return self.bounds.lower == -INF and self.bounds.upper == -INF

    @property
    def isPositive(self):
        return self.bounds.lower >= 0.0

    def cadd(self, other):
        return ConstantOps.add(self, other)

    def cmul(self, other):
        return ConstantOps.multiply(self, other)

    def cdiv(self, other):
        return ConstantOps.divide(self, other)


class Scalar(Constant):  # case class: structural equality
    __slots__ = ("value",)

    def __init__(self, value):
        self.value = float(value)
        self.bounds = Bounds(self.value, self.value)

    def getDouble(self):
        return self.value

    def map(self, fn, vfn=None):
        return Scalar(fn(self.value))

    def mapWith(self, other, fn, vfn):
        if isinstance(other, Scalar):
            return Scalar(fn(self.value, other.value))
        with np.errstate(all="ignore"):
            return Column(vfn(self.value, other.values))

    def __eq__(self, other):
        return self is other or (isinstance(other, Scalar) and self.value == other.value)

    def __hash__(self):
        return hash(self.value)

    def __repr__(self):
        return "Scalar(%r)" % self.value


class Column(Constant):  # identity equality; owns an ir.Param
    __slots__ = ("values", "param_id")

    def __init__(self, values):
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        self.param_id = _Sym.fresh()
        self.bounds = Bounds(float(np.min(self.values)), float(np.max(self.values)))

    def getDouble(self):
        raise RuntimeError("Not a scalar")

    def map(self, fn, vfn=None):
        if vfn is not None:
            with np.errstate(all="ignore"):
                return Column(vfn(self.values))
        return Column(np.array([fn(float(v)) for v in self.values], dtype=np.float64))

    def mapWith(self, other, fn, vfn):
        with np.errstate(all="ignore"):
            if isinstance(other, Scalar):
                return Column(vfn(self.values, other.value))
            return Column(vfn(self.values, other.values))

    def maybeScalar(self):
        if self.bounds.lower == self.bounds.upper:
            return self.bounds.lower
        return None

    __hash__ = object.__hash__


class NonConstant(Real):
    __slots__ = ()


class Prior:
    __slots__ = ("density",)

    def __init__(self, density):
        self.density = density


class Parameter(NonConstant):  # identity
    __slots__ = ("prior", "param_id")

    def __init__(self, prior):
        self.prior = prior
        self.param_id = _Sym.fresh()
        self.bounds = Bounds(-INF, INF)

    __hash__ = object.__hash__


class Unary(NonConstant):  # case class
    __slots__ = ("original", "op", "_h")

    def __init__(self, original, op):
        self.original = original
        self.op = op
        ob = original.bounds
        if op == NoOp:
            self.bounds = ob
        elif op == AbsOp:
            self.bounds = Bounds.abs(ob)
        elif op == ExpOp:
            self.bounds = Bounds.exp(ob)
        elif op == LogOp:
            self.bounds = Bounds.log(ob)
        elif op in (SinOp, CosOp):
            self.bounds = Bounds(-1, 1)
        elif op == TanOp:
            self.bounds = Bounds(-INF, INF)
        else:
            self.bounds = Bounds(0, math.pi / 2.0)
        self._h = hash(("U", op, hash(original)))

    def __eq__(self, other):
        return self is other or (isinstance(other, Unary) and self.op == other.op and self.original == other.original)

    def __hash__(self):
        return self._h


class Line(NonConstant):  # identity (deliberately not a case class, C/Real.scala:199-206)
    __slots__ = ("ax", "b")

    def __init__(self, ax, b):
        assert not ax.isEmpty
        self.ax = ax
        self.b = b
        self.bounds = Bounds.sum([b.bounds] + [Bounds.multiply(x.bounds, a.bounds) for (x, a) in ax.toList()])

    __hash__ = object.__hash__


class LogLine(NonConstant):  # case class over Coefficients
    __slots__ = ("ax", "_h")

    def __init__(self, ax):
        assert not ax.isEmpty
        self.ax = ax
        bs = [Bounds.pow(x.bounds, a.bounds) for (x, a) in ax.toList()]
        b = bs[0]
        for r in bs[1:]:
            b = Bounds.multiply(b, r)
        self.bounds = b
        self._h = hash(("LL", hash(ax)))

    @staticmethod
    def of(nc):  # object LogLine.apply(nc), C/Real.scala:237-243
        if isinstance(nc, LogLine):
            return nc
        return LogLine(Coefficients.of_term(nc))

    def __eq__(self, other):
        return self is other or (isinstance(other, LogLine) and self.ax == other.ax)

    def __hash__(self):
        return self._h


class Compare(NonConstant):  # case class
    __slots__ = ("left", "right", "_h")

    def __init__(self, left, right):
        self.left = left
        self.right = right
        self.bounds = Bounds(-1, 1)
        self._h = hash(("C", hash(left), hash(right)))

    def __eq__(self, other):
        return self is other or (isinstance(other, Compare) and self.left == other.left and self.right == other.right)

    def __hash__(self):
        return self._h


class Pow(NonConstant):  # case class
    __slots__ = ("base", "exponent", "_h")

    def __init__(self, base, exponent):
        self.base = base
        self.exponent = exponent
        self.bounds = Bounds.pow(base.bounds, exponent.bounds)
        self._h = hash(("P", hash(base), hash(exponent)))

    def __eq__(self, other):
        return self is other or (isinstance(other, Pow) and self.base == other.base and self.exponent == other.exponent)

    def __hash__(self):
        return self._h


class Lookup(NonConstant):  # identity
    __slots__ = ("index", "table", "low")

    def __init__(self, index, table, low):
        self.index = index
        self.table = list(table)
        self.low = low
        self.bounds = Bounds.or_([t.bounds for t in self.table])

    __hash__ = object.__hash__


def lookup_apply(index, table, low=0):  # object Lookup.apply, C/Real.scala:287-308
    table = [to_real(t) for t in table]
    if isinstance(index, Scalar):
        return _lookup(index.value, table, low)
    if isinstance(index, Column):
        v = index.maybeScalar()
        if v is not None:
            return _lookup(v, table, low)
        if all(isinstance(t, Scalar) for t in table):
            scalars = np.array([t.value for t in table], dtype=np.float64)
            vals = index.values
            if not np.all(np.floor(vals) == vals):
                raise ArithmeticException("Cannot lookup a non-integral number")
            return Column(scalars[vals.astype(np.int64) - low])
        return Lookup(index, table, low)
    return Lookup(index, table, low)


def _lookup(index, table, low):  # C/Real.scala:310-314
    if is_whole(index):
        return table[int(index) - low]
    raise ArithmeticException("Cannot lookup a non-integral number")


# ------------------------------------------------------------------------------------------------------
# ConstantOps  (C/ConstantOps.scala)
# ------------------------------------------------------------------------------------------------------
class ConstantOps:
    @staticmethod
    def unary(original, op):
        C = Real
        if original.isPosInfinity:
            if op in (ExpOp, LogOp, AbsOp, NoOp):
                return C.infinity
            if op == AtanOp:
                return C.Pi.cdiv(C.two)
            raise ArithmeticException("no limit at +inf for op %d" % op)
        if original.isNegInfinity:
            if op == ExpOp:
                return C.zero
            if op == AbsOp:
                return C.infinity
            if op == AtanOp:
                return C.Pi.cdiv(_NegTwo)
            if op == NoOp:
                return original
            raise ArithmeticException("undefined at -inf for op %d" % op)
        if original.isZero:
            if op in (ExpOp, CosOp):
                return C.one
            if op == LogOp:
                return C.negInfinity
            if op == AcosOp:
                return C.Pi.cdiv(C.two)
            if op == NoOp:
                return original
            return C.zero
        if op == ExpOp:
            return original.map(jexp, np.exp)
        if op == LogOp:
            if not original.isPositive:
                raise ArithmeticException("Cannot take the log of a negative number")
            return original.map(jlog, np.log)
        if op == AbsOp:
            return original.map(abs, np.abs)
        if op == SinOp:
            return original.map(math.sin, np.sin)
        if op == CosOp:
            return original.map(math.cos, np.cos)
        if op == TanOp:
            return original.map(math.tan, np.tan)
        if op == AsinOp:
            return original.map(math.asin, np.arcsin)
        if op == AcosOp:
            return original.map(math.acos, np.arccos)
        if op == AtanOp:
            return original.map(math.atan, np.arctan)
        return original

    @staticmethod
    def add(left, right):
        if (left.isNegInfinity and right.isPosInfinity) or (left.isPosInfinity and right.isNegInfinity):
            raise ArithmeticException("Cannot add +inf and -inf")
        return left.mapWith(right, lambda a, b: a + b, lambda a, b: a + b)

    @staticmethod
    def multiply(left, right):
        if ((left.isPosInfinity or left.isNegInfinity) and right.isZero) or (
                left.isZero and (right.isPosInfinity or right.isNegInfinity)):
            raise ArithmeticException("Cannot multiply inf by zero")
        return left.mapWith(right, lambda a, b: a * b, lambda a, b: a * b)

    @staticmethod
    def divide(left, right):
        if left.isZero and right.isZero:
            raise ArithmeticException("Cannot divide zero by zero")
        return left.mapWith(right, jdiv, lambda a, b: np.divide(a, b))

    @staticmethod
    def pow(left, right):
        return left.mapWith(right, jpow, _vpow)

    @staticmethod
    def compare(left, right):
        def cmp(a, b):
            if a == b:
                return 0.0
            if a < b or a == -INF or b == INF:
                return -1.0
            return 1.0

        def vcmp(a, b):
            a, b = np.broadcast_arrays(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))
            out = np.ones(a.shape, dtype=np.float64)
            out[(a < b) | (a == -INF) | (b == INF)] = -1.0
            out[a == b] = 0.0
            return out

        return left.mapWith(right, cmp, vcmp)


def _vpow(a, b):
    a, b = np.broadcast_arrays(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))
    out = np.power(a, b)
    out = np.where(np.isnan(b) | np.isnan(a), np.nan, out)
    out = np.where(np.isinf(b) & (np.abs(a) == 1.0), np.nan, out)
    out = np.where(b == 0.0, 1.0, out)
    return out


# ------------------------------------------------------------------------------------------------------
# Coefficients  (C/Coefficients.scala)
# ------------------------------------------------------------------------------------------------------
class Coefficients:
    Empty = None  # set below

    @staticmethod
    def of_term(term):
        return Coefficients.of_pair(term, Real.one)

    @staticmethod
    def of_pair(term, coefficient):  # :22-26
        if coefficient.isZero:
            return Coefficients.Empty
        return COne(term, coefficient)

    @staticmethod
    def of_seq(seq):  # :28-36
        filtered = [(x, a) for (x, a) in seq if not a.isZero]
        if not filtered:
            return Coefficients.Empty
        if len(filtered) == 1:
            return Coefficients.of_pair(*filtered[0])
        m = {}
        for x, a in filtered:
            m[x] = a
        return CMany(m, [x for (x, _) in filtered])


class CEmpty(Coefficients):
    isEmpty = True
    size = 0

    def toList(self):
        return []

    def withComplements(self):
        return []

    def mapCoefficients(self, fn):
        return self

    def plus(self, term, coefficient):
        return Coefficients.of_pair(term, coefficient)

    def merge(self, other):
        return other

    def __eq__(self, other):
        return isinstance(other, CEmpty)

    def __hash__(self):
        return 0


Coefficients.Empty = CEmpty()


class COne(Coefficients):  # :53-76
    isEmpty = False
    size = 1

    def __init__(self, term, coefficient):
        self.term = term
        self.coefficient = coefficient

    def toList(self):
        return [(self.term, self.coefficient)]

    def withComplements(self):
        return [(self.term, self.coefficient, Coefficients.Empty)]

    def mapCoefficients(self, fn):
        return COne(self.term, fn(self.coefficient))

    def merge(self, other):
        return other.plus(self.term, self.coefficient)

    def plus(self, term, coefficient):
        if term == self.term:
            nc = self.coefficient.cadd(coefficient)
            if nc.isZero:
                return Coefficients.Empty
            return COne(self.term, nc)
        return Coefficients.of_seq([(term, coefficient)] + self.toList())

    def __eq__(self, other):
        return self is other or (isinstance(other, COne) and self.term == other.term and self.coefficient == other.coefficient)

    def __hash__(self):
        return hash(("1", hash(self.term), hash(self.coefficient)))


class CMany(Coefficients):  # :78-139
    isEmpty = False

    def __init__(self, toMap, terms):
        self.toMap = toMap
        self.terms = terms

    @property
    def size(self):
        return len(self.toMap)

    def toList(self):
        return [(x, self.toMap[x]) for x in self.terms]

    def mapCoefficients(self, fn):
        return CMany({x: fn(a) for (x, a) in self.toMap.items()}, self.terms)

    def withComplements(self):  # :91-112
        acc = []
        a = []
        b = list(self.terms)
        while b:
            head, tail = b[0], b[1:]
            if len(a) > len(tail):
                complementTerms = tail + a
            else:
                complementTerms = a + tail
            if len(complementTerms) == 1:
                complement = COne(complementTerms[0], self.toMap[complementTerms[0]])
            else:
                m = dict(self.toMap)
                del m[head]
                complement = CMany(m, complementTerms)
            acc.insert(0, (head, self.toMap[head], complement))
            a = [head] + a
            b = tail
        return acc

    def merge(self, other):  # :114-120
        if other.size > self.size:
            return other.merge(self)
        acc = self
        for (x, a) in other.toList():
            acc = acc.plus(x, a)
        return acc

    def plus(self, term, coefficient):  # :122-138
        if term in self.toMap:
            nc = coefficient.cadd(self.toMap[term])
            if nc.isZero:
                newMap = dict(self.toMap)
                del newMap[term]
                newTerms = [t for t in self.terms if not (t == term)]
                if len(newTerms) == 1:
                    return COne(newTerms[0], next(iter(newMap.values())))
                return CMany(newMap, newTerms)
            m = dict(self.toMap)
            m[term] = nc
            return CMany(m, self.terms)
        m = dict(self.toMap)
        m[term] = coefficient
        return CMany(m, [term] + self.terms)

    def __eq__(self, other):
        if self is other:
            return True
        if not isinstance(other, CMany):
            return False
        if len(self.terms) != len(other.terms) or len(self.toMap) != len(other.toMap):
            return False
        for x, y in zip(self.terms, other.terms):
            if not (x == y):
                return False
        for k, v in self.toMap.items():
            if k not in other.toMap or not (other.toMap[k] == v):
                return False
        return True

    def __hash__(self):
        return hash(("M", tuple(hash(t) for t in self.terms)))


# ------------------------------------------------------------------------------------------------------
# RealOps  (C/RealOps.scala)
# ------------------------------------------------------------------------------------------------------
class RealOps:
    @staticmethod
    def unary(original, op):  # :8-24
        if isinstance(original, Constant):
            return ConstantOps.unary(original, op)
        nc = original
        opt = None
        if op == ExpOp and isinstance(nc, Unary) and nc.op == LogOp:
            opt = nc.original
        elif op == AbsOp and isinstance(nc, Unary) and nc.op == AbsOp:
            opt = nc
        elif op == AbsOp and isinstance(nc, Unary) and nc.op == ExpOp:
            opt = nc
        elif op == LogOp and isinstance(nc, Unary) and nc.op == ExpOp:
            opt = nc.original
        elif op == LogOp and isinstance(nc, Line):
            opt = LineOps.log(nc)
        elif op == LogOp and isinstance(nc, LogLine):
            opt = LogLineOps.log(nc)
        if opt is not None:
            return opt
        return Unary(nc, op)

    @staticmethod
    def add(left, right):  # :26-39
        lc, rc = isinstance(left, Constant), isinstance(right, Constant)
        if lc and rc:
            return ConstantOps.add(left, right)
        if left == Real.infinity:
            return left
        if right == Real.infinity:
            return right
        if left == Real.negInfinity:
            return left
        if right == Real.negInfinity:
            return right
        if right == Real.zero:
            return left
        if left == Real.zero:
            return right
        if lc:
            return LineOps.translate(right, left)
        if rc:
            return LineOps.translate(left, right)
        return LineOps.sum(left, right)

    @staticmethod
    def multiply(left, right):  # :41-55
        lc, rc = isinstance(left, Constant), isinstance(right, Constant)
        if lc and rc:
            return ConstantOps.multiply(left, right)
        if left == Real.infinity:
            return Real.gt(right, 0, Real.infinity, Real.negInfinity)
        if right == Real.infinity:
            return Real.gt(left, 0, Real.infinity, Real.negInfinity)
        if left == Real.negInfinity:
            return Real.gt(right, Real.zero, Real.negInfinity, Real.infinity)
        if right == Real.negInfinity:
            return Real.gt(left, Real.zero, Real.negInfinity, Real.infinity)
        if right == Real.zero:
            return Real.zero
        if left == Real.zero:
            return Real.zero
        if right == Real.one:
            return left
        if left == Real.one:
            return right
        if lc:
            return LineOps.scale(right, left)
        if rc:
            return LineOps.scale(left, right)
        return LogLineOps.multiply(LogLine.of(left), LogLine.of(right))

    @staticmethod
    def divide(left, right):  # :57-62
        if isinstance(left, Constant) and isinstance(right, Constant):
            return ConstantOps.divide(left, right)
        if right == Real.zero:
            return left * Real.infinity
        return left * right.pow(-1)

    @staticmethod
    def min(left, right):
        return Real.lt(left, right, left, right)

    @staticmethod
    def max(left, right):
        return Real.gt(left, right, left, right)

    @staticmethod
    def pow(original, exponent):  # :70-89
        if not isinstance(exponent, Constant):
            return Pow(original, exponent)
        if isinstance(original, Constant):
            return ConstantOps.pow(original, exponent)
        if exponent == Real.infinity:
            return Real.infinity
        if exponent == Real.negInfinity:
            return Real.zero
        if exponent == Real.zero:
            return Real.one
        if exponent == Real.one:
            return original
        if isinstance(original, Line):
            r = LineOps.pow(original, exponent)
            if r is not None:
                return r
            return LogLineOps.pow(LogLine.of(original), exponent)
        return LogLineOps.pow(LogLine.of(original), exponent)

    @staticmethod
    def compare(left, right):  # :91-99
        if isinstance(left, Constant) and isinstance(right, Constant):
            return ConstantOps.compare(left, right)
        if left == Real.infinity:
            return Real.one
        if right == Real.infinity:
            return Real.negOne
        if left == Real.negInfinity:
            return Real.negOne
        if right == Real.negInfinity:
            return Real.one
        return Compare(left, right)


# ------------------------------------------------------------------------------------------------------
# LineOps / LogLineOps
# ------------------------------------------------------------------------------------------------------
class LineOps:
    @staticmethod
    def axb(nc):  # C/LineOps.scala:5-13
        if isinstance(nc, Line):
            return nc.ax, nc.b
        if isinstance(nc, LogLine):
            d = LogLineOps.distribute(nc)
            if d is not None:
                return d
            return Coefficients.of_term(nc), Real.zero
        return Coefficients.of_term(nc), Real.zero

    @staticmethod
    def sum(left, right):  # :15-24
        lax, lb = LineOps.axb(left)
        rax, rb = LineOps.axb(right)
        merged = lax.merge(rax)
        if merged.isEmpty:
            return lb.cadd(rb)
        return LineOps.simplify(merged, lb.cadd(rb))

    @staticmethod
    def scale(nc, v):  # :26-29
        ax, b = LineOps.axb(nc)
        return LineOps.simplify(ax.mapCoefficients(lambda a: a.cmul(v)), b.cmul(v))

    @staticmethod
    def translate(nc, v):  # :31-34
        ax, b = LineOps.axb(nc)
        return LineOps.simplify(ax, b.cadd(v))

    @staticmethod
    def multiply(left, right):  # :39-57
        allLeft = [(Real.one, left.b)] + left.ax.toList()
        allRight = [(Real.one, right.b)] + right.ax.toList()
        terms = []
        for (x, a) in allLeft:
            for (y, c) in allRight:
                terms.append((x * y, a.cmul(c)))
        nAx = Coefficients.Empty
        nB = Real.zero
        for (x, a) in terms:
            if isinstance(x, NonConstant):
                nAx = nAx.merge(Coefficients.of_pair(x, a))
            else:
                nB = nB.cadd(x.cmul(a))
        return Line(nAx, nB)

    @staticmethod
    def log(line):  # :68-73
        if isinstance(line.ax, COne) and line.ax.coefficient.isPositive and line.b.isZero:
            return line.ax.term.log() + line.ax.coefficient.log()
        return None

    @staticmethod
    def pow(line, exponent):  # :83-88
        if isinstance(line.ax, COne) and line.b.isZero:
            return line.ax.term.pow(exponent) * RealOps.pow(line.ax.coefficient, exponent)
        return None

    @staticmethod
    def simplify(ax, b):  # :90-96
        if ax.isEmpty:
            return b
        if isinstance(ax, COne) and ax.coefficient.isOne and b.isZero:
            return ax.term
        return Line(ax, b)


class LogLineOps:
    DistributeToMaxTerms = 20

    @staticmethod
    def multiply(left, right):  # C/LogLineOps.scala:7-13
        merged = left.ax.merge(right.ax)
        if merged.isEmpty:
            return Real.one
        return LogLine(merged)

    @staticmethod
    def pow(line, v):  # :15-16
        return LogLine(line.ax.mapCoefficients(lambda a: a.cmul(v)))

    @staticmethod
    def log(line):  # :25
        return None

    @staticmethod
    def distribute(line):  # :43-91
        MAX = LogLineOps.DistributeToMaxTerms

        def nTerms(l):
            return l.ax.size if l.b.isZero else l.ax.size + 1

        def nTerms2(l):
            n = nTerms(l)
            return (n * (n + 1)) // 2

        factors = []
        terms = None
        for (l, c) in line.ax.toList():
            if isinstance(l, Line):
                if terms is None and c.isOne and nTerms(l) < MAX:
                    terms = l
                    continue
                if terms is not None and c.isOne and (nTerms(terms) * nTerms(l)) < MAX:
                    terms = LineOps.multiply(terms, l)
                    continue
                if terms is None and c.isTwo and nTerms2(l) < MAX:
                    terms = LineOps.multiply(l, l)
                    continue
                if terms is not None and c.isTwo and (nTerms(terms) * nTerms2(l)) < MAX:
                    terms = LineOps.multiply(terms, LineOps.multiply(l, l))
                    continue
            factors.insert(0, (l, c))
        if terms is None:
            return None
        l = terms
        if not factors:
            return l.ax, l.b
        ll = LogLine(Coefficients.of_seq(factors))
        nAx = Coefficients.of_pair(ll, l.b)
        nB = Real.zero
        for (x, a) in l.ax.toList():
            m = LogLineOps.multiply(ll, LogLine.of(x))
            if isinstance(m, Constant):
                nB = nB.cadd(m.cmul(a))
            else:
                nAx = nAx.merge(Coefficients.of_pair(m, a))
        return nAx, nB


# constants, C/Real.scala:133-142
Real.zero = Scalar(0.0)
Real.one = Scalar(1.0)
Real.two = Scalar(2.0)
Real.negOne = Scalar(-1.0)
_NegTwo = Scalar(-2.0)
Real.Pi = Scalar(math.pi)
Real.infinity = Scalar(INF)
Real.negInfinity = Scalar(-INF)


# ------------------------------------------------------------------------------------------------------
# Gradient  (C/Gradient.scala)
# ------------------------------------------------------------------------------------------------------
class _CompoundDiff:
    def __init__(self):
        self.parts = []
        self._real = None

    def register(self, part):
        self.parts.insert(0, part)

    def toReal(self):
        if self._real is None:
            if len(self.parts) == 1:
                self._real = self.parts[0].toReal()
            else:
                self._real = Real.sum([p.toReal() for p in self.parts])
        return self._real


class _ConstDiff:
    def toReal(self):
        return Real.one


class _ProductDiff:  # :89-92
    def __init__(self, other, gradient):
        self.other, self.gradient = other, gradient

    def toReal(self):
        return self.gradient.toReal() * self.other


class _UnaryDiff:  # :94-115
    def __init__(self, child, gradient):
        self.child, self.gradient = child, gradient

    def toReal(self):
        child, g = self.child, self.gradient
        op = child.op
        if op == LogOp:
            return g.toReal() * (Real.one / child.original)
        if op == ExpOp:
            return g.toReal() * child
        if op == AbsOp:
            return Real.eq(child.original, Real.zero, Real.zero, g.toReal() * child.original / child)
        if op == NoOp:
            return g.toReal()
        if op == SinOp:
            return g.toReal() * child.original.cos()
        if op == CosOp:
            return g.toReal() * (Real.zero - child.original.sin())
        if op == TanOp:
            return g.toReal() / child.original.cos().pow(2)
        if op == AsinOp:
            return g.toReal() / (Real.one - child.original.pow(2)).pow(0.5)
        if op == AcosOp:
            return -g.toReal() / (Real.one - child.original.pow(2)).pow(0.5)
        if op == AtanOp:
            return g.toReal() / (Real.one + child.original.pow(2))
        raise AssertionError


class _PowDiff:  # :117-128
    def __init__(self, child, gradient, isExponent):
        self.child, self.gradient, self.isExponent = child, gradient, isExponent

    def toReal(self):
        child = self.child
        if self.isExponent:
            return self.gradient.toReal() * child * Real.eq(child.base, Real.zero, Real.one, child.base).log()
        return self.gradient.toReal() * child.exponent * child.base.pow(child.exponent - 1)


class _LogLineDiff:  # :130-146
    def __init__(self, gradient, term, exponent, complement):
        self.gradient, self.term, self.exponent, self.complement = gradient, term, exponent, complement

    def toReal(self):
        otherTerms = Real.one if self.complement.isEmpty else LogLine(self.complement)
        return self.gradient.toReal() * self.exponent * self.term.pow(self.exponent - Real.one) * otherTerms


class _LookupDiff:  # :148-152
    def __init__(self, child, gradient, index):
        self.child, self.gradient, self.index = child, gradient, index

    def toReal(self):
        return Real.eq(self.child.index, self.index, self.gradient.toReal(), Real.zero)


def gradient_derive(parameters, output):  # C/Gradient.scala:8-69
    diffs = {}

    def diff(real):
        d = diffs.get(real)
        if d is None:
            d = _CompoundDiff()
            diffs[real] = d
        return d

    diff(output).register(_ConstDiff())
    visited = set()

    def visit(real):
        if real in visited:
            return
        visited.add(real)
        if isinstance(real, (Parameter, Constant)):
            return
        if isinstance(real, Pow):
            diff(real.base).register(_PowDiff(real, diff(real), False))
            diff(real.exponent).register(_PowDiff(real, diff(real), True))
            visit(real.base)
            visit(real.exponent)
        elif isinstance(real, Unary):
            diff(real.original).register(_UnaryDiff(real, diff(real)))
            visit(real.original)
        elif isinstance(real, Line):
            for (x, a) in real.ax.toList():
                diff(x).register(_ProductDiff(a, diff(real)))
                visit(x)
        elif isinstance(real, LogLine):
            for (x, a, c) in real.ax.withComplements():
                diff(x).register(_LogLineDiff(diff(real), x, a, c))
                visit(x)
        elif isinstance(real, Lookup):
            for i, x in enumerate(real.table):
                diff(x).register(_LookupDiff(real, diff(real), i + real.low))
                visit(x)
            visit(real.index)
        elif isinstance(real, Compare):
            visit(real.left)
            visit(real.right)

    visit(output)
    return [diff(v).toReal() for v in parameters]


# ------------------------------------------------------------------------------------------------------
# PartialEvaluator  (C/PartialEvaluator.scala)
# ------------------------------------------------------------------------------------------------------
class PartialEvaluator:
    def __init__(self, noChange, rowIndex):
        self.noChange = noChange  # shared, mutated set (the Scala var holds an immutable Set; `next()` copies the reference)
        self.rowIndex = rowIndex
        self.cache = {}

    def next(self):
        return PartialEvaluator(set(self.noChange), self.rowIndex + 1)

    def apply(self, real):
        if real in self.noChange:
            return real, False
        if real in self.cache:
            return self.cache[real], True
        v, changed = self.eval(real)
        if changed:
            self.cache[real] = v
        else:
            self.noChange.add(real)
        return v, changed

    def eval(self, real):
        if isinstance(real, Scalar):
            return real, False
        if isinstance(real, Column):
            return to_real(float(real.values[self.rowIndex])), True
        if isinstance(real, Line):
            terms = [(self.apply(x), self.apply(a)) for (x, a) in real.ax.toList()]
            b, bModified = self.apply(real.b)
            anyModified = any(m1 or m2 for ((_, m1), (_, m2)) in terms) or bModified
            if anyModified:
                s = Real.sum([x * a for ((x, _), (a, _)) in terms])
                return s + b, True
            return real, False
        if isinstance(real, LogLine):
            terms = [(self.apply(x), self.apply(a)) for (x, a) in real.ax.toList()]
            anyModified = any(m1 or m2 for ((_, m1), (_, m2)) in terms)
            if anyModified:
                ps = [x.pow(a) for ((x, _), (a, _)) in terms]
                product = ps[0]
                for p in ps[1:]:
                    product = product * p
                return product, True
            return real, False
        if isinstance(real, Unary):
            r, modified = self.apply(real.original)
            if modified:
                return RealOps.unary(r, real.op), True
            return real, False
        if isinstance(real, Compare):
            nl, lm = self.apply(real.left)
            nr, rm = self.apply(real.right)
            if lm or rm:
                return RealOps.compare(nl, nr), True
            return real, False
        if isinstance(real, Pow):
            nb, bm = self.apply(real.base)
            ne, em = self.apply(real.exponent)
            if bm or em:
                return nb.pow(ne), True
            return real, False
        if isinstance(real, Lookup):
            ni, im = self.apply(real.index)
            nt = [self.apply(t) for t in real.table]
            anyModified = any(m for (_, m) in nt)
            if im or anyModified:
                return lookup_apply(ni, [t for (t, _) in nt], real.low), True
            return real, False
        if isinstance(real, Parameter):
            return real, False
        raise AssertionError

    @staticmethod
    def inline(real, nRows):  # :90-97
        acc = Real.zero
        pe = PartialEvaluator(set(), 0)
        for _ in range(nRows):
            acc = acc + pe.apply(real)[0]
            pe = pe.next()
        return acc


# ------------------------------------------------------------------------------------------------------
# Target / TargetGroup  (C/Target.scala)
# ------------------------------------------------------------------------------------------------------
def _leaves(real):  # :87-129  (returns ordered, de-duplicated lists in first-seen order)
    seen = set()
    params, cols = [], []

    def loop(r):
        if r in seen:
            return
        seen.add(r)
        if isinstance(r, Scalar):
            return
        if isinstance(r, Column):
            cols.append(r)
        elif isinstance(r, Parameter):
            params.append(r)
            loop(r.prior.density)
        elif isinstance(r, Unary):
            loop(r.original)
        elif isinstance(r, Line):
            for (x, a) in r.ax.toList():
                loop(x)
                loop(a)
            loop(r.b)
        elif isinstance(r, LogLine):
            for (x, a) in r.ax.toList():
                loop(x)
                loop(a)
        elif isinstance(r, Compare):
            loop(r.left)
            loop(r.right)
        elif isinstance(r, Pow):
            loop(r.base)
            loop(r.exponent)
        elif isinstance(r, Lookup):
            loop(r.index)
            for t in r.table:
                loop(t)

    loop(real)
    # the reference prepends (`leaves = x :: leaves`) then takes `.toSet`; order is layout-only
    return params[::-1], cols[::-1]


def find_parameters(real):
    return _leaves(real)[0]


def find_columns(real):
    return _leaves(real)[1]


def inlinable(real):  # :136-207
    seen = {}

    class State:
        __slots__ = ("hasParameter", "hasPlaceholder", "nonlinearCombination")

        def __init__(self, a, b, c):
            self.hasParameter, self.hasPlaceholder, self.nonlinearCombination = a, b, c

        def or_(self, o):
            return State(self.hasParameter or o.hasParameter, self.hasPlaceholder or o.hasPlaceholder,
                         self.nonlinearCombination or o.nonlinearCombination)

        @property
        def combination(self):
            return self.hasParameter and self.hasPlaceholder

        def nonlinearOp(self):
            return State(self.hasParameter, self.hasPlaceholder, self.combination)

    def loopMerge(rs):
        st = [loop(r) for r in rs]
        s = st[0]
        for t in st[1:]:
            s = s.or_(t)
        return s

    def loop(r):
        if r in seen:
            return seen[r]
        if isinstance(r, Scalar):
            result = State(False, False, False)
        elif isinstance(r, Column):
            result = State(False, True, False)
        elif isinstance(r, Parameter):
            result = State(True, False, False)
        elif isinstance(r, Unary):
            result = loopMerge([r.original]).nonlinearOp()
        elif isinstance(r, Line):
            items = [r.b]
            for (x, a) in r.ax.toList():
                items += [x, a]
            result = loopMerge(items)
        elif isinstance(r, LogLine):
            termStates = [loopMerge([x, a]).nonlinearOp() for (x, a) in r.ax.toList()]
            state = termStates[0]
            for t in termStates[1:]:
                state = state.or_(t)
            if state.nonlinearCombination or not state.combination:
                result = state
            elif any(t.combination for t in termStates):
                result = state.nonlinearOp()
            else:
                result = state
        elif isinstance(r, Compare):
            result = loopMerge([r.left, r.right]).nonlinearOp()
        elif isinstance(r, Pow):
            result = loopMerge([r.base, r.exponent]).nonlinearOp()
        elif isinstance(r, Lookup):
            tableState = loopMerge(r.table)
            indexState = loop(r.index)
            state = tableState.or_(indexState)
            result = state.nonlinearOp() if indexState.hasParameter else state
        else:
            raise AssertionError
        seen[r] = result
        return result

    return not loop(real + real).nonlinearCombination


class Target:  # :5-33
    def __init__(self, name, real, parameters, with_gradient=True):
        columns = find_columns(real)
        nRows = 0 if not columns else len(columns[0].values)
        if nRows > 0 and inlinable(real):
            real2, columns2 = PartialEvaluator.inline(real, nRows), []
        else:
            real2, columns2 = real, columns
        if parameters and with_gradient:
            gradient = gradient_derive(parameters, real2)
        else:
            gradient = []
        gcols = []
        seen_g = set()
        for g in gradient:
            if g in seen_g:
                continue
            seen_g.add(g)
            for c in find_columns(g):
                if c not in columns2 and c not in gcols:
                    gcols.append(c)
        self.name = name
        self.real = real2
        self.columns = columns2
        self.gradient = gradient
        self.gradientColumns = gcols
        self.nRows = len(columns2[0].values) if columns2 else (len(gcols[0].values) if gcols else 0)


class TargetGroup:  # :36-80
    def __init__(self, reals, track=(), with_gradient=True):
        pset = []
        seenp = set()
        allr = []
        for r in list(reals) + list(track):
            if r not in allr:
                allr.append(r)
        for r in allr:
            for p in find_parameters(r):
                if p not in seenp:
                    seenp.add(p)
                    pset.append(p)
        parameters = sorted(pset, key=lambda p: p.param_id)
        priors = []
        for p in parameters:
            if not any(p.prior is q for q in priors):
                priors.append(p.prior)
        densities = []
        for pr in priors:
            if pr.density not in densities:
                densities.append(pr.density)
        priorTarget = Target("prior", Real.sum(densities), parameters, with_gradient)
        others = [Target("t_%d" % i, r, parameters, with_gradient) for i, r in enumerate(reals)]
        self.targets = [priorTarget] + others
        self.parameters = parameters
        self.with_gradient = with_gradient

    @property
    def columns(self):
        out = []
        for t in self.targets:
            out += t.columns + t.gradientColumns
        return out

    @property
    def outputs(self):
        out = []
        for t in self.targets:
            out.append((t.name, t.real))
            for i, g in enumerate(t.gradient):
                out.append(("%s_grad_%d" % (t.name, i), g))
        return out


# ------------------------------------------------------------------------------------------------------
# Translator -> RIR  (C/Translator.scala; flat SSA per include/rainier_rir.h)
# ------------------------------------------------------------------------------------------------------
RIR_INPUT, RIR_CONST, RIR_UNARY, RIR_BINARY, RIR_LOOKUP = range(5)
RIR_FLAG_GRADIENT = 1


class Translator:
    """Refs are tuples: ('p', param_id) | ('c', float) | ('v', sym).  A fresh VarDef is ('d', sym)."""

    def __init__(self, input_index):
        self.input_index = input_index  # param_id -> input position
        self.binary = {}
        self.unary = {}
        self.reals = {}
        self.nodes = []          # (kind, op, a, b, c, d, value)
        self.lookup_refs = []
        self.node_of_sym = {}    # sym -> node id
        self.node_of_leaf = {}   # ('p', id) / ('c', bits) -> node id

    # --- node plumbing -------------------------------------------------------------------------------
    def _emit(self, rec):
        self.nodes.append(rec)
        return len(self.nodes) - 1

    def node_id(self, ref):
        k = ref[0]
        if k in ("v", "d"):
            return self.node_of_sym[ref[1]]
        if k == "p":
            if ref not in self.node_of_leaf:
                self.node_of_leaf[ref] = self._emit((RIR_INPUT, 0, self.input_index[ref[1]], 0, 0, 0, 0.0))
            return self.node_of_leaf[ref]
        key = ("c", struct.pack("<d", ref[1]))
        if key not in self.node_of_leaf:
            self.node_of_leaf[key] = self._emit((RIR_CONST, 0, 0, 0, 0, 0, ref[1]))
        return self.node_of_leaf[key]

    @staticmethod
    def ref(expr):  # :159-163
        if expr[0] == "d":
            return ("v", expr[1])
        return expr

    # --- toExpr, :10-27 ---------------------------------------------------------------------------------
    def toExpr(self, r):
        hit = self.reals.get(r)
        if hit is not None:
            return self.ref(hit)
        if isinstance(r, Parameter):
            expr = ("p", r.param_id)
        elif isinstance(r, Constant):
            expr = self.constToExpr(r)
        elif isinstance(r, Unary):
            expr = self.unaryExpr(self.toExpr(r.original), r.op)
        elif isinstance(r, Line):
            expr = self.makeLine(r.ax, r.b, True)
        elif isinstance(r, LogLine):
            expr = self.makeLine(r.ax, Real.one, False)
        elif isinstance(r, Pow):
            expr = self.binaryExpr(self.toExpr(r.base), self.toExpr(r.exponent), PowOp)
        elif isinstance(r, Compare):
            expr = self.binaryExpr(self.toExpr(r.left), self.toExpr(r.right), CompareOp)
        elif isinstance(r, Lookup):
            expr = self.lookupExpr(r)
        else:
            raise AssertionError
        self.reals[r] = expr
        return expr

    def constToExpr(self, c):  # :29-36
        if isinstance(c, Scalar):
            return ("c", c.value)
        v = c.maybeScalar()
        if v is not None:
            return ("c", v)
        return ("p", c.param_id)

    @staticmethod
    def _refkey(ref):
        # Const is a case class over Double: Const(0.0) == Const(-0.0)
        if ref[0] == "c":
            return ("c", ref[1])
        return ("v" if ref[0] == "d" else ref[0], ref[1])

    def _memoize(self, cache, exprKeys, opKey, make):  # SymCache.memoize, :166-187
        refKeys = [tuple(self._refkey(e) for e in l) for l in exprKeys]
        hit = None
        for k in refKeys:
            if hit is None:
                hit = cache.get((k, opKey))
        if hit is not None:
            if any(e[0] == "d" for e in exprKeys[0]):
                raise RuntimeError("VarRef was used before its VarDef")
            return ("v", hit)
        sym = _Sym.fresh()
        cache[(refKeys[0], opKey)] = sym
        self.node_of_sym[sym] = make()
        return ("d", sym)

    def unaryExpr(self, original, op):  # :38-39
        return self._memoize(self.unary, [[original]], op,
                             lambda: self._emit((RIR_UNARY, op, self.node_id(original), 0, 0, 0, 0.0)))

    def binaryExpr(self, left, right, op):  # :41-49 (the key/commutativity quirk is the reference's)
        key = [left, right]
        keys = [key] if _COMMUTATIVE[op] else [key, key[::-1]]
        rir_op = {AddOp: 0, MultiplyOp: 1, SubtractOp: 2, DivideOp: 3, PowOp: 4, CompareOp: 5}[op]
        return self._memoize(self.binary, keys, op,
                             lambda: self._emit((RIR_BINARY, rir_op, self.node_id(left), self.node_id(right), 0, 0, 0.0)))

    def lookupExpr(self, lookup):  # :51-61
        tableExprs = [self.toExpr(t) for t in lookup.table]
        index = self.toExpr(lookup.index)
        refs = [self.ref(e) for e in tableExprs]
        sym = _Sym.fresh()
        off = len(self.lookup_refs)
        ids = [self.node_id(r) for r in refs]
        idx_node = self.node_id(index)
        self.lookup_refs += ids
        self.node_of_sym[sym] = self._emit((RIR_LOOKUP, 0, idx_node, off, len(ids), lookup.low, 0.0))
        # SeqIR(defs :+ lookupExpr): evaluation order == emission order in the flat form
        return ("d", sym)

    def makeLine(self, ax, b, is_sum):  # :91-141
        terms = [(x, self.constToExpr(a)) for (x, a) in ax.toList()]
        allTerms = terms if b.isZero else [(b, ("c", 1.0))] + terms
        plus = AddOp if is_sum else MultiplyOp
        times = MultiplyOp if is_sum else PowOp

        def lazy(x, a):
            if a[0] == "c" and a[1] == 1.0:
                return lambda: self.toExpr(x)
            if a[0] == "c" and a[1] == 2.0:
                return lambda: self.binaryExpr(self.toExpr(x), self.toExpr(x), plus)
            return lambda: self.binaryExpr(self.toExpr(x), a, times)

        lazyExprs = [lazy(x, a) for (x, a) in allTerms]
        if not is_sum:
            return self.combineTree(lazyExprs, plus)
        accum = lazyExprs[0]()
        for t in lazyExprs[1:]:
            accum = self.binaryExpr(accum, t(), plus)
        return accum

    def combineTree(self, terms, plus):  # :143-157
        while len(terms) != 1:
            grouped = []
            for i in range(0, len(terms), 2):
                if i + 1 < len(terms):
                    l, r = terms[i], terms[i + 1]
                    grouped.append((lambda l=l, r=r: self.binaryExpr(l(), r(), plus)))
                else:
                    grouped.append(terms[i])
            terms = grouped
        return terms[0]()


def compile_rir(group):
    """Compiler.compileTargets (C/Compiler.scala:14-30) with the bytecode emitter replaced by RIR serialisation.
    Returns (rir_bytes, columns) where columns is the list of float64 arrays in input order."""
    params = group.parameters
    n = len(params)
    input_index = {p.param_id: i for i, p in enumerate(params)}
    cols = []
    tmeta = []
    pos = n
    for t in group.targets:
        tc = t.columns + t.gradientColumns
        first = pos
        for c in tc:
            input_index[c.param_id] = pos
            cols.append(c.values)
            pos += 1
        tmeta.append((first, len(tc), len(tc[0].values) if tc else 0))
    tr = Translator(input_index)
    out_nodes = []
    for t in group.targets:
        outs = [t.real] + list(t.gradient)
        ids = []
        for r in outs:
            e = tr.toExpr(r)
            ids.append(tr.node_id(e))
        out_nodes.append(ids)
    flags = RIR_FLAG_GRADIENT if group.with_gradient else 0
    blob = bytearray()
    blob += struct.pack("<8I", 0x31524952, 1, n, pos, len(tr.nodes), len(group.targets), len(tr.lookup_refs), flags)
    for (kind, op, a, b, c, d, value) in tr.nodes:
        blob += struct.pack("<BBHiiiiid", kind, op, 0, a, b, c, d, 0, value)
    lr = struct.pack("<%di" % len(tr.lookup_refs), *tr.lookup_refs)
    blob += lr + b"\0" * ((-len(lr)) % 8)
    for (first, ncols, nrows), ids in zip(tmeta, out_nodes):
        blob += struct.pack("<QIIII", nrows, first, ncols, len(ids), 0)
        ob = struct.pack("<%dI" % len(ids), *ids)
        blob += ob + b"\0" * ((-len(ob)) % 8)
    return bytes(blob), cols


RIR_FLAG_FUNCTION = 2


def compile_function_rir(parameters, outputs):
    """Compiler.compile(inputs: Seq[ir.Param], outputs: Seq[(String, Real)]) (C/Compiler.scala:22-30) with the bytecode
    emitter replaced by RIR serialisation (RIR_FLAG_FUNCTION container, include/rainier_rir.h).  parameters: the model's
    Parameter list (input order); outputs: list of Real -- one Translator over all of them, in order, exactly like the
    `outputs.map { case (s, r) => s -> translator.toExpr(r) }` of :25-28."""
    n = len(parameters)
    input_index = {p.param_id: i for i, p in enumerate(parameters)}
    tr = Translator(input_index)
    ids = []
    for r in outputs:
        e = tr.toExpr(to_real(r))
        ids.append(tr.node_id(e))
    blob = bytearray()
    blob += struct.pack("<8I", 0x31524952, 1, n, n, len(tr.nodes), 1, len(tr.lookup_refs), RIR_FLAG_FUNCTION)
    for (kind, op, a, b, c, d, value) in tr.nodes:
        blob += struct.pack("<BBHiiiiid", kind, op, 0, a, b, c, d, 0, value)
    lr = struct.pack("<%di" % len(tr.lookup_refs), *tr.lookup_refs)
    blob += lr + b"\0" * ((-len(lr)) % 8)
    blob += struct.pack("<QIIII", 0, n, 0, len(ids), 0)
    ob = struct.pack("<%dI" % len(ids), *ids)
    blob += ob + b"\0" * ((-len(ob)) % 8)
    return bytes(blob)


# ------------------------------------------------------------------------------------------------------
# Evaluator  (C/Evaluator.scala)
# ------------------------------------------------------------------------------------------------------
class Evaluator:
    def __init__(self, cache=None):
        self.cache = dict(cache or {})

    def toDouble(self, x):
        x = to_real(x)
        if isinstance(x, Constant):
            return x.getDouble()
        if x in self.cache:
            return self.cache[x]
        v = self._eval(x)
        self.cache[x] = v
        return v

    def toInt(self, x):
        return jd2i(self.toDouble(x))

    def toLong(self, x):
        return int(self.toDouble(x))

    def _eval(self, real):
        if isinstance(real, Constant):
            return real.getDouble()
        if isinstance(real, Line):
            s = 0.0
            for (r, d) in real.ax.toList():
                s += self.toDouble(r) * d.getDouble()
            return s + real.b.getDouble()
        if isinstance(real, LogLine):
            p = 1.0
            for (r, d) in real.ax.toList():
                p *= jpow(self.toDouble(r), d.getDouble())
            return p
        if isinstance(real, Unary):
            ev = to_real(self.toDouble(real.original))
            return self._eval(RealOps.unary(ev, real.op))
        if isinstance(real, Compare):
            return self._eval(RealOps.compare(to_real(self.toDouble(real.left)), to_real(self.toDouble(real.right))))
        if isinstance(real, Pow):
            return jpow(self.toDouble(real.base), self.toDouble(real.exponent))
        if isinstance(real, Lookup):
            return self.toDouble(real.table[jd2i(self.toDouble(real.index)) - real.low])
        raise RuntimeError("No value provided for parameter")


# ------------------------------------------------------------------------------------------------------
# Vec  (C/Vec.scala)
# ------------------------------------------------------------------------------------------------------
class Vec:
    def take(self, k):
        return self.mapLeaves(lambda r: RealVec(r.reals[:k]))

    def drop(self, k):
        return self.mapLeaves(lambda r: RealVec(r.reals[k:]))

    def slice(self, frm, until):
        return self.mapLeaves(lambda r: RealVec(r.reals[frm:until]))

    def map(self, fn):
        return MapVec(self, fn)

    def zip(self, other):
        assert self.size == other.size
        return ZipVec(self, other)

    def toList(self):
        return [self.at(i) for i in range(self.size)]

    def columnize(self):  # :37-38
        return self.at(Column(np.arange(self.size, dtype=np.float64)))

    def dot(self, other):  # :40-43
        return Real.sum([self.at(i) * other.at(i) for i in range(self.size)])

    @staticmethod
    def of(*seq):
        return Vec.from_(list(seq))

    @staticmethod
    def from_(seq):  # ToVec instances, :97-175
        seq = list(seq)
        head = seq[0]
        if isinstance(head, Vec):
            raise TypeError("Vec.from of Vecs is not defined in the reference")
        if isinstance(head, tuple):
            k = len(head)
            vs = [Vec.from_([s[i] for s in seq]) for i in range(k)]
            z = vs[0]
            for v in vs[1:]:
                z = z.zip(v)
            if k == 2:
                return z
            if k == 3:
                return z.map(lambda t: (t[0][0], t[0][1], t[1]))
            if k == 4:
                return z.map(lambda t: (t[0][0][0], t[0][0][1], t[0][1], t[1]))
            raise TypeError("tuple arity")
        if isinstance(head, dict):
            keys = list(head.keys())
            valueVecs = [Vec.from_([m[k] for m in seq]) for k in keys]
            return TraverseVec(valueVecs).map(lambda us: dict(zip(keys, us)))
        if isinstance(head, (list, np.ndarray)):
            size = len(head)
            valueVecs = [Vec.from_([m[k] for m in seq]) for k in range(size)]
            return TraverseVec(valueVecs).map(lambda s: Vec.from_(s))
        return RealVec([to_real(s) for s in seq])


class RealVec(Vec):
    def __init__(self, reals):
        self.reals = list(reals)
        self.size = len(self.reals)

    def at(self, index):
        if isinstance(index, Real):
            return lookup_apply(index, self.reals)
        return self.reals[index]

    def mapLeaves(self, g):
        return g(self)


class MapVec(Vec):
    def __init__(self, original, fn):
        self.original, self.fn = original, fn

    @property
    def size(self):
        return self.original.size

    def at(self, index):
        return self.fn(self.original.at(index))

    def mapLeaves(self, g):
        return MapVec(self.original.mapLeaves(g), self.fn)


class ZipVec(Vec):
    def __init__(self, left, right):
        self.left, self.right = left, right

    @property
    def size(self):
        return self.left.size

    def at(self, index):
        return (self.left.at(index), self.right.at(index))

    def mapLeaves(self, g):
        return ZipVec(self.left.mapLeaves(g), self.right.mapLeaves(g))


class TraverseVec(Vec):
    def __init__(self, lst):
        self.list = list(lst)
        self.size = self.list[0].size
        assert all(v.size == self.size for v in self.list)

    def at(self, index):
        return [v.at(index) for v in self.list]

    def mapLeaves(self, g):
        return TraverseVec([v.mapLeaves(g) for v in self.list])
