"""
oracle/rainier_py/core.py -- TEST INFRASTRUCTURE ONLY (oracle tooling).

Python restatement of the slice of rainier-core needed to build the benchmark / golden-vector models exactly
as the reference would.  Cites are relative to rainier-core/src/main/scala/com/stripe/rainier/core/ (K/).

  Support               K/Support.scala:10-96
  Injection             K/Injection.scala:9-107
  Continuous family     K/Continuous.scala:10-248   (Normal, Cauchy, Laplace, Gamma, Exponential, Beta, LogNormal, Uniform)
  Discrete family       K/Discrete.scala:6-282, K/Multinomial.scala:11-26
  Combinatorics         K/Combinatorics.scala:9-35
  Generator             K/Generator.scala:10-137     (only Const/From, map/flatMap/zip/repeat, categorical, traverse)
  Model                 K/Model.scala:7-133
  SBC                   K/SBC.scala:15-66 (synthesize / fit only)
"""
import math

import numpy as np

from .compute import (Bounds, Column, Evaluator, Real, Scalar, TargetGroup, Vec, compile_function_rir, compile_rir, jd2i,
                      jexp, jlog, jpow, to_real)


# ------------------------------------------------------------------------------------------------------
# Generator  (K/Generator.scala)
# ------------------------------------------------------------------------------------------------------
class Generator:
    """From(requirements, fn) or Const(requirements, t): `const` is None for From."""

    def __init__(self, requirements, fn=None, const=None, is_const=False):
        self.requirements = list(requirements)
        self.fn = fn
        self.const = const
        self.is_const = is_const

    def get(self, r, n):
        return self.const if self.is_const else self.fn(r, n)

    def map(self, f):
        if self.is_const:
            return Generator(self.requirements, const=f(self.const), is_const=True)
        fromFn = self.fn
        return Generator(self.requirements, fn=lambda r, n: f(fromFn(r, n)))

    def flatMap(self, f):  # :20-35
        if self.is_const:
            g = to_generator(f(self.const))
            if g.is_const:
                return Generator(self.requirements + g.requirements, const=g.const, is_const=True)
            return Generator(self.requirements + g.requirements, fn=g.fn)
        fromFn = self.fn

        def inner(r, n):
            g = to_generator(f(fromFn(r, n)))
            return g.const if g.is_const else g.fn(r, n)

        return Generator(self.requirements, fn=inner)

    def zip(self, other):  # :37-47
        reqs = self.requirements + other.requirements
        if self.is_const and other.is_const:
            return Generator(reqs, const=(self.const, other.const), is_const=True)
        if not self.is_const and not other.is_const:
            lf, rf = self.fn, other.fn
            return Generator(reqs, fn=lambda r, n: (lf(r, n), rf(r, n)))
        if self.is_const:
            t, rf = self.const, other.fn
            return Generator(reqs, fn=lambda r, n: (t, rf(r, n)))
        lf, u = self.fn, other.const
        return Generator(reqs, fn=lambda r, n: (lf(r, n), u))

    def repeat(self, k):  # :49-58
        if self.is_const:
            u = self.const
            return Generator(self.requirements, fn=lambda r, n: [u for _ in range(n.toInt(k))])
        fromFn = self.fn
        return Generator(self.requirements, fn=lambda r, n: [fromFn(r, n) for _ in range(n.toInt(k))])

    MaxRequirements = 500  # :101

    def reqs(self):
        """requirements.toList.take(Generator.MaxRequirements) (:61): `requirements` is a Set[Real] -- duplicates
        collapse; the hash order of a Scala Set with more than 4 elements is resolved to insertion order here."""
        seen, out = set(), []
        for r in self.requirements:
            r = to_real(r)
            if r not in seen:
                seen.add(r)
                out.append(r)
        return out[: Generator.MaxRequirements]

    def prepare(self, parameters, rng, make_function):  # :59-94
        """Returns fn(array) like the reference.  make_function(rir_bytes) -> callable([count][n] -> [count][m]) stands
        for `Compiler.default.compile(parameters.map(_.param), namedReqs)` + the CompiledFunction.output loop: the CPU
        oracle's OracleFunction or the CUDA path's CudaFunction."""
        reqs = self.reqs()
        if not reqs:
            return lambda array: self.get(rng, Evaluator({p: float(v) for p, v in zip(parameters, array)}))
        cf = make_function(compile_function_rir(parameters, reqs))

        def fn(array):
            reqValues = cf(np.asarray(array, dtype=np.float64)[None, :])[0]
            cache = {p: float(v) for p, v in zip(parameters, array)}
            cache.update({r: float(v) for r, v in zip(reqs, reqValues)})
            return self.get(rng, Evaluator(cache))

        return fn

    def predict(self, parameters, draws, rng, make_function):
        """Trace.predict (K/Trace.scala:34-41) over draws [count][n] given in predict's order (chain-major): the batched
        form of prepare -- ALL draws' requirement values come from one call of the compiled function, then Generator.get
        runs per draw on the host exactly as in the reference (same RNG consumption order)."""
        draws = np.asarray(draws, dtype=np.float64)
        reqs = self.reqs()
        if not reqs:
            return [self.get(rng, Evaluator({p: float(v) for p, v in zip(parameters, a)})) for a in draws]
        values = make_function(compile_function_rir(parameters, reqs))(draws)
        out = []
        for a, rv in zip(draws, values):
            cache = {p: float(v) for p, v in zip(parameters, a)}
            cache.update({r: float(v) for r, v in zip(reqs, rv)})
            out.append(self.get(rng, Evaluator(cache)))
        return out

    # companion
    @staticmethod
    def constant(t):
        return Generator([], const=t, is_const=True)

    @staticmethod
    def from_(fn):
        return Generator([], fn=fn)

    @staticmethod
    def real(x):
        return Generator([x], fn=lambda _, n: n.toDouble(x))

    @staticmethod
    def require(reqs, fn):
        return Generator(reqs, fn=fn)

    @staticmethod
    def categorical(pmf):  # :121-132 ; pmf: ordered list of (t, p)
        cdf = []
        acc = Real.zero
        for (t, p) in pmf:
            acc = to_real(p) + acc
            cdf.append((t, acc))

        def fn(r, n):
            v = r.standardUniform()
            for (t, p) in cdf:
                if n.toDouble(p) >= v:
                    return t
            return cdf[-1][0]

        return Generator([p for (_, p) in cdf], fn=fn)

    @staticmethod
    def traverse(seq):  # :136-141
        g = Generator.constant([])
        for t in seq:
            g = g.zip(to_generator(t)).map(lambda lr: lr[0] + [lr[1]])
        return g


def to_generator(t):  # ToGenerator instances, K/Generator.scala:152-248
    if isinstance(t, Generator):
        return t
    if isinstance(t, Distribution):
        return t.generator
    if isinstance(t, Real):
        return Generator.require([t], lambda _, n: n.toDouble(t))
    if isinstance(t, tuple):
        g = to_generator(t[0])
        for u in t[1:]:
            g = g.zip(to_generator(u))
        return g
    if isinstance(t, list):
        return Generator.traverse([to_generator(x) for x in t])
    raise TypeError("no ToGenerator for %r" % (t,))


# ------------------------------------------------------------------------------------------------------
# Support  (K/Support.scala)
# ------------------------------------------------------------------------------------------------------
class UnboundedSupport:
    def transform(self, v):
        return v

    def logJacobian(self, v):
        return Real.zero


class BoundedSupport:
    def __init__(self, mn, mx):
        self.min, self.max = to_real(mn), to_real(mx)

    def transform(self, v):  # :59-60
        return v.logistic() * (self.max - self.min) + self.min

    def logJacobian(self, v):  # :62-63
        return v.logistic().log() + (1 - v.logistic()).log() + (self.max - self.min).log()


class BoundedBelowSupport:
    def __init__(self, mn=Real.zero):
        self.min = to_real(mn)

    def transform(self, v):  # :72-73
        return v.exp() + self.min

    def logJacobian(self, v):
        return v


class BoundedAboveSupport:
    def __init__(self, mx=Real.zero):
        self.max = to_real(mx)

    def transform(self, v):  # :84-85
        return self.max - (-1 * v).exp()

    def logJacobian(self, v):
        return v * -1


# ------------------------------------------------------------------------------------------------------
# Distribution / Continuous  (K/Distribution.scala, K/Continuous.scala:10-34)
# ------------------------------------------------------------------------------------------------------
class Distribution:
    pass


class Continuous(Distribution):
    support = None

    def logDensitySeq(self, seq):  # :13
        return Vec.from_([float(v) for v in seq]).map(self.logDensity).columnize()

    def scale(self, a):
        return Scale(a).transform(self)

    def translate(self, b):
        return Translate(b).transform(self)

    def exp(self):
        return Exp.transform(self)

    def latentVec(self, k):
        return Vec.from_([self.latent() for _ in range(k)])


class StandardContinuous(Continuous):  # :27-34
    def latent(self):
        x = Real.parameter(lambda x: self.support.logJacobian(x) + self.logDensity(self.support.transform(x)))
        return self.support.transform(x)


class _Fn(StandardContinuous):
    def __init__(self, support, logDensity, generator):
        self.support = support
        self._ld = logDensity
        self.generator = generator

    def logDensity(self, x):
        return self._ld(to_real(x))


# ------------------------------------------------------------------------------------------------------
# Injection  (K/Injection.scala)
# ------------------------------------------------------------------------------------------------------
class Injection:
    requirements = []

    def fastForwards(self, x, n):
        return n.toDouble(self.forwards(to_real(x)))

    def whenDefinedAt(self, y, ifDefined, notDefined):
        return ifDefined

    def transform(self, dist):  # :25-41
        inj = self

        class _T(Continuous):
            support = inj.transformSupport(dist.support)

            def logDensity(self, real):
                real = to_real(real)
                return inj.whenDefinedAt(real, dist.logDensity(inj.backwards(real)) + inj.logJacobian(real),
                                         Real.negInfinity)

            def latent(self):
                return inj.forwards(dist.latent())

        t = _T()
        distGen = dist.generator
        t.generator = Generator.require(list(inj.requirements) + distGen.requirements,
                                        lambda r, n: inj.fastForwards(distGen.get(r, n), n))
        return t


class Scale(Injection):  # :48-66
    def __init__(self, a):
        self.a = to_real(a)
        self.lj = self.a.log() * -1
        self.requirements = [self.a]

    def forwards(self, x):
        return x * self.a

    def fastForwards(self, x, n):
        return x * n.toDouble(self.a)

    def backwards(self, y):
        return y / self.a

    def logJacobian(self, y):
        return self.lj

    def transformSupport(self, supp):
        if isinstance(supp, UnboundedSupport):
            return supp
        if isinstance(supp, BoundedBelowSupport):
            return BoundedBelowSupport(self.forwards(supp.min))
        if isinstance(supp, BoundedAboveSupport):
            return BoundedAboveSupport(self.forwards(supp.max))
        return BoundedSupport(self.forwards(supp.min), self.forwards(supp.max))


class Translate(Injection):  # :71-86
    def __init__(self, b):
        self.b = to_real(b)
        self.requirements = [self.b]

    def forwards(self, x):
        return x + self.b

    def fastForwards(self, x, n):
        return x + n.toDouble(self.b)

    def backwards(self, y):
        return y - self.b

    def logJacobian(self, y):
        return Real.zero

    def transformSupport(self, supp):
        if isinstance(supp, UnboundedSupport):
            return supp
        if isinstance(supp, BoundedBelowSupport):
            return BoundedBelowSupport(self.forwards(supp.min))
        if isinstance(supp, BoundedAboveSupport):
            return BoundedAboveSupport(self.forwards(supp.max))
        return BoundedSupport(self.forwards(supp.min), self.forwards(supp.max))


class _Exp(Injection):  # :91-107
    requirements = []

    def forwards(self, x):
        return x.exp()

    def fastForwards(self, x, n):
        return jexp(x)

    def backwards(self, y):
        return y.log()

    def logJacobian(self, y):
        return y.log() * -1

    def whenDefinedAt(self, y, whenDefined, notDefined):
        return Real.gt(y, Real.zero, whenDefined, notDefined)

    def transformSupport(self, supp):
        if isinstance(supp, UnboundedSupport):
            return supp
        if isinstance(supp, BoundedBelowSupport):
            return BoundedBelowSupport(self.forwards(supp.min))
        if isinstance(supp, BoundedAboveSupport):
            return BoundedSupport(Real.zero, self.forwards(supp.max))
        return BoundedSupport(self.forwards(supp.min), self.forwards(supp.max))


Exp = _Exp()


# ------------------------------------------------------------------------------------------------------
# Combinatorics  (K/Combinatorics.scala)
# ------------------------------------------------------------------------------------------------------
class Combinatorics:
    @staticmethod
    def gamma(z):
        z = to_real(z)
        if z == Real.zero:
            return Real.infinity
        if z == Real.one or z == Real.two:
            return Real.zero
        return Combinatorics._approxGamma(z)

    @staticmethod
    def beta(a, b):
        a, b = to_real(a), to_real(b)
        return Combinatorics.gamma(a) + Combinatorics.gamma(b) - Combinatorics.gamma(a + b)

    @staticmethod
    def factorial(k):
        return Combinatorics.gamma(to_real(k) + 1)

    @staticmethod
    def choose(n, k):
        n, k = to_real(n), to_real(k)
        return Combinatorics.factorial(n) - Combinatorics.factorial(k) - Combinatorics.factorial(n - k)

    @staticmethod
    def _approxGamma(z):  # :25-34
        v = z + 1
        w = v + (Real.one / ((12 * v) - (Real.one / (10 * v))))
        return (to_real(math.pi * 2).log() / 2) - (v.log() / 2) + (v * (w.log() - 1)) - z.log()


# ------------------------------------------------------------------------------------------------------
# Continuous distributions  (K/Continuous.scala:36-248)
# ------------------------------------------------------------------------------------------------------
class LocationScaleFamily:  # :39-58
    def __init__(self, logDensity, generate):
        self.logDensity = logDensity
        self.generate = generate
        self.standard = _Fn(UnboundedSupport(), logDensity, Generator.from_(lambda r, _: generate(r)))

    def __call__(self, location, scale):
        return self.standard.scale(to_real(scale)).translate(to_real(location))


Normal = LocationScaleFamily(  # :63-67
    lambda x: ((x * x) / -2.0) - 0.5 * to_real(2 * math.pi).log(),
    lambda r: r.standardNormal())

Cauchy = LocationScaleFamily(  # :72-77
    lambda x: (((x * x) + 1) * math.pi).log() * -1,
    lambda r: _div(r.standardNormal(), r.standardNormal()))


def _div(a, b):
    return a / b if b != 0.0 else (math.copysign(float("inf"), a) if a != 0 else float("nan"))


def _laplace_gen(r):  # :85-88
    u = r.standardUniform() - 0.5
    sgn = 0.0 if u == 0 else math.copysign(1.0, u)
    return sgn * -1 * jlog(1 - (2 * abs(u)))


Laplace = LocationScaleFamily(lambda x: to_real(0.5).log() - x.abs(), _laplace_gen)  # :82-89


class Gamma:  # :94-146
    def __new__(cls, shape, scale):
        return Gamma.standard(to_real(shape)).scale(to_real(scale))

    @staticmethod
    def meanAndScale(mean, scale):
        mean, scale = to_real(mean), to_real(scale)
        return Gamma(mean / scale, scale)

    @staticmethod
    def standard(shape):
        shape = to_real(shape)

        def logDensity(real):
            return Bounds.positive(real, lambda: (shape - 1) * real.log() - Combinatorics.gamma(shape) - real)

        def generate(a, r):  # :125-144 (Marsaglia-Tsang)
            while True:
                d = a - 1.0 / 3.0
                c = (1.0 / 3.0) / math.sqrt(d)
                x = r.standardNormal()
                v = 1.0 + c * x
                while v <= 0:
                    x = r.standardNormal()
                    v = 1.0 + c * x
                v3 = v * v * v
                u = r.standardUniform()
                if (u < 1 - 0.0331 * x * x * x * x) or (jlog(u) < 0.5 * x * x + d * (1 - v3 + jlog(v3))):
                    return d * v3

        def gen(r, n):  # :114-122
            a = n.toDouble(shape)
            if a < 1:
                u = r.standardUniform()
                return generate(a + 1, r) * jpow(u, 1.0 / a)
            return generate(a, r)

        return _Fn(BoundedBelowSupport(Real.zero), logDensity, Generator.require([shape], gen))


class Exponential:  # :151-157
    standard = None

    def __new__(cls, rate):
        if Exponential.standard is None:
            Exponential.standard = Gamma.standard(1.0)
        return Exponential.standard.scale(Real.one / to_real(rate))


class Beta(StandardContinuous):  # :162-184
    def __init__(self, a, b):
        self.a, self.b = to_real(a), to_real(b)
        self.support = BoundedSupport(Real.zero, Real.one)
        self.generator = Gamma(self.a, 1).generator.zip(Gamma(self.b, 1).generator).map(lambda z: z[0] / (z[0] + z[1]))

    def logDensity(self, real):
        real = to_real(real)
        return Bounds.zeroToOne(real, lambda: self._betaDensity(real))

    def _betaDensity(self, u):
        a, b = self.a, self.b
        return (a - 1) * u.log() + (b - 1) * (1 - u).log() - Combinatorics.beta(a, b)


def LogNormal(location, scale):  # :196-199
    return Normal(location, scale).exp()


class Uniform:  # :204-218
    _standard = None

    def __new__(cls, frm, to):
        if Uniform._standard is None:
            beta11 = Beta(1, 1)
            Uniform._standard = _Fn(beta11.support, beta11.logDensity, Generator.from_(lambda r, _: r.standardUniform()))
        frm, to = to_real(frm), to_real(to)
        return Uniform._standard.scale(to - frm).translate(frm)


# ------------------------------------------------------------------------------------------------------
# Discrete distributions  (K/Discrete.scala, K/Multinomial.scala)
# ------------------------------------------------------------------------------------------------------
class Discrete(Distribution):
    def logDensitySeq(self, seq):  # :7-8
        return Vec.from_([float(v) for v in seq]).map(self.logDensity).columnize()


class Bernoulli(Discrete):  # :38-52
    def __init__(self, p):
        self.p = p = to_real(p)

        def gen(r, n):
            u = r.standardUniform()
            l = n.toDouble(p)
            return 1 if u <= l else 0

        self.generator = Generator.require([p], gen)

    def logDensity(self, v):
        return Real.eq(to_real(v), Real.zero, (1 - self.p).log(), self.p.log())


class Geometric(Discrete):  # :59-73
    def __init__(self, p):
        self.p = p = to_real(p)

        def gen(r, n):
            u = r.standardUniform()
            q = n.toDouble(p)
            return int(math.floor(_div(jlog(u), jlog(1 - q))))

        self.generator = Generator.require([p], gen)

    def logDensity(self, v):
        return self.p.log() + to_real(v) * (1 - self.p).log()


class NegativeBinomial(Discrete):  # :81-115
    def __init__(self, p, n):
        self.p, self.n = p, nn = to_real(p), to_real(n)
        p = self.p

        def nb(r, m):
            total = 0
            for _ in range(m.toLong(nn)):
                total += Geometric(1 - p).generator.get(r, m)
            return total

        normalGenerator = Normal(nn * p / (1 - p), (nn * p).pow(1.0 / 2.0) / (1 - p)).generator.map(lambda x: max(int(x), 0))

        def gen(r, m):
            pD, nD = m.toDouble(p), m.toDouble(nn)
            if pD < -100 / nD + 1 and pD > 100 / nD - .25:
                return normalGenerator.get(r, m)
            return nb(r, m)

        self.generator = Generator.from_(gen)

    def logDensity(self, v):
        v = to_real(v)
        n, p = self.n, self.p
        return (Combinatorics.factorial(n + v - 1) - Combinatorics.factorial(v) - Combinatorics.factorial(n - 1) +
                n * (1 - p).log() + v * p.log())


class Poisson(Discrete):  # :122-198
    def __init__(self, lam):
        self.lam = lam = to_real(lam)

        def gen(r, n):
            l = n.toDouble(lam)
            if l < 30.0:
                return Poisson.small(l, r)
            return Poisson.large(l, r)

        self.generator = Generator.require([lam], gen)

    def logDensity(self, v):
        v = to_real(v)
        return self.lam.log() * v - self.lam - Combinatorics.factorial(v)

    @staticmethod
    def small(lam, r):  # :142-153
        l = jexp(-lam)
        if l >= 1.0:
            return 0
        k = 0
        p = 1.0
        while p > l:
            k += 1
            p *= r.standardUniform()
        return k - 1

    @staticmethod
    def large(lam, r):  # :156-178
        c = 0.767 - 3.36 / lam
        beta = math.pi / math.sqrt(3.0 * lam)
        alpha = beta * lam
        k = jlog(c) - lam - jlog(beta)
        while True:
            u = r.standardUniform()
            x = (alpha - jlog(_div(1.0 - u, u))) / beta
            n = int(math.floor(x + 0.5))
            if n >= 0:
                v = r.standardUniform()
                y = alpha - beta * x
                lhs = y + jlog(v / jpow(1.0 + jexp(y), 2))
                rhs = k + n * jlog(lam) - Poisson._logFactorial(n)
                if lhs <= rhs:
                    return n

    @staticmethod
    def _logFactorial(n):  # :183-186
        x = float(n + 1)
        return ((x - 0.5) * jlog(x)) - x + (0.5 * jlog(2 * math.pi))


class Multinomial(Distribution):  # K/Multinomial.scala:11-26 ; pmf is an ordered list of (key, p)
    def __init__(self, pmf, k):
        self.pmf = [(t, to_real(p)) for (t, p) in pmf]
        self.k = to_real(k)

        def count(seq):
            out = {}
            for t in seq:
                out[t] = out.get(t, 0) + 1
            return out

        self.generator = Generator.categorical(self.pmf).repeat(self.k).map(count)

    def logDensity(self, v):  # v: ordered list of (key, Real)
        terms = []
        for (t, i) in v:
            i = to_real(i)
            p = Real.zero
            for (tt, pp) in self.pmf:
                if tt == t:
                    p = pp
            pTerm = Real.eq(i, Real.zero, Real.zero, i * p.log())
            terms.append(pTerm - Combinatorics.factorial(i))
        return Combinatorics.factorial(self.k) + Real.sum(terms)


class Binomial(Discrete):  # K/Discrete.scala:206-243
    def __init__(self, p, k):
        self.p, self.k = p, k = to_real(p), to_real(k)
        self.multi = Multinomial([(True, p), (False, 1 - p)], k)
        kGenerator = Generator.real(k)
        lazy = {}

        def poissonGenerator():
            if "p" not in lazy:
                lazy["p"] = Poisson(p * k).generator.zip(kGenerator).map(lambda xk: min(xk[0], int(xk[1])))
            return lazy["p"]

        def normalGenerator():
            if "n" not in lazy:
                lazy["n"] = Normal(k * p, (k * p * (1 - p)).pow(0.5)).generator.zip(kGenerator).map(
                    lambda xk: min(max(int(xk[0]), 0), int(xk[1])))
            return lazy["n"]

        binomialGenerator = self.multi.generator.map(lambda m: m.get(True, 0))

        def gen(r, n):
            pD, kD = n.toDouble(p), n.toDouble(k)
            if kD >= 100 and kD * pD <= 10:
                return poissonGenerator().get(r, n)
            if kD >= 100 and kD * pD >= 9 and kD * (1.0 - pD) >= 9:
                return normalGenerator().get(r, n)
            return binomialGenerator.get(r, n)

        self.generator = Generator.require([p, k], gen)

    def logDensity(self, v):
        v = to_real(v)
        return self.multi.logDensity([(True, v), (False, self.k - v)])


# ------------------------------------------------------------------------------------------------------
# Model  (K/Model.scala)
# ------------------------------------------------------------------------------------------------------
NumSplits = 8


def _split_seq(ts):  # :115-132
    n = len(ts)
    splitSize = (n - 1) // NumSplits
    initSize = n - (splitSize * NumSplits)
    init = list(ts[:initSize])
    if splitSize == 0:
        return init, []
    return init, [list(ts[initSize + i * splitSize: initSize + (i + 1) * splitSize]) for i in range(NumSplits)]


def _split_vec(ts):  # :98-113
    n = ts.size
    splitSize = (n - 1) // NumSplits
    initSize = n - (splitSize * NumSplits)
    init = ts.take(initSize)
    if splitSize == 0:
        return init, []
    return init, [ts.slice(initSize + i * splitSize, initSize + (i + 1) * splitSize) for i in range(NumSplits)]


class Model:
    def __init__(self, likelihoods, track=()):
        self.likelihoods = list(likelihoods)
        self.track = list(track)
        self._tg = {}

    def merge(self, other):
        return Model(self.likelihoods + other.likelihoods, self.track + other.track)

    @staticmethod
    def track_(track):  # :67
        return Model([Real.zero], track)

    @staticmethod
    def likelihood(real):
        return Model([real])

    @staticmethod
    def observe(ys, lh):  # :71-96
        if isinstance(lh, Vec):
            initX, splitsX = _split_vec(lh)
            if not splitsX:
                return Model([initX.columnize().logDensitySeq(ys)])
            initY, splitsY = _split_seq(ys)
            return Model([initX.columnize().logDensitySeq(initY),
                          Real.sum([sx.columnize().logDensitySeq(sy) for sx, sy in zip(splitsX, splitsY)])])
        if not isinstance(ys, (list, tuple, np.ndarray)):
            ys = [ys]
        init, splits = _split_seq(list(ys))
        initReal = lh.logDensitySeq(init)
        if not splits:
            return Model([initReal])
        return Model([initReal, Real.sum([lh.logDensitySeq(s) for s in splits])])

    def targetGroup(self, with_gradient=True):  # :32
        if with_gradient not in self._tg:
            self._tg[with_gradient] = TargetGroup(self.likelihoods, self.track, with_gradient)
        return self._tg[with_gradient]

    @property
    def parameters(self):
        if self._tg:
            return next(iter(self._tg.values())).parameters
        return self.targetGroup(True).parameters

    def compile(self, with_gradient=True):
        """Compiler.default.compileTargets(targetGroup) -> (rir bytes, [column arrays])"""
        return compile_rir(self.targetGroup(with_gradient))


Model.empty = Model([Real.zero])


# ------------------------------------------------------------------------------------------------------
# SBC  (K/SBC.scala:15-66)
# ------------------------------------------------------------------------------------------------------
class SBC:
    def __init__(self, priors, fn):
        if not isinstance(priors, (list, tuple)):
            prior = priors
            f1 = fn
            priors = [prior]
            fn = lambda l: (f1(l[0]), l[0])  # noqa: E731  (SBC.apply(prior)(fn), :167-170)
        self.priors = list(priors)
        self.fn = fn
        self.priorGenerator = Generator.traverse([p.generator for p in self.priors])

    def synthesize(self, samples, rng):  # :53-61
        def inner(priorParams):
            d, r = self.fn([to_real(p) for p in priorParams])
            return d.generator.repeat(to_real(samples)).zip(Generator.real(r))

        return self.priorGenerator.flatMap(inner).get(rng, Evaluator())

    def fit(self, values):  # :63-66
        d, r = self.fn([p.latent() for p in self.priors])
        return Model.observe(values, d), r
