"""
oracle/rainier_py/configs.py -- TEST INFRASTRUCTURE ONLY: the five BASELINE.json configurations built the way
the reference would build them, plus the synthetic inputs of SURVEY.md §8(d).  All draws come
from the restated java.util.Random so that a Scala harness can reproduce the inputs.
"""
import math

import numpy as np

from .binding import JRandom, ScalaRNG
from .compute import Real, Vec
from .core import Bernoulli, Cauchy, Exponential, Model, Normal, Poisson, Uniform


def funnel(dim=10):
    """cfg 1: Neal's funnel, centred form: theta0 ~ N(0,3), theta_i ~ N(0, exp(theta0/2)).
    Real.parameters(10){...} + Model.track (compute/Real.scala:70-78, core/Model.scala:67)."""
    thetas = Real.parameters(dim, lambda th: Normal(0, 3).logDensity(th[0]) + Real.sum(
        [Normal(0, (th[0] / 2).exp()).logDensity(th[i]) for i in range(1, dim)]))
    return Model.track_(list(thetas))


def linreg_data(n=10000, seed=20260923):
    r = JRandom(seed)
    xs, ys = [], []
    for _ in range(n):
        u, v, w = (r.nextGaussian() * 3 for _ in range(3))
        y = 0.5 + 1.0 * u - 2.0 * v + 0.5 * w + 0.7 * r.nextGaussian()
        xs.append((u, v, w))
        ys.append(y)
    return xs, ys


def linreg(n=10000, seed=20260923, covariates=3):
    """cfg 2: README linear regression (README.md:20-31).  With 3 covariates the reference inlines the data away;
    with >= 4 it streams (SURVEY.md §8 'why cfg 2 ends up data-free')."""
    if covariates == 3:
        xs, ys = linreg_data(n, seed)
        sigma = Exponential(1).latent()
        alpha = Normal(0, 1).latent()
        betas = Normal(0, 1).latentVec(3)
        return Model.observe(ys, Vec.from_(xs).map(lambda t: Normal(alpha + Vec.of(*t).dot(betas), sigma)))
    r = JRandom(seed)
    g = r.gaussians(n * (covariates + 1)).reshape(n, covariates + 1)
    X = g[:, :covariates] * 3
    coef = np.array([(-1) ** j * (0.5 + 0.25 * j) for j in range(covariates)])
    ys = 0.5 + X @ coef + 0.7 * g[:, covariates]
    sigma = Exponential(1).latent()
    alpha = Normal(0, 1).latent()
    betas = Normal(0, 1).latentVec(covariates)
    return Model.observe(list(ys), Vec.from_([list(row) for row in X]).map(lambda x: Normal(alpha + x.dot(betas), sigma)))


def logreg_data(n=100000, d=50, seed=20260924):
    r = JRandom(seed)
    beta_true = r.gaussians(d)
    X = np.empty((n, d))
    ys = np.empty(n, dtype=np.int64)
    sd = math.sqrt(d)
    for i in range(n):
        X[i] = r.gaussians(d) / sd
        p = 1.0 / (1.0 + math.exp(-float(X[i] @ beta_true)))
        ys[i] = 1 if r.nextDouble() <= p else 0
    return X, ys


def logreg(n=100000, d=50, seed=20260924):
    """cfg 3: logistic regression, Bernoulli(x.dot(betas).logistic) (core/Discrete.scala:38-52)."""
    X, ys = logreg_data(n, d, seed)
    betas = Normal(0, 1).latentVec(d)
    return Model.observe([int(y) for y in ys], Vec.from_([list(row) for row in X]).map(lambda x: Bernoulli(x.dot(betas).logistic())))


def eight_schools_parts():
    """cfg 4: rainier-benchmark/.../bench/stan/EightSchools.scala:9-24 verbatim; returns (model, mu, tau, thetas, sigmas)."""
    ys = [28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0]
    sigmas = [15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0]
    mu = Normal(0, 5).latent()
    tau = Cauchy(0, 5).latent().abs()
    thetas = Normal(mu, tau).latentVec(8)
    model = Model.empty
    for i, (y, s) in enumerate(zip(ys, sigmas)):
        model = model.merge(Model.observe(y, Normal(thetas.at(i), s)))
    return model, mu, tau, thetas, sigmas


def eight_schools():
    return eight_schools_parts()[0]


def eight_schools_derived(mu, tau, thetas):
    """12 derived quantities of eight schools used as posterior-predictive requirements (Trace.predict) in fixtures,
    tests and scripts/bench_function.py: mu, tau, the 8 thetas, log|theta_0 - theta_1|, tau^mu."""
    t = [thetas.at(i) for i in range(8)]
    return [mu, tau] + t + [(t[0] - t[1]).abs().log(), tau.pow(mu)]


def poisson_glm_data(groups=1000, n=1000000, seed=20260925):
    rng = ScalaRNG(seed)
    r = rng.rand
    a = 1.0 + 0.5 * r.gaussians(groups)
    g = np.arange(n) % groups
    xs = np.empty(n)
    ys = np.empty(n, dtype=np.int64)
    for i in range(n):
        x = r.nextGaussian()
        xs[i] = x
        ys[i] = Poisson.small(math.exp(a[g[i]] + 0.3 * x), rng)
    return g, xs, ys


def poisson_glm(groups=1000, n=1000000, seed=20260925):
    """cfg 5: rainier-benchmark/.../bench/stan/GLMMPoisson2.scala:24-57 scaled."""
    g, xs, ys = poisson_glm_data(groups, n, seed)
    mu = Normal(0, 10).latent()
    sdAlpha = Uniform(0, 2).latent()
    alphas = Normal(mu, sdAlpha).latentVec(groups)
    beta = Normal(0, 10).latent()
    rows = [(float(gi), float(xi)) for gi, xi in zip(g, xs)]
    return Model.observe([int(y) for y in ys], Vec.from_(rows).map(lambda t: Poisson((alphas.at(t[0]) + beta * t[1]).exp())))
