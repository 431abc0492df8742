"""
oracle/rainier_py/sbc_models.py -- TEST INFRASTRUCTURE ONLY.

The reference's SBC golden-vector models (rainier-test/src/main/scala/com/stripe/rainier/core/SBCModel.scala:46-267)
and the harness around them (SBCModel trait, :6-42), restated so the CPU oracle can be pinned to the reference's
own known-answer vectors without a JVM.
"""
from rainier_b200 import abi

from .binding import OracleModel, ScalaRNG, default_config
from .compute import Evaluator
from .core import (SBC, Bernoulli, Binomial, Exponential, Gamma, Geometric, Laplace, LogNormal, NegativeBinomial,
                   Normal, Poisson, Uniform)

# name -> () => SBC           (SBCModel.scala line of the `def sbc`)
MODELS = {
    "SBCUniformNormal": lambda: SBC(Uniform(0, 1), lambda x: Normal(x, 1)),                      # :47
    "SBCLogNormal": lambda: SBC(LogNormal(0, 1), lambda x: LogNormal(x, x)),                     # :65
    "SBCExponential": lambda: SBC(LogNormal(0, 1), lambda x: Exponential(x)),                    # :83
    "SBCLaplace": lambda: SBC(LogNormal(0, 1), lambda x: Laplace(x, x)),                         # :97
    "SBCGamma": lambda: SBC(LogNormal(0, 1), lambda x: Gamma(x, x)),                             # :111
    "SBCBernoulli": lambda: SBC(Uniform(0, 1), lambda x: Bernoulli(x)),                          # :127
    "SBCBinomial": lambda: SBC(Uniform(0, 1), lambda x: Binomial(x, 10)),                        # :145
    "SBCGeometric": lambda: SBC(Uniform(0, 1), lambda x: Geometric(x)),                          # :163
    "SBCNegativeBinomial": lambda: SBC(Uniform(0, 1), lambda x: NegativeBinomial(x, 10)),        # :199
    "SBCBinomialPoissonApproximation": lambda: SBC(Uniform(0, 0.04), lambda x: Binomial(x, 200)),  # :217
    "SBCLargePoisson": lambda: SBC(Uniform(0.8, 1), lambda x: Poisson(x * 1000)),                # :255
}
# enabled list of rainier-test/src/test/scala/com/stripe/rainier/core/SBCTest.scala:20-34
ENABLED = list(MODELS.keys())


def sbc_config(iterations, warmup=10000):
    """SBCModel.sampler(it), SBCModel.scala:11-21: HMCSampler(1), DualAvgTuner(0.8), IdentityMassMatrixTuner."""
    c = default_config()
    c.iterations = iterations
    c.warmup_iterations = warmup
    c.stats_window = 100
    c.sampler = abi.RN_SAMPLER_HMC
    c.n_steps = 1
    c.step_size_tuner = abi.RN_STEP_DUAL_AVG
    c.delta = 0.8
    c.mass_tuner = abi.RN_MASS_IDENTITY
    return c


def build(name, seed=1528673302081, synthetic_samples=1000):
    """SBCModel.scala:31-36 up to (and excluding) model.sample: returns (model, real, rng, trueValue)."""
    rng = ScalaRNG(seed)
    sbc = MODELS[name]()
    values, trueValue = sbc.synthesize(synthetic_samples, rng)
    model, real = sbc.fit(values)
    return model, real, rng, trueValue


def run(name, n_samples, seed=1528673302081, synthetic_samples=1000, warmup=10000):
    """SBCModel.scala:31-39: synthesize, fit, sample 1 chain on the SAME rng stream, predict(real)."""
    model, real, rng, _ = build(name, seed, synthetic_samples)
    rir, cols = model.compile(with_gradient=True)
    om = OracleModel(rir, cols)
    cfg = sbc_config(n_samples, warmup)
    res = om.sample(cfg, rng_states=[rng.rand.state()])
    params = model.parameters
    out = []
    for a in res["samples"][0]:
        ev = Evaluator({p: float(v) for p, v in zip(params, a)})
        out.append(ev.toDouble(real))
    return out
