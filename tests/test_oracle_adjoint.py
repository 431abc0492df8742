"""
The oracle's reverse-mode density for PRIMAL RIRs (oracle/rainier_oracle.cpp: AdjointDensity -- the reference's Diff rules,
compute/Gradient.scala:71-152, applied to the interpreter's node values) against the oracle's evaluation of the reference's own
symbolic gradient outputs, which the golden vectors pin.  This is the checker of the full-size hierarchical Poisson GLM
(BASELINE configs[4]: 1000 groups / 1M rows), whose symbolic form needs one one-hot column per table entry and cannot be built.
"""
import numpy as np
import pytest

from oracle.rainier_py import configs
from oracle.rainier_py.binding import OracleModel, default_config
from rainier_b200 import abi


@pytest.mark.parametrize("name", ["poisson_100x20k", "logreg", "linreg5", "eight_schools", "funnel"])
def test_adjoint_density_matches_the_symbolic_gradient(name):
    model = {"poisson_100x20k": lambda: configs.poisson_glm(100, 20000), "logreg": lambda: configs.logreg(800, 6),
             "linreg5": lambda: configs.linreg(600, covariates=5), "eight_schools": configs.eight_schools, "funnel": configs.funnel}[name]()
    rir, cols = model.compile(True)
    prir, pcols = model.compile(False)
    a, b = OracleModel(rir, cols), OracleModel(prir, pcols)
    q = np.random.default_rng(5).normal(size=(4, a.n)) * 0.3
    da, db = a.density_batch(q), b.density_batch(q)
    assert np.max(np.abs(da - db) / np.maximum(np.abs(da), 1e-9)) < 1e-10


def test_sampling_on_a_primal_rir_follows_the_symbolic_run():
    model = configs.poisson_glm(20, 2000)
    rir, cols = model.compile(True)
    prir, pcols = model.compile(False)
    c = default_config()
    c.iterations, c.warmup_iterations, c.sampler, c.n_steps = 5, 0, abi.RN_SAMPLER_HMC, 3
    c.step_size_tuner, c.static_step_size, c.mass_tuner = abi.RN_STEP_STATIC, 0.002, abi.RN_MASS_IDENTITY
    ra = OracleModel(rir, cols).sample(c, seeds=np.arange(3) + 1, trace=True)
    rb = OracleModel(prir, pcols).sample(c, seeds=np.arange(3) + 1, trace=True)
    assert np.array_equal(ra["trace"][:, :, 1], rb["trace"][:, :, 1])
    assert np.max(np.abs(ra["samples"] - rb["samples"]) / np.maximum(np.abs(ra["samples"]), 1e-9)) < 1e-9
