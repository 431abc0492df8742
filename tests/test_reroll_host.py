"""The re-roll of the invariant sections (rainier_b200/csrc/rn_emit.cpp: rr_plan) on model shapes beyond cfg 5's: every
warp-per-chain density must equal its unrolled form (RN_NO_REROLL=1) to rounding of the re-associated sums and the oracle to
1e-9, whatever mixture of families, uniform operands, gathers and leftovers the model produces.  Host emulation, no GPU."""
import os

import numpy as np
import pytest

from oracle.rainier_py import configs
from oracle.rainier_py.core import Cauchy, Gamma, LogNormal, Model, Normal, Poisson, Uniform, Vec
from oracle.rainier_py.binding import OracleModel
from rainier_b200 import abi, api

import host_emulation as he


def _glm(groups, n, seed, prior):
    """Poisson GLM over `groups` latent effects drawn from `prior(mu, sd)`; rows = (group, x)"""
    rng = np.random.default_rng(seed)
    g = rng.permutation(np.arange(n) % groups)  # every group present, random order: scatter conflicts inside a warp
    xs = rng.normal(size=n)
    ys = rng.poisson(np.exp(0.3 + 0.2 * xs))
    mu = Normal(0, 2).latent()
    sd = Uniform(0, 2).latent()
    alphas = prior(mu, sd, groups)
    beta = Normal(0, 3).latent()
    rows = [(float(gi), float(xi)) for gi, xi in zip(g, xs)]
    return Model.observe([int(y) for y in ys], Vec.from_(rows).map(lambda t: Poisson((alphas.at(t[0]) + beta * t[1]).exp())))


def _two_vectors(groups, n, seed):
    """two latent vectors of different lengths and priors, one looked up, one summed into the rate directly"""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, groups, size=n)
    xs = rng.normal(size=n)
    ys = rng.poisson(1.5, size=n)
    a = Normal(0.1, 0.7).latentVec(groups)
    tau = Gamma(2.0, 0.5).latent()
    b = Normal(0, tau).latentVec(33)
    extra = b.at(0)
    for i in range(1, 33):
        extra = extra + b.at(i) * (0.01 * i)
    rows = [(float(gi), float(xi)) for gi, xi in zip(g, xs)]
    return Model.observe([int(y) for y in ys], Vec.from_(rows).map(lambda t: Poisson((a.at(t[0]) + extra * t[1]).exp())))


CASES = {
    "noncentred_normal_40": lambda: _glm(40, 480, 1, lambda mu, sd, k: Normal(mu, sd).latentVec(k)),
    "lognormal_effects_64": lambda: _glm(64, 512, 2, lambda mu, sd, k: LogNormal(mu * 0.1, sd * 0.2 + 0.1).latentVec(k)),
    "cauchy_effects_35": lambda: _glm(35, 420, 3, lambda mu, sd, k: Cauchy(mu, sd + 0.5).latentVec(k)),
    "two_vectors_48_33": lambda: _two_vectors(48, 384, 4),
    "logreg_40_covariates": lambda: configs.logreg(320, 40),  # 40 iid priors: a fold family and a gradient family, no table
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("k", ["1", "2"])
def test_rerolled_density_equals_unrolled_and_oracle(name, k):
    model = CASES[name]()
    rir, cols = model.compile(True)
    prir, pcols = model.compile(False)
    om = OracleModel(rir, cols)
    cfg = api.make_config(iterations=2, warmupIterations=0, sampler=api.HMCSampler(2), stepSizeTuner=api.StaticStepSize(0.01),
                          massMatrixTuner=api.IdentityMassMatrixTuner())
    cfg.backend = abi.RN_BACKEND_WARP
    os.environ["RN_WPC_K"] = k
    os.environ["RN_MMA"] = "0"
    try:
        cm = api.CudaModel(prir, pcols, device=-1)
        src = cm.emit_source(cfg)
        os.environ["RN_NO_REROLL"] = "1"
        src0 = api.CudaModel(prir, pcols, device=-1).emit_source(cfg)
    finally:
        for v in ("RN_WPC_K", "RN_MMA", "RN_NO_REROLL"):
            os.environ.pop(v, None)
    dens = src[src.index("// ---- emitted"):src.index("// rn_sampler_wpc.cuh --")]
    dens0 = src0[src0.index("// ---- emitted"):src0.index("// rn_sampler_wpc.cuh --")]
    assert "k += RN_G) {" in dens and dens.count("\n") < dens0.count("\n"), "no family was re-rolled"
    q = np.random.default_rng(11).normal(size=(3, cm.nVars)) * 0.4
    d1, e1 = he.density(src, q, None, cm)
    d0, e0 = he.density(src0, q, None, cm)
    ref = om.density_batch(q)
    assert e0 == 0 and e1 == 0
    scale = np.maximum(np.abs(d0), 1e-6)
    assert np.max(np.abs(d1 - d0) / scale) < 1e-11, name
    assert np.max(np.abs(d1 - ref) / np.maximum(np.abs(ref), 1e-6)) < 1e-9, name
