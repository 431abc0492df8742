"""
GPU parity of the chain-batched fp64 tensor-core path (DMMA, rn_emit.cpp: Emitter::mma_block): the dot products of a streamed
row body (the Translator's fold of a `Line` with column coefficients, compute/Translator.scala:91-125) evaluated for the 8
chains of a CTA with mma.sync.m8n8k4.f64, the elementwise code on the C fragments, the adjoints X^T w by a second DMMA.
Checked against the oracle like every warp-per-chain test: bit-equal accept decisions, 1e-9 samples (stable regime), and
against the rows-across-lanes path of the same kernel.
"""
import os

import numpy as np
import pytest

from oracle.rainier_py import configs
from rainier_b200 import abi, api

import parity

pytestmark = pytest.mark.gpu


def _cfg(it, warm, sampler, step, mass=None, **kw):
    return api.make_config(iterations=it, warmupIterations=warm, sampler=sampler, stepSizeTuner=step,
                           massMatrixTuner=mass or api.IdentityMassMatrixTuner(), backend=abi.RN_BACKEND_WARP, **kw)


@pytest.mark.parametrize("name", ["logreg", "linreg5"])
def test_dmma_path_matches_the_oracle(name):
    """logistic regression: pure dots; Gaussian regression on 5 covariates: the fold starts from the observation column and
    the elementwise code reads per-chain invariants (sigma) published through shared memory"""
    model, eps = {"logreg": (configs.logreg(1500, 6), 0.01), "linreg5": (configs.linreg(900, covariates=5), 0.002)}[name]
    rir, cols = model.compile(True)
    prir, pcols = model.compile(False)  # primal RIR: the emitter differentiates it (what the Scala CudaCompiler sends)
    cfg = _cfg(12, 0, api.HMCSampler(4), api.StaticStepSize(eps))
    src = api.CudaModel(prir, pcols, device=-1).emit_source(cfg)
    assert "rn_dmma(z" in src and "#define RN_MMA_BARS 8" in src, "the model should take the DMMA path"
    r = parity.run_both(rir, cols, cfg, seeds=np.arange(64) + 3, rir_gpu=prir, cols_gpu=pcols)
    parity.assert_parity(r, tol=1e-9, check_mass=False)


def test_dmma_and_rows_across_lanes_agree_and_ragged_chain_counts_fall_back():
    """same kernel source, RN_MMA=0 -> rows across lanes; 13 chains (not a multiple of 8) run the per-warp path of the DMMA build"""
    model = configs.logreg(1500, 6)
    prir, pcols = model.compile(False)
    cfg = _cfg(10, 0, api.HMCSampler(3), api.StaticStepSize(0.01))
    seeds = np.arange(40) + 11
    a = api.CudaModel(prir, pcols).sample(cfg, seeds=seeds)
    os.environ["RN_MMA"] = "0"
    try:
        m0 = api.CudaModel(prir, pcols)
        assert "rn_dmma(z" not in m0.emit_source(cfg)
        b = m0.sample(cfg, seeds=seeds)
    finally:
        del os.environ["RN_MMA"]
    assert parity.rel_err(a.chains, b.chains) < 1e-9
    c = api.CudaModel(prir, pcols).sample(cfg, seeds=seeds[:13])
    assert parity.rel_err(c.chains, b.chains[:13]) < 1e-9
    # 40 chains = two full groups of 16 on the DMMA path + a tail of 8 on the per-warp path (second launch): the first 32 are
    # bit-identical to the same chains run as a batch of 32
    d = api.CudaModel(prir, pcols).sample(cfg, seeds=seeds[:32])
    assert np.array_equal(a.chains[:32], d.chains)


def test_dmma_with_adaptation_short_horizon():
    """DualAvg + diagonal mass adaptation on the DMMA path: decision for decision with the oracle over the early warmup
    (before rounding differences of the reordered row sums are amplified, see tests/test_zz_gpu_wpc_dense.py)"""
    model = configs.logreg(1500, 6)
    rir, cols = model.compile(True)
    prir, pcols = model.compile(False)
    cfg = _cfg(0, 10, api.HMCSampler(3), api.DualAvgTuner(0.8), api.DiagonalMassMatrixTuner(4, 1.5, 2, 2))
    r = parity.run_both(rir, cols, cfg, seeds=np.arange(32) + 5, rir_gpu=prir, cols_gpu=pcols)
    gt, rt = r["gpu_trace"], r["ref_trace"]
    assert np.array_equal(gt[:, :, 1], rt[:, :, 1]) and np.array_equal(gt[:, :, 3], rt[:, :, 3])
    assert parity.rel_err(gt[:, :, 2], rt[:, :, 2]) < 1e-6
