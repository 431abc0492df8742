"""
GPU parity of the device-side inlining (rn_model_create -> rn_inline.cpp + rn_k_eval / rn_k_reduce_rows): a
Gaussian regression arrives as the STREAMED primal container (what the Scala side would send without running
TargetGroup.inlinable / PartialEvaluator.inline), the column-only monomials are summed over the rows on the device, and the
model that samples is data-free.  Against the oracle on the streamed form, and against the reference's own inlined model.
"""
import os

import numpy as np
import pytest

from oracle.rainier_py import compute, configs
from oracle.rainier_py.binding import OracleModel
from rainier_b200 import abi, api

import parity

pytestmark = pytest.mark.gpu


def _streamed(model, grad):
    keep = compute.inlinable
    compute.inlinable = lambda real: False
    try:
        return model.compile(grad)
    finally:
        compute.inlinable = keep


@pytest.mark.parametrize("cov", [3, 5])
def test_streamed_gaussian_regression_is_inlined_on_the_device(cov):
    n_obs = 5000
    srir, scols = _streamed(configs.linreg(n_obs, covariates=cov), False)
    m = api.CudaModel(srir, scols)
    targets, monos, rows = m.inlined()
    assert targets == 2 and monos > 0 and rows == 8 + (n_obs - 1) // 8  # both observe() targets: the 8-row init block and the 8 splits
    assert "for (long long row" not in m.emit_source(api.make_config(sampler=api.HMCSampler(2)))  # nothing left to stream
    q = np.random.default_rng(2).normal(size=(16, m.nVars)) * 0.3
    q[:, 0] = np.abs(q[:, 0]) + 0.3
    ref_rir, ref_cols = configs.linreg(n_obs, covariates=cov).compile(True)  # the reference's own form (inlined for 3, streamed for 5)
    ref = OracleModel(ref_rir, ref_cols)
    assert parity.rel_err(m.density_batch(q), ref.density_batch(q), 1e-9) < 1e-9
    # sampling: decision for decision with the oracle on the reference's form (stable regime, like every tolerance-parity test)
    cfg = api.make_config(iterations=30, warmupIterations=0, sampler=api.HMCSampler(4), stepSizeTuner=api.StaticStepSize(0.002),
                          massMatrixTuner=api.IdentityMassMatrixTuner())
    seeds = np.arange(64) + 1
    got = m.sample(cfg, seeds=seeds)
    want = ref.sample(api.lower_config(cfg)[0], seeds=seeds, trace=True)
    assert parity.rel_err(got.chains, want["samples"], 1e-9) < 1e-8
    assert [s.accepted for s in got.stats] == [s.accepted for s in want["stats"]]
    # switched off: the same container keeps streaming, same answers
    os.environ["RN_INLINE"] = "0"
    try:
        m0 = api.CudaModel(srir, scols)
        assert m0.inlined()[0] == 0
        assert parity.rel_err(m0.density_batch(q), ref.density_batch(q), 1e-9) < 1e-9
    finally:
        del os.environ["RN_INLINE"]


def test_nonlinear_likelihood_keeps_streaming():
    prir, pcols = configs.logreg(1500, 6).compile(False)
    assert api.CudaModel(prir, pcols).inlined() == (0, 0, 0)
