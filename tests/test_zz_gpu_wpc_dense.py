"""
GPU parity of dense mass matrices on the warp-per-chain kernels (MassMatrix.scala:15-117, MassMatrixEstimator.scala:9-50;
opt-in with RN_BACKEND_WARP -- AUTO keeps dense configurations on the thread-per-chain kernels): adaptive
(DenseMassMatrixTuner) and static (StaticMassMatrix(DenseMassMatrix)), data-free (bit-exact accept decisions, 1e-9 samples)
and streamed.  (Named test_zz_* so that it runs after the files of the hot path proper.)
"""
import numpy as np
import pytest

from oracle.rainier_py import configs
from rainier_b200 import abi, api

import parity

pytestmark = pytest.mark.gpu


def _cfg(it, warm, sampler, step, mass):
    return api.make_config(iterations=it, warmupIterations=warm, sampler=sampler, stepSizeTuner=step, massMatrixTuner=mass,
                           backend=abi.RN_BACKEND_WARP)


def test_adaptive_dense_data_free():
    rir, cols = configs.eight_schools().compile(True)
    cfg = _cfg(40, 300, api.EHMCSampler(64, 1, 20, 0.1), api.DualAvgTuner(0.8), api.DenseMassMatrixTuner(40, 1.5, 20, 20))
    parity.assert_parity(parity.run_both(rir, cols, cfg, seeds=np.arange(70) + 1))


def test_static_dense_data_free():
    rir, cols = configs.funnel().compile(True)
    n = 10
    A = np.random.default_rng(3).normal(size=(n, n)) * 0.2 + np.eye(n) * 1.5
    dense = api.DenseMassMatrix((A @ A.T).reshape(-1))
    r = parity.run_both(rir, cols, _cfg(30, 50, api.HMCSampler(4), api.DualAvgTuner(0.8), api.StaticMassMatrix(dense)), seeds=np.arange(33) + 1)
    parity.assert_parity(r, tol=1e-8)


def test_adaptive_dense_streamed_logistic_regression():
    model = configs.logreg(700, 4)
    rir, cols = model.compile(True)  # the reference's symbolic gradient outputs on both sides
    cfg = _cfg(20, 120, api.HMCSampler(3), api.DualAvgTuner(0.8), api.DenseMassMatrixTuner(30, 1.5, 10, 10))
    r = parity.run_both(rir, cols, cfg, seeds=np.arange(40) + 9)
    parity.assert_parity(r, tol=1e-7)
