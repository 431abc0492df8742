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


def test_adaptive_dense_streamed_thread_shape_is_bit_exact():
    """Streamed rows + DualAvg + dense mass adaptation, the reference's sequential row order (thread per chain): accept
    decisions, step counts and RNG positions identical to the oracle through the whole adaptive warmup."""
    rir, cols = configs.logreg(700, 4).compile(True)  # the reference's symbolic gradient outputs on both sides
    cfg = api.make_config(iterations=20, warmupIterations=120, sampler=api.HMCSampler(3), stepSizeTuner=api.DualAvgTuner(0.8),
                          massMatrixTuner=api.DenseMassMatrixTuner(30, 1.5, 10, 10), backend=abi.RN_BACKEND_THREAD)
    parity.assert_parity(parity.run_both(rir, cols, cfg, seeds=np.arange(40) + 9), tol=1e-9)


def test_adaptive_dense_streamed_logistic_regression():
    """The same run on the warp-per-chain shape.  Rows are summed as a tree there (32 per-lane partials + butterfly), so a
    density differs from the reference's sequential sum in the last bits (~1e-13), and an ADAPTIVE warmup amplifies that:
    measured on B200 (scripts/r2/diag_wpc_dense.py, profiles/r2_diag_wpc_dense_v1.txt) the log-acceptance error grows
    smoothly 1e-10 (iteration 9) -> 1e-8 (20) -> 1e-6 (39) -> O(1) (60), for identity, diagonal and dense matrices alike,
    while the thread shape stays at exactly 0.  Nothing that reorders a floating-point sum can follow the reference bit
    for bit through a chaotic map, so what is asserted here is what the shape can promise:
      (1) the early warmup, before the amplification matters, agrees decision for decision and to 1e-6;
      (2) after the full adaptive run the two ensembles are the same distribution: adapted covariances, step sizes,
          acceptance rates and posterior means agree within Monte-Carlo error over 256 chains."""
    model = configs.logreg(700, 4)
    rir, cols = model.compile(True)
    cfg = _cfg(20, 120, api.HMCSampler(3), api.DualAvgTuner(0.8), api.DenseMassMatrixTuner(30, 1.5, 10, 10))
    r = parity.run_both(rir, cols, cfg, seeds=np.arange(256) + 9)
    gt, rt = r["gpu_trace"], r["ref_trace"]
    early = slice(0, 12)
    assert np.array_equal(gt[:, early, 1], rt[:, early, 1]) and np.array_equal(gt[:, early, 3], rt[:, early, 3])
    assert parity.rel_err(gt[:, early, 2], rt[:, early, 2]) < 1e-6
    fin = np.isfinite(gt[:, early, 0]) & np.isfinite(rt[:, early, 0])
    assert np.max(np.abs(gt[:, early, 0][fin] - rt[:, early, 0][fin])) < 1e-6
    # (2) ensemble agreement
    n = 4
    gm, rm = r["gpu_mass"].reshape(-1, n, n), r["ref_mass"].reshape(-1, n, n)
    se = rm.std(axis=0) / np.sqrt(rm.shape[0]) + gm.std(axis=0) / np.sqrt(gm.shape[0])
    assert np.all(np.abs(gm.mean(axis=0) - rm.mean(axis=0)) < 5 * se + 1e-12), "adapted covariance matrices differ in distribution"
    gs = np.array([s.stepSize for s in r["gpu_stats"]]); rs = np.array([s.step_size for s in r["ref_stats"]])
    assert abs(gs.mean() - rs.mean()) < 5 * (gs.std() + rs.std()) / np.sqrt(len(gs))
    ga = np.array([s.accepted / s.iterations for s in r["gpu_stats"]]); ra = np.array([s.accepted / s.iterations for s in r["ref_stats"]])
    assert abs(ga.mean() - ra.mean()) < 5 * (ga.std() + ra.std()) / np.sqrt(len(ga)) + 1e-3
    gq, rq = r["gpu"].reshape(-1, n), r["ref"].reshape(-1, n)
    assert np.all(np.abs(gq.mean(axis=0) - rq.mean(axis=0)) < 0.1 * rq.std(axis=0))
