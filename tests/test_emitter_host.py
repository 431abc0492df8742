"""Emitter checks that need no GPU: the emitted rn_density(), compiled for the host (tests/host_emulation.py), against
the oracle -- bit-exact in symbolic-gradient mode (same operations in the same order), 1e-9 in adjoint mode."""
import numpy as np
import pytest

from oracle.rainier_py import configs, sbc_models
from oracle.rainier_py.binding import OracleModel
from rainier_b200 import abi, api

import host_emulation as he

CASES = {
    "funnel": lambda: configs.funnel(),
    "eight_schools": configs.eight_schools,
    "linreg_inlined": lambda: configs.linreg(200),
    "linreg_streamed": lambda: configs.linreg(300, covariates=5),
    "logreg": lambda: configs.logreg(400, 6),
    "SBCLaplace": lambda: sbc_models.build("SBCLaplace")[0],
    "SBCGamma": lambda: sbc_models.build("SBCGamma")[0],
    "SBCNegativeBinomial": lambda: sbc_models.build("SBCNegativeBinomial")[0],
}


@pytest.mark.parametrize("name", list(CASES))
def test_emitted_density_matches_oracle(name):
    model = CASES[name]()
    rir, cols = model.compile(True)
    om = OracleModel(rir, cols)
    q = np.random.default_rng(0).normal(size=(5, om.n)) * 0.7
    ref = om.density_batch(q)
    cm = api.CudaModel(rir, cols, device=-1)
    for gm, tol in ((abi.RN_GRAD_SYMBOLIC, 0.0), (abi.RN_GRAD_ADJOINT, 1e-9)):
        cfg = api.make_config(sampler=api.HMCSampler(1), gradientMode=gm, backend=abi.RN_BACKEND_THREAD)
        out, err = he.density(cm.emit_source(cfg), q, cols, cm)
        assert err == 0
        rel = np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1e-300))
        assert rel <= tol, (name, gm, rel)


def test_primal_rir_scatter_gradient_for_large_lookup_tables():
    """Poisson GLM with a 40-entry lookup table: the adjoint of Lookup is a scatter-add; it must
    agree with the reference's one-hot symbolic gradient evaluated by the oracle."""
    rir, cols = configs.poisson_glm(40, 640).compile(True)
    om = OracleModel(rir, cols)
    q = np.random.default_rng(1).normal(size=(3, om.n)) * 0.3
    ref = om.density_batch(q)
    prir, pcols = configs.poisson_glm(40, 640).compile(False)
    cm = api.CudaModel(prir, pcols, device=-1)
    src = cm.emit_source(api.make_config(sampler=api.HMCSampler(1), backend=abi.RN_BACKEND_THREAD))
    assert "+ k] +=" in src  # the scatter statement
    out, err = he.density(src, q, pcols, cm)
    assert err == 0
    assert np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1e-12)) < 1e-9


def test_warp_per_chain_source_shape():
    """CPU-side checks of what the emitter writes for a streamed model on the warp-per-chain shape (the device code itself
    is exercised by the GPU tests): TMA tile loop + per-warp fallback, no fdlibm pow in emitter-derived adjoints, CUDA libm
    only in row regions, K-warps reduction scratch, and the same source when emitted twice (cache key stability)."""
    rir, cols = configs.logreg(700, 4).compile(False)
    m = api.CudaModel(rir, cols, device=-1)
    cfg = api.make_config(sampler=api.HMCSampler(2), backend=abi.RN_BACKEND_WARP)
    src = m.emit_source(cfg)
    dens = src[src.index("// ---- emitted"):src.index("// rn_sampler_wpc.cuh --")]
    # logistic regression = dot products + elementwise code: the chain-batched DMMA path (per-warp TMA column blocks) ...
    assert "rn_dmma(" in dens and "rn_tma_load_raw(" in dens and "rn_mbar_wait_warp(" in dens and "rn_cta_bar(" in dens
    assert "#define RN_MMA_BARS 8" in src and dens.count("RN_DEVICE void rn_mma_e") == 1  # one copy serves the 8 unrolled observations
    dens_mma = dens
    import os
    os.environ["RN_MMA"] = "0"  # ... and, switched off, the CTA-shared tile pipeline of the rows-across-lanes body
    try:
        src = api.CudaModel(rir, cols, device=-1).emit_source(cfg)
    finally:
        del os.environ["RN_MMA"]
    dens = src[src.index("// ---- emitted"):src.index("// rn_sampler_wpc.cuh --")]
    assert "rn_dmma(" not in dens
    assert "rn_tma_load(" in dens and "rn_mbar_wait(" in dens and "rn_cta_bar(" in dens  # CTA-shared tiles
    assert "RN_LDG(rp +" in dens                                                          # independent per-warp path
    assert "rn_pow(" not in dens and "rn_pow_libm(" not in dens                           # d/dx x^-1 strength-reduced
    rows = dens[dens.index("// target 1"):]
    # row regions: the branch-free row functions (kernels without the DMMA path), never fdlibm; with the DMMA path: CUDA's libm
    assert "rn_row_exp(" in rows and "rn_row_log(" in rows and "rn_row_rcp(" in rows and "rn_exp(" not in rows and " exp(" not in rows
    mma_rows = dens_mma[dens_mma.index("RN_DEVICE void rn_mma_e"):]
    assert " exp(" in mma_rows and " log(" in mma_rows and "rn_row_exp(" not in mma_rows and "rn_exp(" not in mma_rows
    assert "RN_FENCE();" in rows                                                          # reverse sweep reloads columns
    assert "#define RN_TMA_STAGES" in src and "#define RN_WPC_K 1" in src
    assert m.emit_source(cfg) == api.CudaModel(rir, cols, device=-1).emit_source(cfg)
    os.environ["RN_WPC_K"] = "2"
    try:
        src2 = api.CudaModel(rir, cols, device=-1).emit_source(cfg)
    finally:
        del os.environ["RN_WPC_K"]
    assert "#define RN_WPC_K 2" in src2 and "double* red = scr +" in src2
    cub = api.CudaModel(rir, cols, device=-1).emit_cubin(cfg)  # NVRTC accepts it for sm_100a
    assert cub[:4] == b"\x7fELF"


def test_dense_structure_recognition():
    """rn_model_dot_structure: the parameter x column dot products of the streamed row bodies (the Translator's fold of a
    `Line` with column coefficients, compute/Translator.scala:91-125) -- one per observation, d terms each, for a
    regression on d covariates; none in a data-free model or where coefficients are not columns.  Groundwork of the
    chain-batched DMMA contraction: this is the detector of "where the DAG really is a dense mat-vec"."""
    for nobs, d in ((700, 4), (160, 37)):
        for with_gradient in (True, False):
            rir, cols = configs.logreg(nobs, d).compile(with_gradient)
            st = api.CudaModel(rir, cols, device=-1).dot_structure()
            assert st["dots_per_gradient"] == nobs and st["dot_fmas_per_gradient"] == nobs * d
            assert st["longest_dot"] == d and st["distinct_dots"] == 9  # 8 unrolled splits (Model.scala:98-132) + the init block
    rir, cols = configs.eight_schools().compile(True)
    assert api.CudaModel(rir, cols, device=-1).dot_structure()["distinct_dots"] == 0
    model, real, rng, _ = sbc_models.build("SBCLaplace")  # streamed, but no parameter x column products
    rir, cols = model.compile(True)
    assert len(cols) > 0 and api.CudaModel(rir, cols, device=-1).dot_structure()["dots_per_gradient"] == 0


def test_separability_analysis():
    """rn_model_separable_structure: which streamed targets have a row sum of the form sum_k S_k * p_k(parameters) -- what
    the reference's inliner folds into constants on the JVM (compute/Target.scala:136-207).  A Gaussian regression on 5
    covariates is NOT inlined by the reference (6 additive terms: 21 >= 20, compute/LogLineOps.scala:43-66) and streams
    its rows, yet it is separable (its atoms are the entries of X^T X, X^T y, y^T y per unrolled split); a logistic
    regression and a Laplace likelihood are not (exp / abs of a parameter x column mix).  Groundwork of the device-side
    inliner."""
    rir, cols = configs.linreg(400, covariates=3).compile(False)
    assert not cols  # the reference inlines this one itself: nothing is streamed
    assert api.CudaModel(rir, cols, device=-1).separable_structure()["streamed_targets"] == 0
    for with_gradient in (True, False):
        rir, cols = configs.linreg(400, covariates=5).compile(with_gradient)
        st = api.CudaModel(rir, cols, device=-1).separable_structure()
        assert st["streamed_targets"] == 2 and st["separable_targets"] == 2 and st["rows_removed"] == 49 + 8
        assert 8 * 28 <= st["atoms"] <= 9 * 28  # (5 covariates + y + 1)(..+1)/2 = 28 products per unrolled observation
        rir, cols = configs.logreg(700, 4).compile(with_gradient)
        st = api.CudaModel(rir, cols, device=-1).separable_structure()
        assert st["streamed_targets"] == 2 and st["separable_targets"] == 0
    model, real, rng, _ = sbc_models.build("SBCLaplace")
    assert api.CudaModel(*model.compile(True), device=-1).separable_structure()["separable_targets"] == 0
