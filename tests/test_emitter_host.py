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
