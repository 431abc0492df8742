"""
Device-side inlining (rainier_b200/csrc/rn_inline.cpp), host halves: the separability plan and the rewrite of
the container -- checked WITHOUT a device by doing the one device step (row sums of the column-only monomials) with numpy over
the oracle's evaluation of the plan's own function-flavour program.
  * README linear regression, 3 covariates: the reference's simplifier does NOT expand it in its streamed form here (we hand
    the non-inlined DAG), the rewrite must reproduce the reference's own inlined, data-free model (compute/Target.scala:136-207,
    compute/PartialEvaluator.scala:86-97 as restated in oracle/rainier_py/compute.py) to 1e-9;
  * Gaussian regression on 5 covariates: the reference streams it (compute/LogLineOps.scala:43-66 stops expanding squares at 5
    additive terms); the rewrite makes it data-free and must agree with the oracle on the streamed form;
  * logistic / Poisson regression: not separable (exp of a parameter x column mix) -> left alone.
"""
import numpy as np
import pytest

from oracle.rainier_py import compute, configs
from oracle.rainier_py.binding import OracleFunction, OracleModel
from rainier_b200 import api


def _streamed(model):
    """the model's primal container WITHOUT the reference's inlining step (what the Scala side would send)"""
    keep = compute.inlinable
    compute.inlinable = lambda real: False
    try:
        return model.compile(False)
    finally:
        compute.inlinable = keep


def _inline_on_host(rir, cols, n_params):
    plan = api.inline_plan(rir)
    sums = []
    info = []
    for target, n_mono, frir in plan:
        f = OracleFunction(frir)
        # the function's inputs are the target's columns, in column order: find them by length bookkeeping of the container
        info.append((target, n_mono, f.nInputs))
        sums.append((f, n_mono))
    return plan, sums


@pytest.mark.parametrize("name", ["linreg3", "linreg5"])
def test_rewrite_matches_the_reference(name):
    model = configs.linreg(400) if name == "linreg3" else configs.linreg(400, covariates=5)
    rir_ref, cols_ref = model.compile(True)       # the reference's own result (inlined for 3 covariates, streamed for 5)
    model2 = configs.linreg(400) if name == "linreg3" else configs.linreg(400, covariates=5)
    srir, scols = _streamed(model2)               # streamed primal container
    assert len(scols) > 0
    plan = api.inline_plan(srir)
    assert len(plan) >= 1
    # columns of target t: consecutive in input order; recover each target's column slice from the function's input count
    om = OracleModel(srir, scols)
    n = om.n
    pos, sums = 0, []
    by_target = sorted(plan, key=lambda p: p[0])
    for target, n_mono, frir in by_target:
        f = OracleFunction(frir)
        block = scols[pos:pos + f.nInputs]
        pos += f.nInputs
        vals = f(np.stack(block, axis=1))          # [rows][monomials]
        assert vals.shape[1] == n_mono
        sums.append(vals.sum(axis=0))
    assert pos == len(scols), "every streamed target of a Gaussian regression is separable"
    order = {t: i for i, (t, _, _) in enumerate(by_target)}
    flat = np.concatenate([sums[order[t]] for t, _, _ in plan])
    new_rir = api.inline_apply(srir, flat)
    q = np.random.default_rng(1).normal(size=(6, n)) * 0.4
    q[:, 0] = np.abs(q[:, 0]) + 0.2
    got = OracleModel(new_rir, scols).density_batch(q)      # data-free now: the columns are not read
    ref = OracleModel(rir_ref, cols_ref).density_batch(q)
    streamed = om.density_batch(q)
    assert np.max(np.abs(got - streamed) / np.maximum(np.abs(streamed), 1e-9)) < 1e-9
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-9)) < 1e-9
    # and it really is data-free: emitted for the thread-per-chain shape, no row loop
    src = api.CudaModel(new_rir, scols, device=-1).emit_source(api.make_config(sampler=api.HMCSampler(2)))
    assert "for (long long row" not in src


@pytest.mark.parametrize("name", ["logreg", "poisson"])
def test_nonlinear_likelihoods_are_left_streamed(name):
    model = configs.logreg(300, 4) if name == "logreg" else configs.poisson_glm(7, 300)
    prir, pcols = model.compile(False)
    assert api.inline_plan(prir) == []
