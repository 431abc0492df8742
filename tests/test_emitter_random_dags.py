"""Randomised frozen DAGs (seeded): small expression trees over 1-3 parameters and, for half of them, observation columns,
built with the Python restatement of the reference's DAG builder (so they pass through its simplifier, Gradient.derive and
the Translator), then evaluated (a) by the oracle's interpreter and (b) by the CUDA source the emitter writes, compiled for
the host.  Symbolic-gradient mode must agree bit for bit; the emitter's own reverse mode within 1e-9 wherever the value is
finite.  Complements the fixed models of test_emitter_host.py and the RealTest mirror."""
import numpy as np
import pytest

from oracle.rainier_py.binding import OracleModel
from oracle.rainier_py.compute import Real, Vec, to_real
from oracle.rainier_py.core import Model, Normal
from rainier_b200 import abi, api

import host_emulation as he


def _random_expr(rng, leaves, depth):
    if depth == 0 or rng.random() < 0.2:
        return leaves[rng.integers(len(leaves))]
    k = rng.integers(9)
    a = _random_expr(rng, leaves, depth - 1)
    if k == 0:
        return a + _random_expr(rng, leaves, depth - 1)
    if k == 1:
        return a * _random_expr(rng, leaves, depth - 1)
    if k == 2:
        return a - _random_expr(rng, leaves, depth - 1) * float(rng.normal())
    if k == 3:
        return (a * 0.3).exp()
    if k == 4:
        return (a.abs() + 0.5).log()
    if k == 5:
        return (a * a + 1.0).pow(float(rng.choice([-1.0, -0.5, 0.5, 2.0, 3.0])))
    if k == 6:
        return Real.gt(a, 0.1, a * 2.0, _random_expr(rng, leaves, depth - 1))
    if k == 7:
        return a.max(_random_expr(rng, leaves, depth - 1))
    return a / ((_random_expr(rng, leaves, depth - 1)).abs() + 1.5)


@pytest.mark.parametrize("seed", range(24))
def test_random_dag(seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 4))
    streamed = seed % 2 == 1
    holder = {}

    def prior(t):
        holder["t"] = t
        return _random_expr(rng, list(t) + [to_real(0.7)], 3)

    params = Real.parameters(n, prior)
    model = Model.track_(list(params))
    if streamed:  # a likelihood that cannot be inlined: exp of a parameter x column mix, over 45 rows (ragged split sizes)
        xs = rng.normal(size=(45, 2))
        ys = rng.normal(size=45)
        t = holder["t"]
        lik = Model.observe(list(ys), Vec.from_([list(r) for r in xs]).map(
            lambda r: Normal((t[0] * r.at(0)).exp() + _random_expr(rng, [t[-1], r.at(1)], 2), 1.5)))
        model = model.merge(lik)
    rir, cols = model.compile(True)
    assert (len(cols) > 0) == streamed  # the likelihood really streams its rows
    om = OracleModel(rir, cols)
    q = rng.normal(size=(4, om.n)) * 0.8
    ref = om.density_batch(q)
    cm = api.CudaModel(rir, cols, device=-1)
    cfg = api.make_config(sampler=api.HMCSampler(1), gradientMode=abi.RN_GRAD_SYMBOLIC, backend=abi.RN_BACKEND_THREAD)
    sym, err = he.density(cm.emit_source(cfg), q, cols, cm, opt="-O0")
    assert err == 0
    assert np.all((sym == ref) | (np.isnan(sym) & np.isnan(ref))), "symbolic-gradient emission is not bit-identical to the interpreter"
    prir, pcols = model.compile(False)
    pm = api.CudaModel(prir, pcols, device=-1)
    adj, err = he.density(pm.emit_source(api.make_config(sampler=api.HMCSampler(1), backend=abi.RN_BACKEND_THREAD)), q, pcols, pm, opt="-O0")
    assert err == 0
    if not np.any(np.isfinite(ref[:, 1:]) & (ref[:, 1:] != 0.0)):
        pytest.skip("degenerate draw: the random expression does not depend on any parameter")
    fin = np.isfinite(ref) & np.isfinite(adj)
    assert np.array_equal(np.isfinite(ref[:, 0]), np.isfinite(adj[:, 0]))
    rel = np.abs(adj - ref)[fin] / np.maximum(np.abs(ref[fin]), 1e-6)
    assert rel.size == 0 or rel.max() < 1e-9, rel.max()
    if streamed:  # the same DAG on the warp-per-chain shape (rows across 32 emulated lanes, tree-ordered sums), both modes
        import os
        os.environ["RN_TMA"], os.environ["RN_WPC_K"] = "2", "1"
        try:
            for mdl, cc, gm in ((cm, cols, abi.RN_GRAD_SYMBOLIC), (pm, pcols, abi.RN_GRAD_AUTO)):
                src = mdl.emit_source(api.make_config(sampler=api.HMCSampler(1), gradientMode=gm, backend=abi.RN_BACKEND_WARP))
                w, err = he.density(src, q[:2], cc, mdl, opt="-O0")
                assert err == 0
                f2 = np.isfinite(ref[:2]) & np.isfinite(w)
                r2 = np.abs(w - ref[:2])[f2] / np.maximum(np.abs(ref[:2][f2]), 1e-6)
                assert r2.size == 0 or r2.max() < 1e-9, r2.max()
        finally:
            del os.environ["RN_TMA"], os.environ["RN_WPC_K"]
