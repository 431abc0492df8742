"""
GPU parity of SURVEY.md 8f-2 (posterior-predictive requirements: Trace.predict -> Generator.prepare -> Compiler.compile +
CompiledFunction.output, core/Trace.scala:34-41, core/Generator.scala:59-94) through the C ABI (rn_function_*):
the reference's golden vectors end to end on the device, bit-equality with the oracle, both addressing modes, errors.
(Named test_zz_* so that it runs after the files of the hot path proper.)
"""
import json
import os

import numpy as np
import pytest

from oracle.rainier_py import sbc_models
from oracle.rainier_py.binding import OracleFunction, ScalaRNG
from oracle.rainier_py.compute import compile_function_rir, lookup_apply
from oracle.rainier_py.core import Generator, Normal, to_generator
from rainier_b200 import abi, api

from test_function_host import _derived, schools

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sbc_goldsets.json")))


@pytest.mark.parametrize("name", ["SBCUniformNormal", "SBCGamma", "SBCLargePoisson"])
def test_reference_goldsets_sampled_and_predicted_on_gpu(name):
    """SBCModel.scala:31-39 with BOTH halves on the device: rn_sample draws the chain, rn_function_eval evaluates
    predict(real) for all draws in one launch; 1e-10 against the reference's goldset as in SBCTest.scala:7-15."""
    gold = GOLD["models"][name]["goldset"]
    model, real, rng, _ = sbc_models.build(name, GOLD["seed"], GOLD["synthetic_samples"])
    rir, cols = model.compile(True)
    cfg = api.make_config(iterations=len(gold), warmupIterations=GOLD["warmup"], sampler=api.HMCSampler(1),
                          stepSizeTuner=api.DualAvgTuner(0.8), massMatrixTuner=api.IdentityMassMatrixTuner())
    tr = api.CudaModel(rir, cols).sample(cfg, rng_states=[rng.rand.state()])
    out = to_generator(real).predict(model.parameters, tr.chains[0], rng, api.CudaFunction)
    assert len(out) == len(gold)
    for a, b in zip(out, gold):
        assert abs((a - b) / b) < 1e-10
    f = api.CudaFunction(compile_function_rir(model.parameters, [real]))
    assert np.array_equal(tr.requirements(f)[:, 0], np.array(out))


@pytest.mark.parametrize("count", [1, 127, 128, 129, 4097, 300000])
def test_function_bit_identical_to_oracle(count):
    model, mu, tau, thetas, _ = schools()
    reals = _derived(mu, tau, thetas)
    rir = compile_function_rir(model.parameters, reals)
    rows = np.random.default_rng(count).normal(size=(count, 10)) * 1.3
    if count > 3:
        rows[0, 0], rows[1, 1], rows[2, 2] = np.nan, np.inf, -np.inf  # values, not errors
    f = api.CudaFunction(rir)
    got = f(rows)
    assert np.array_equal(got, OracleFunction(rir)(rows), equal_nan=True)
    assert f.launches() >= 1
    assert f(np.zeros((0, 10))).shape == (0, len(reals))  # empty batch


def test_device_resident_draws_in_sampler_layout():
    """rn_sampler_run leaves draws as [iteration][n][chain]; rn_function_eval_device reads them in place and writes
    [chain][iteration][m] = the order of Trace.predict (chains.flatMap(_.map(fn)))."""
    import torch
    model, mu, tau, thetas, sigmas = schools()
    reals = _derived(mu, tau, thetas)
    m = len(reals)
    rir, cols = model.compile(True)
    iters, chains = 37, 333
    cfg = api.SamplerConfig(iterations=iters, warmupIterations=120)
    cm = api.CudaModel(rir, cols)
    s = api.CudaSampler(cm, cfg, seeds=np.arange(chains) + 5)
    d = torch.empty((iters, cm.nVars, chains), dtype=torch.float64, device="cuda")
    s.warmup(-1)
    s.run(iters, d.data_ptr())
    s.sync()
    frir = compile_function_rir(model.parameters, reals)
    f = api.CudaFunction(frir)
    out = torch.full((chains, iters, m), float("nan"), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    f.eval_device(d.data_ptr(), iters, chains, out.data_ptr())
    f.sync()
    draws = d.permute(2, 0, 1).contiguous().cpu().numpy()  # [chain][iteration][n]
    ref = OracleFunction(frir)(draws.reshape(-1, cm.nVars)).reshape(chains, iters, m)
    assert np.array_equal(out.cpu().numpy(), ref)
    # the same through host buffers and through the row layout on the device
    tr = api.Trace(draws, None, None)
    assert np.array_equal(tr.requirements(f).reshape(chains, iters, m), ref)
    rows = torch.from_numpy(draws.reshape(-1, cm.nVars)).cuda()
    out2 = torch.empty((chains * iters, m), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    f.eval_device(rows.data_ptr(), iters, chains, out2.data_ptr(), layout=abi.RN_LAYOUT_ROWS)
    f.sync()
    assert np.array_equal(out2.cpu().numpy().reshape(chains, iters, m), ref)
    # posterior-predictive draws with an RNG-consuming generator: identical to the CPU oracle's function on one stream
    gen = Generator.traverse([to_generator(Normal(thetas.at(i), sigmas[i])) for i in range(8)])
    a = gen.predict(model.parameters, draws.reshape(-1, cm.nVars)[:200], ScalaRNG(9), api.CudaFunction)
    b = gen.predict(model.parameters, draws.reshape(-1, cm.nVars)[:200], ScalaRNG(9), OracleFunction)
    assert a == b
    s.close()


def test_function_lookup_error_and_fast_math():
    model, mu, tau, thetas, _ = schools()
    look = lookup_apply(mu.abs(), [thetas.at(i) for i in range(8)])
    rir = compile_function_rir(model.parameters, [look, tau])
    f = api.CudaFunction(rir)
    rows = np.abs(np.random.default_rng(2).normal(size=(500, 10))) * 0.2
    assert np.array_equal(f(rows), OracleFunction(rir)(rows))
    bad = rows.copy()
    bad[321, 0] = 100.0
    with pytest.raises(api.RainierCudaError) as e:
        f(bad)
    assert e.value.code == abi.RN_E_LOOKUP
    assert np.array_equal(f(rows), OracleFunction(rir)(rows))  # the flag is cleared; the handle stays usable
    reals = _derived(mu, tau, thetas)
    rir2 = compile_function_rir(model.parameters, reals)
    x = np.random.default_rng(4).normal(size=(1000, 10))
    np.testing.assert_allclose(api.CudaFunction(rir2, fast=True)(x), OracleFunction(rir2)(x), rtol=1e-9, atol=1e-12)  # north_star's 1e-9 relative; FMA contraction moves last bits of cancelling sums


def test_sample_predict_in_one_call():
    """object Model.sample(t, config) (core/Model.scala:56-63) = model.sample(config).predict(gen): rn_sample_predict keeps
    the draws on the device and returns only the requirement values -- identical to rn_sample followed by rn_function_eval."""
    model, mu, tau, thetas, _ = schools()
    reals = _derived(mu, tau, thetas)
    rir, cols = model.compile(True)
    cm = api.CudaModel(rir, cols)
    f = api.CudaFunction(compile_function_rir(model.parameters, reals))
    cfg = api.SamplerConfig(iterations=41, warmupIterations=150)
    seeds = np.arange(777) + 11
    pred, tr = cm.sample_predict(f, cfg, seeds=seeds)
    full = cm.sample(cfg, seeds=seeds)
    assert pred.shape == (777, 41, len(reals))
    assert np.array_equal(pred.reshape(-1, len(reals)), full.requirements(f))
    assert np.array_equal(pred, OracleFunction(f._rir)(full.chains.reshape(-1, cm.nVars)).reshape(pred.shape))
    assert [s.gradientEvaluations for s in tr.stats] == [s.gradientEvaluations for s in full.stats]
    assert np.array_equal(tr.mass, full.mass)
    # page-locked destination: one DMA
    buf = api.PinnedBuffer((777, 41, len(reals)))
    pred2, _ = cm.sample_predict(f, cfg, seeds=seeds, out=buf.array)
    assert np.array_equal(pred2, pred)
    buf.close()
