"""
SURVEY.md 8f-2 -- posterior-predictive requirements (Trace.predict -> Generator.prepare -> Compiler.compile + CompiledFunction.output,
rainier-core/.../core/Trace.scala:34-41, core/Generator.scala:59-94, rainier-compute/.../compute/Compiler.scala:22-30), CPU side:

  * the oracle of this row (oracle rno_function_* + the Python restatement of Generator.prepare / Trace.predict) is PINNED
    to the reference's golden vectors: SBCModel.scala:37 produces every goldset with `model.sample(...).predict(real)`,
    i.e. through exactly this path;
  * the emitted CUDA source of the function flavour (rn_function() + rn_k_eval, compiled for the host) is bit-identical
    to the oracle, in both addressing modes of rn_function_eval_device;
  * the batched predict consumes the RNG exactly like the reference's per-draw loop;
  * container validation, lookup errors, no CPU fallback.
"""
import json
import os
import struct

import numpy as np
import pytest

from oracle.rainier_py import configs, sbc_models
from oracle.rainier_py.binding import OracleError, OracleFunction, OracleModel, ScalaRNG
from oracle.rainier_py.compute import Real, compile_function_rir
from oracle.rainier_py.core import Cauchy, Generator, Model, Normal, to_generator
from rainier_b200 import abi, api

import host_emulation

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sbc_goldsets.json")))


def schools():
    """rainier-benchmark/.../bench/stan/EightSchools.scala:9-24, keeping the latent handles"""
    return configs.eight_schools_parts()


@pytest.mark.parametrize("name", sbc_models.ENABLED)
def test_goldset_through_compiled_predict(name):
    """SBCModel.scala:31-39 end to end the reference's way: predict(real) = Generator.real -> prepare -> compiled
    function -> CompiledFunction.output (not the tree-walking Evaluator); 1e-10 as in SBCTest.scala:7-15."""
    gold = GOLD["models"][name]["goldset"]
    model, real, rng, _ = sbc_models.build(name, GOLD["seed"], GOLD["synthetic_samples"])
    rir, cols = model.compile(True)
    res = OracleModel(rir, cols).sample(sbc_models.sbc_config(len(gold), GOLD["warmup"]), rng_states=[rng.rand.state()])
    out = to_generator(real).predict(model.parameters, res["samples"][0], rng, OracleFunction)
    assert len(out) == len(gold)
    for a, b in zip(out, gold):
        assert abs((a - b) / b) < 1e-10
    # and the emitted CUDA source of that function, compiled for the host, gives the same bits
    frir = compile_function_rir(model.parameters, [real])
    src = api.CudaFunction(frir, device=-1).emit_source()
    emu, err = host_emulation.eval_function(src, res["samples"][0], 1)
    assert err == 0 and np.array_equal(emu[:, 0], np.array(out))


def _derived(mu, tau, thetas):
    t = [thetas.at(i) for i in range(8)]
    return configs.eight_schools_derived(mu, tau, thetas) + [Real.sum(t) / 8.0, (t[2] * t[3]).exp(), mu, Real.zero + 3.5]


def test_emitted_function_bit_identical_to_oracle_both_layouts():
    model, mu, tau, thetas, _ = schools()
    params = model.parameters
    reals = _derived(mu, tau, thetas)
    rir = compile_function_rir(params, reals)
    of = OracleFunction(rir)
    cf = api.CudaFunction(rir, device=-1)
    assert (cf.nInputs, cf.nOutputs) == (10, len(reals)) == (of.nInputs, of.nOutputs)
    src = cf.emit_source()
    iters, chains = 7, 45  # 315 points: more than one CTA of 128, a ragged tail, a grid-stride second pass with grid=2
    draws = np.random.default_rng(3).normal(size=(iters, 10, chains)) * 1.3  # [iteration][n][chain], as rn_sampler_run
    rows = np.ascontiguousarray(draws.transpose(2, 0, 1)).reshape(-1, 10)    # predict's order: chain-major
    ref = of(rows)
    out_rows, err = host_emulation.eval_function(src, rows, len(reals), grid=2)
    assert err == 0 and np.array_equal(out_rows, ref)
    out_s, err = host_emulation.eval_function(src, draws, len(reals), layout="sampler", iterations=iters, chains=chains, grid=2)
    assert err == 0 and np.array_equal(out_s.reshape(-1, len(reals)), ref)
    # constants / duplicated outputs / inputs as outputs are stored too
    assert np.all(ref[:, -1] == 3.5) and np.array_equal(ref[:, 0], ref[:, -2])
    # non-finite inputs are values, not errors (NaN / inf propagate like in the JVM)
    bad = rows[:4].copy()
    bad[0, 0], bad[1, 1], bad[2, 2] = np.nan, np.inf, -np.inf
    o, err = host_emulation.eval_function(src, bad, len(reals))
    assert err == 0 and np.array_equal(o, of(bad), equal_nan=True)


def test_fast_math_function_within_tolerance():
    model, mu, tau, thetas, _ = schools()
    reals = _derived(mu, tau, thetas)
    rir = compile_function_rir(model.parameters, reals)
    rows = np.random.default_rng(5).normal(size=(64, 10))
    ref = OracleFunction(rir)(rows)
    out, err = host_emulation.eval_function(api.CudaFunction(rir, device=-1, fast=True).emit_source(), rows, len(reals), fast=True)
    assert err == 0
    np.testing.assert_allclose(out, ref, rtol=1e-9, atol=1e-12)  # north_star's 1e-9 relative; FMA contraction moves last bits of cancelling sums


def test_batched_predict_consumes_rng_like_the_per_draw_loop():
    """posterior-predictive draws y_i ~ Normal(theta_i, sigma_i): a From generator that consumes the RNG.  The batched
    predict (all requirement values first, then Generator.get per draw) must return what the reference's per-draw closure
    (Generator.prepare's fn(array), core/Generator.scala:76-93) returns on the same RNG stream."""
    model, mu, tau, thetas, sigmas = schools()
    params = model.parameters
    gen = Generator.traverse([to_generator(Normal(thetas.at(i), sigmas[i])) for i in range(8)]).zip(to_generator(tau))
    draws = np.random.default_rng(11).normal(size=(40, 10))
    r1, r2 = ScalaRNG(77), ScalaRNG(77)
    fn = gen.prepare(params, r1, OracleFunction)
    per_draw = [fn(a) for a in draws]
    batched = gen.predict(params, draws, r2, OracleFunction)
    assert per_draw == batched
    assert r1.rand.state().seed48 == r2.rand.state().seed48
    # requirements: 8 thetas + tau, deduplicated (tau appears in every theta's Normal and on its own)
    assert len(gen.reqs()) <= Generator.MaxRequirements and len(set(gen.reqs())) == len(gen.reqs())
    # a generator without requirements never compiles anything
    const = Generator.constant(5).predict(params, draws, r1, lambda rir: pytest.fail("compiled a requirement-free generator"))
    assert const == [5] * 40


def test_lookup_function_and_out_of_range_error():
    from oracle.rainier_py.compute import lookup_apply
    model, mu, tau, thetas, _ = schools()
    params = model.parameters
    look = lookup_apply(mu.abs(), [thetas.at(i) for i in range(8)])  # Lookup(index, table), compute/Real.scala:287-308
    rir = compile_function_rir(params, [look, tau])
    of = OracleFunction(rir)
    src = api.CudaFunction(rir, device=-1).emit_source()
    rows = np.random.default_rng(9).normal(size=(50, 10)) * 0.2  # |mu| = |5 q0| < 8
    rows[:, 0] = np.abs(rows[:, 0])
    ok = rows[np.abs(rows[:, 0] * 5.0) < 7.9]
    out, err = host_emulation.eval_function(src, ok, 2)
    assert err == 0 and np.array_equal(out, of(ok))
    bad = ok[:3].copy()
    bad[1, 0] = 100.0  # index 500: outside the 8-entry table -> NullPointerException in the reference
    out, err = host_emulation.eval_function(src, bad, 2)
    assert err & 1
    with pytest.raises(OracleError):
        of(bad)


def test_function_container_validation_and_no_cpu_fallback():
    model, mu, tau, thetas, _ = schools()
    rir = compile_function_rir(model.parameters, [mu, tau])
    f = api.CudaFunction(rir, device=-1)
    with pytest.raises(api.RainierCudaError) as e:  # no device: emit/compile only, evaluation fails loudly
        f(np.zeros((1, 10)))
    assert e.value.code == abi.RN_E_CUDA
    with pytest.raises(api.RainierCudaError) as e:  # sample + predict in one call: same rule
        api.CudaModel(*model.compile(True), device=-1).sample_predict(f, api.SamplerConfig(iterations=2, warmupIterations=2), seeds=[1])
    assert e.value.code == abi.RN_E_CUDA
    # a model container is not a function container and vice versa
    mrir, cols = model.compile(True)
    with pytest.raises(api.RainierCudaError) as e:
        api.CudaFunction(mrir, device=-1)
    assert e.value.code == abi.RN_E_INVALID
    with pytest.raises(api.RainierCudaError):
        api.CudaModel(rir, [], device=-1)
    # truncated / corrupted containers
    for cut in (10, 40, len(rir) - 4):
        with pytest.raises(api.RainierCudaError):
            api.CudaFunction(rir[:cut], device=-1)
    hdr = list(struct.unpack("<8I", rir[:32]))
    hdr[3] += 1  # n_inputs != n_params
    with pytest.raises(api.RainierCudaError):
        api.CudaFunction(struct.pack("<8I", *hdr) + rir[32:], device=-1)
    # the function flavour assembles for sm_100a through NVRTC
    cubin = f.emit_cubin()
    assert cubin[:4] == b"\x7fELF" and f.op_counts()["flops"] >= 2


@pytest.mark.parametrize("seed", range(16))
def test_random_functions(seed):
    """Randomised requirement lists (seeded): 1-6 random expression trees over 1-4 parameters -- through the reference's
    simplifier and Translator (restated), then (a) the oracle's interpreter, (b) the emitted CUDA source compiled for the
    host: bit-identical; and (c) the tree-walking Evaluator (compute/Evaluator.scala) agrees to rounding -- the reference's
    RealTest asserts exactly that triangle for single expressions (rainier-test/.../compute/RealTest.scala:24-40)."""
    from oracle.rainier_py.compute import Evaluator, to_real
    from test_emitter_random_dags import _random_expr
    rng = np.random.default_rng(7000 + seed)
    n, m = int(rng.integers(1, 5)), int(rng.integers(1, 7))
    holder = {}

    def keep(t):
        holder["t"] = list(t)
        return Real.sum(list(t))

    params = Real.parameters(n, keep)
    model = Model.track_(list(params))
    leaves = holder["t"] + [to_real(0.7)]
    reals = [_random_expr(rng, leaves, 3) for _ in range(m)]
    if seed % 4 == 0:
        reals.append(reals[0])  # a duplicated requirement keeps its own output slot
    plist = model.parameters
    rir = compile_function_rir(plist, reals)
    x = rng.normal(size=(9, len(plist))) * 0.8
    ref = OracleFunction(rir)(x)
    out, err = host_emulation.eval_function(api.CudaFunction(rir, device=-1).emit_source(), x, len(reals))
    assert err == 0 and np.array_equal(out, ref, equal_nan=True)
    for p in range(3):
        ev = Evaluator({q: float(v) for q, v in zip(plist, x[p])})
        for j, r in enumerate(reals):
            v = ev.toDouble(r)
            if np.isfinite(v) and np.isfinite(ref[p, j]):
                assert abs(v - ref[p, j]) <= 1e-9 * max(abs(v), 1e-6), (p, j, v, ref[p, j])
