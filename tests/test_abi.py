"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/rainier_cuda.h
declares, agrees with the ctypes mirror on struct sizes, emits + NVRTC-compiles sm_100a kernels without a device, and
fails loudly (no CPU fallback) when asked to execute without one."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rainier_b200 import abi, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rainier_cuda.h")).read()
    names = set(re.findall(r"\b(rn_[a-z_0-9]+)\s*\(", hdr))
    assert {"rn_model_create", "rn_density_batch", "rn_sample", "rn_last_error", "rn_model_destroy", "rn_emit_source"} <= names
    L = api.lib()
    for n in sorted(names):
        assert hasattr(L, n), "librainier_cuda.so does not export %s" % n


def test_struct_sizes_match():
    sizes = (C.c_int32 * 4)()
    api.lib().rn_abi_sizes(sizes)
    assert sizes[0] == C.sizeof(abi.Config) and sizes[1] == C.sizeof(abi.ChainStats) and sizes[2] == C.sizeof(abi.RngState)


def test_default_config_is_the_reference_default():
    c = abi.Config()
    api.lib().rn_config_default(C.byref(c))
    assert (c.iterations, c.warmup_iterations, c.stats_window) == (1000, 1000, 100)  # Sampler.scala:18-20
    assert c.sampler == abi.RN_SAMPLER_EHMC and c.max_steps == 1024 and c.min_steps == 1 and c.buf_size == 100 and c.p_count == 0.1
    assert c.step_size_tuner == abi.RN_STEP_DUAL_AVG and c.delta == 0.8
    assert c.mass_tuner == abi.RN_MASS_DIAGONAL and (c.initial_window_size, c.window_expansion, c.skip_first, c.skip_last) == (50, 1.5, 50, 50)


@pytest.mark.parametrize("name", ["funnel10", "eight_schools", "funnel10.primal", "eight_schools.primal"])
def test_emit_and_nvrtc_compile_without_device(name):
    rir = open(os.path.join(ROOT, "rainier_b200", "models", name + ".rir"), "rb").read()
    m = api.CudaModel(rir, [], device=-1)
    src = m.emit_source(api.SamplerConfig())
    assert "rn_density" in src and "rn_k_iter" in src
    cubin = m.emit_cubin(api.SamplerConfig())
    assert cubin[:4] == b"\x7fELF"


def test_no_cpu_fallback():
    rir = open(os.path.join(ROOT, "rainier_b200", "models", "funnel10.rir"), "rb").read()
    m = api.CudaModel(rir, [], device=-1)
    with pytest.raises(api.RainierCudaError) as e:
        m.density_batch(np.zeros((1, 10)))
    assert e.value.code == abi.RN_E_CUDA
    with pytest.raises(api.RainierCudaError) as e:
        m.sample(api.SamplerConfig(), seeds=[1])
    assert e.value.code == abi.RN_E_CUDA


def test_malformed_rir_is_rejected():
    with pytest.raises(api.RainierCudaError) as e:
        api.CudaModel(b"\0" * 64, [], device=-1)
    assert e.value.code == abi.RN_E_INVALID


def test_user_defined_sampler_cannot_be_lowered():
    class MySampler(api.Sampler):
        pass

    with pytest.raises(api.RainierCudaError) as e:
        api.lower_config(api.make_config(sampler=MySampler()))
    assert e.value.code == abi.RN_E_UNSUPPORTED


def test_scala_side_offsets():
    """scala/com/stripe/rainier/cuda/CudaSampling.scala hard-codes rn_config / rn_chain_stats field offsets (it cannot
    be compiled here); keep them pinned to the real struct layout."""
    src = open(os.path.join(ROOT, "scala", "com", "stripe", "rainier", "cuda", "CudaSampling.scala")).read()
    offs = dict((k, int(v)) for k, v in re.findall(r"private val (Off\w+) = (\d+)", src))
    expect = {"OffIterations": "iterations", "OffWarmup": "warmup_iterations", "OffStatsWindow": "stats_window",
              "OffSampler": "sampler", "OffNSteps": "n_steps", "OffMaxSteps": "max_steps", "OffMinSteps": "min_steps",
              "OffBufSize": "buf_size", "OffPCount": "p_count", "OffStepTuner": "step_size_tuner", "OffDelta": "delta",
              "OffStaticStep": "static_step_size", "OffMassTuner": "mass_tuner", "OffInitWindow": "initial_window_size",
              "OffExpansion": "window_expansion", "OffSkipFirst": "skip_first", "OffSkipLast": "skip_last"}
    for k, f in expect.items():
        assert offs[k] == getattr(abi.Config, f).offset, k
    assert "configSize() == %d" % C.sizeof(abi.Config) in src
    for field, off in (("gradient_evaluations", 0), ("iterations", 16), ("divergences", 20), ("energy_mean", 40),
                       ("energy_raw", 48), ("energy_transitions2", 56), ("energy_samples", 64), ("step_sizes_mean", 96),
                       ("acceptance_rates_mean", 104), ("grads_per_iteration_mean", 112), ("gradient_time_ns_mean", 144),
                       ("iteration_time_ns_mean", 152)):
        assert getattr(abi.ChainStats, field).offset == off
        assert "o + %d" % off in src or off == 0
