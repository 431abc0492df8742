"""Shared helpers of the GPU parity tests: run the same config/seeds through the CUDA path (C ABI) and through the
CPU oracle, and compare per chain."""
import numpy as np

from oracle.rainier_py.binding import OracleModel
from rainier_b200 import api


def rel_err(a, b, floor=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    same = (a == b) | both_nan
    d = np.abs(a - b) / np.maximum(np.abs(b), floor)
    d = np.where(same, 0.0, d)
    return float(np.nanmax(d)) if d.size else 0.0


def run_both(rir, cols, config, seeds=None, rng_states=None, rir_gpu=None, device=0, cols_gpu=None):
    """returns dict with gpu/oracle samples [chains][iters][n], traces [chains][iters_total][4], stats, mass"""
    cfg, keep = api.lower_config(config)
    dense = cfg.mass_tuner == 2 or (cfg.mass_tuner == 3 and cfg.static_matrix == 2)
    om = OracleModel(rir, cols)
    ref = om.sample(cfg, seeds=seeds, rng_states=rng_states, trace=True, dense_mass=dense)
    gm = api.CudaModel(rir_gpu if rir_gpu is not None else rir, cols_gpu if cols_gpu is not None else cols, device=device)
    s = api.CudaSampler(gm, config, seeds=seeds, rng_states=rng_states, trace=True)
    import torch

    n = gm.nVars
    d_samples = torch.empty((max(cfg.iterations, 1), n, s.chains), dtype=torch.float64, device="cuda:%d" % device)
    s.warmup(-1)
    s.run(cfg.iterations, d_samples.data_ptr())
    s.sync()
    gpu_samples = d_samples[: cfg.iterations].permute(2, 0, 1).contiguous().cpu().numpy()
    stats, mass = s.stats()
    trace = s.read_trace()
    s.close()
    gm.close()
    return {"gpu": gpu_samples, "ref": ref["samples"], "gpu_trace": trace, "ref_trace": ref["trace"], "gpu_stats": stats,
            "ref_stats": ref["stats"], "gpu_mass": mass, "ref_mass": ref["mass"]}


def assert_parity(r, tol=1e-9, check_mass=True):
    """Bit-exact accept decisions and trajectory lengths; samples / log-accept-probabilities / step sizes within tol."""
    gt, rt = r["gpu_trace"], r["ref_trace"]
    assert gt.shape == rt.shape
    assert np.array_equal(gt[:, :, 1], rt[:, :, 1]), "accept decisions differ: %d of %d" % (
        int(np.sum(gt[:, :, 1] != rt[:, :, 1])), gt[:, :, 1].size)
    assert np.array_equal(gt[:, :, 3], rt[:, :, 3]), "leapfrog step counts differ"
    assert rel_err(gt[:, :, 2], rt[:, :, 2]) < tol, "step sizes differ: %g" % rel_err(gt[:, :, 2], rt[:, :, 2])
    e = rel_err(r["gpu"], r["ref"])
    assert e < tol, "samples differ: max rel err %g" % e
    for g, o in zip(r["gpu_stats"], r["ref_stats"]):
        assert g.gradientEvaluations == o.gradient_evaluations
        assert g.leapfrogSteps == o.leapfrog_steps
        assert g.iterations == o.iterations
        assert g.accepted == o.accepted
        assert g.rng[0] == o.rng.seed48, "RNG streams diverged"
        assert rel_err(g.stepSize, o.step_size) < tol
        assert rel_err(g.acceptanceRatesMean, o.acceptance_rates_mean, 1e-6) < 1e-6
        assert rel_err(g.stepSizesMean, o.step_sizes_mean) < tol
        assert rel_err(g.gradsPerIterationMean, o.grads_per_iteration_mean) < tol
        assert rel_err(g.energyTransitions2, o.energy_transitions2, 1e-6) < 1e-6
        assert rel_err(g.energyVarianceRaw, o.energy_raw, 1e-6) < 1e-6
    if check_mass:
        assert rel_err(r["gpu_mass"], r["ref_mass"]) < 1e-7, rel_err(r["gpu_mass"], r["ref_mass"])
    return e
