#!/usr/bin/env python
"""
scripts/bench_function.py -- measurement of SURVEY.md 8f-2 (posterior-predictive requirements, rn_function_*) on a GPU box.

Workload: the headline run's draws -- Neal's funnel is replaced here by eight schools (n = 10 parameters) because its
predict has something to compute: 151 552 chains x 100 iterations left on the device by rn_sampler_run in the sampler's
own [iteration][n][chain] layout; the function evaluates m = 12 derived quantities (mu, tau, the 8 thetas, two
transcendental ones) per draw.  One launch of rn_k_eval per step.

  value    : draws/s, device-resident (CUDA events on the function's stream), L2 flushed between steps
  roofline : HBM -- algorithmic bytes = (n + m) * 8 per draw (read the draw once, write its m values once) x draws per
             launch / measured launch time, against MEASURED_PEAKS.json hbm_gbs
  e2e      : the same through rn_function_eval with HOST buffers ([count][n] in, [count][m] out; H2D + D2H inside)
  cpu_baseline : the oracle's rno_function_eval on one host core (the reference's per-draw CompiledFunction.output loop is
             single-threaded too, core/Generator.scala:76-93), bounded sample
Prints one JSON line.  Not part of the product; models come from the committed fixtures, oracle/ is used as checker and CPU baseline only.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=151552)
    ap.add_argument("--iterations", type=int, default=100)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--fast", action="store_true")
    args = ap.parse_args()
    import torch

    from rainier_b200 import api

    # committed fixtures (oracle/make_fixtures.py): the frozen DAG of eight schools and its 12 derived quantities as a function
    mdir = os.path.join(ROOT, "rainier_b200", "models")
    rir, cols = open(os.path.join(mdir, "eight_schools.rir"), "rb").read(), []
    frir = open(os.path.join(mdir, "eight_schools.derived.fn.rir"), "rb").read()
    n, m = 10, 12
    chains, iters = args.chains, args.iterations
    count = chains * iters
    cm = api.CudaModel(rir, cols)
    cfg = api.make_config(iterations=iters, warmupIterations=50, sampler=api.HMCSampler(5), stepSizeTuner=api.StaticStepSize(0.1),
                          massMatrixTuner=api.IdentityMassMatrixTuner())
    s = api.CudaSampler(cm, cfg, seeds=np.arange(chains) + 1000)
    d = torch.empty((iters, n, chains), dtype=torch.float64, device="cuda")
    s.warmup(-1)
    s.run(iters, d.data_ptr())
    s.sync()
    f = api.CudaFunction(frir, fast=args.fast)
    out = torch.empty((chains, iters, m), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    f.eval_device(d.data_ptr(), iters, chains, out.data_ptr())  # compiles + loads
    f.sync()
    st = torch.cuda.ExternalStream(f.stream())
    times = []
    for k in range(args.warmup + args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        f.eval_device(d.data_ptr(), iters, chains, out.data_ptr())
        e1.record(st)
        f.sync()
        if k >= args.warmup:
            times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times))
    alg = (n + m) * 8.0 * count
    try:
        peak, src = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        peak, src = 6650.0, "fallback"
    # parity spot check against the oracle (first chain block) -- the oracle is the checker / CPU baseline only
    from oracle.rainier_py.binding import OracleFunction
    draws = d[:, :, :64].permute(2, 0, 1).contiguous().cpu().numpy().reshape(-1, n)
    ref = OracleFunction(frir)(draws).reshape(64, iters, m)
    same = bool(np.array_equal(out[:64].cpu().numpy(), ref)) if not args.fast else None
    # e2e: host buffers
    hx = d.permute(2, 0, 1).contiguous().cpu().numpy().reshape(-1, n)
    t_e2e = []
    for k in range(3):
        t0 = time.perf_counter()
        f(hx)
        t_e2e.append(time.perf_counter() - t0)
    # one call for sample + predict (object Model.sample(t, config)): only the m requirement values cross PCIe
    pb = api.PinnedBuffer((chains, iters, m))
    t_sp = []
    for k in range(3):
        t0 = time.perf_counter()
        cm.sample_predict(f, cfg, seeds=np.arange(chains) + 1000, out=pb.array)
        t_sp.append(time.perf_counter() - t0)
    # cpu baseline: bounded sample on one core
    sample = hx[: min(count, 2_000_000)]
    of = OracleFunction(frir)
    t0 = time.perf_counter()
    of(sample)
    t_cpu = time.perf_counter() - t0
    print(json.dumps({
        "metric": "posterior_draws_evaluated_per_sec", "value": count / (ms * 1e-3), "unit": "draws/s", "ms_per_step": ms,
        "steps": args.steps, "warmup": args.warmup, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "eight_schools_predict_12_requirements", "chains": chains, "iterations": iters, "n": n, "m": m,
                   "math": "fast" if args.fast else "parity", "l2": "L2 flushed (256 MB write) between timed steps",
                   "layout": "[iteration][n][chain] -> [chain][iteration][m]"},
        "gpu_launches": args.steps, "bit_identical_to_oracle_first_64_chains": same,
        "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None, "algorithmic_bytes_per_launch": alg,
                     "bytes_per_draw": (n + m) * 8.0, "peak_source": src},
        "e2e": {"value": count / min(t_e2e), "unit": "draws/s", "h2d_bytes_per_step": count * n * 8, "d2h_bytes_per_step": count * m * 8,
                "api": "rn_function_eval (C ABI), pageable host buffers",
                "sample_predict_ms_per_call": min(t_sp) * 1e3,
                "sample_predict_note": "rn_sample_predict: 50 warmup + %d sampling iterations of HMC(5) for %d chains, predictions "
                                       "[chains][iterations][m] into a page-locked buffer (%d MB instead of %d MB of draws)"
                                       % (iters, chains, count * m * 8 >> 20, count * n * 8 >> 20)},
        "cpu_baseline": {"value": len(sample) / t_cpu, "unit": "draws/s", "cores": 1, "kind": "port",
                         "sample": "%d draws through rno_function_eval" % len(sample)},
    }))


if __name__ == "__main__":
    main()
