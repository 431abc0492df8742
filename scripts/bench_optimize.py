#!/usr/bin/env python
"""
scripts/bench_optimize.py -- measurement of SURVEY.md 8f-4 (rn_optimize: batched multi-start L-BFGS) on a GPU box.

Workload: eight schools (n = 10, the reference's EightSchools benchmark model), `starts` starts drawn N(0, 0.7^2) around
the reference's start x = 0, m = 5, eps = 0.1 (Optimizer.scala:12-13).  A "step" = one rn_optimize call (one launch of
rn_k_lbfgs).  Metric: density+gradient evaluations per second summed over starts (the unit Optimizer.lbfgs spends its
time in); `starts_per_sec` beside it.  The kernel is FP64-latency bound like rn_k_iter, not HBM bound (the whole state
is thread-local); the roofline object therefore reports emitter-counted fp64 flops against the fp64 pipe.
cpu_baseline: the Python/C++ oracle on one core, bounded sample (the reference's optimizer is single-threaded too).
Prints one JSON line.  Not part of the product.
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
    ap.add_argument("--starts", type=int, default=151552)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--fast", action="store_true")
    args = ap.parse_args()
    from rainier_b200 import api

    rir, cols = open(os.path.join(ROOT, "rainier_b200", "models", "eight_schools.rir"), "rb").read(), []  # committed fixture
    cm = api.CudaModel(rir, cols)
    x0 = np.random.default_rng(0).normal(size=(args.starts, 10)) * 0.7
    x0[0] = 0.0
    times, res = [], None
    for k in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = cm.optimize(x0, fast=args.fast, max_evals=400)
        if k >= args.warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.median(times))
    evals = int(res["evals"].sum())
    from oracle.rainier_py.binding import OracleModel  # checker / CPU baseline only
    from oracle.rainier_py.optimizer import lbfgs
    om = OracleModel(rir, cols)
    t0 = time.perf_counter()
    ref = [lbfgs(om.density_batch, 10, x0=x, max_evals=400) for x in x0[:64]]
    t_cpu = time.perf_counter() - t0
    same = all(np.array_equal(res["x"][c], np.array(r["x"]), equal_nan=True) and res["evals"][c] == r["evals"] for c, r in enumerate(ref))
    oc = cm.op_counts()
    flops_eval = oc["flops_invariant"] + oc["flops_rows"]
    print(json.dumps({
        "metric": "density_gradient_evaluations_per_sec", "value": evals / t, "unit": "evaluations/s", "starts_per_sec": args.starts / t,
        "ms_per_step": t * 1e3, "steps": args.steps, "warmup": args.warmup, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "eight_schools_lbfgs_multi_start", "starts": args.starts, "m": 5, "eps": 0.1,
                   "math": "fast" if args.fast else "parity", "timing": "host wall clock around rn_optimize (H2D of x0 and D2H of x inside)"},
        "gpu_launches": args.steps, "converged_frac": float((res["info"] == 0).mean()), "mean_evals_per_start": evals / args.starts,
        "bit_identical_to_oracle_first_64_starts": bool(same) if not args.fast else None,
        "roofline": {"bound": "fp64", "achieved": evals / t * flops_eval / 1e12, "unit": "TFLOP/s (emitter-counted density flops only)",
                     "peak": None, "frac": None, "traffic": None},
        "cpu_baseline": {"value": sum(r["evals"] for r in ref) / t_cpu, "unit": "evaluations/s", "cores": 1, "kind": "port",
                         "sample": "64 starts through oracle/rainier_py/optimizer.py over rno_density_batch"},
    }))


if __name__ == "__main__":
    main()
