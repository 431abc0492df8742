"""Secondary measurements: the other BASELINE.json configurations (cfg 2..5) on one B200, device-resident, HMC nSteps=5
with a static step size in the stable regime, next to the CPU oracle on a bounded sample.  Writes JSON lines.
Usage: python scripts/bench_configs.py [cfg2 cfg2s cfg3 cfg4 cfg5] [--small]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

try:
    import torch
except Exception:  # precompile mode on a box without torch/cuda is fine
    torch = None

from oracle.rainier_py import configs
from oracle.rainier_py.binding import OracleModel, lib as olib
from rainier_b200 import abi, api

small = "--small" in sys.argv
precompile = "--precompile" in sys.argv  # CPU box: emit + NVRTC into $RN_KERNEL_CACHE, no device needed
only_math = [a.split("=")[1] for a in sys.argv if a.startswith("--math=")]
no_cpu = "--no-cpu" in sys.argv
which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg2", "cfg2s", "cfg3", "cfg4", "cfg5"]


def cached(name, build):
    """RIR bytes + columns of a config, cached under build/models/ (the Python DAG restatement needs minutes for the
    1M-row model; the cache travels to the GPU box with the snapshot)."""
    d = os.path.join(ROOT, "build", "models")
    os.makedirs(d, exist_ok=True)
    f = os.path.join(d, name + ("_small" if small else "") + ".npz")
    if os.path.exists(f):
        z = np.load(f)
        return z["rir"].tobytes(), [z["c%d" % i] for i in range(int(z["ncols"]))]
    rir, cols = build()
    np.savez(f, rir=np.frombuffer(rir, dtype=np.uint8), ncols=len(cols), **{"c%d" % i: np.asarray(c, dtype=np.float64) for i, c in enumerate(cols)})
    return rir, cols


def timed_run(model, cfg, chains, iters, reps=3):
    s = api.CudaSampler(model, cfg, seeds=np.arange(chains) + 1000)
    s.warmup(-1)
    stream = torch.cuda.ExternalStream(s.stream)
    n = model.nVars
    d = torch.empty((iters, n, chains), dtype=torch.float64, device="cuda")
    s.run(iters, d.data_ptr())
    s.sync()
    st0 = sum(x.leapfrogSteps for x in s.stats()[0])
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        s.run(iters, d.data_ptr())
        e1.record(stream)
        s.sync()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    stats = s.stats()[0]
    steps = (sum(x.leapfrogSteps for x in stats) - st0) / reps  # stats accumulate over the sampling phase
    acc = float(np.mean([x.accepted / max(1, x.iterations) for x in stats]))
    s.close()
    return steps / best, best, acc


def cpu_rate(rir, cols, cfg, iters):
    om = OracleModel(rir, cols)
    cores = olib().rno_hardware_threads()
    c = api.lower_config(cfg)[0]
    c.iterations = iters
    t = time.perf_counter()
    r = om.sample(c, seeds=np.arange(cores) + 1000)
    dt = time.perf_counter() - t
    steps = sum(x.leapfrog_steps for x in r["stats"])
    return steps / dt, cores, dt


def static(eps, iters, **kw):
    return api.make_config(iterations=iters, warmupIterations=0, sampler=api.HMCSampler(5), stepSizeTuner=api.StaticStepSize(eps),
                           massMatrixTuner=api.IdentityMassMatrixTuner(), launchIterations=iters, **kw)


out = []
for name in which:
    t0 = time.perf_counter()
    if name == "cfg2":  # README linear regression, 3 covariates: inlined by the reference -> data-free
        n_obs = 2000 if small else 10000
        rir, cols = cached("cfg2", lambda: configs.linreg(n_obs).compile(True))
        chains, iters, eps, label = 4096, 200, 0.002, "linreg 3 cov, %d obs (inlined, data-free), 4096 chains" % n_obs
        prir, pcols = rir, cols
    elif name == "cfg2s":  # 5 covariates: the reference streams the data
        n_obs = 2000 if small else 10000
        rir, cols = cached("cfg2s", lambda: configs.linreg(n_obs, covariates=5).compile(True))
        chains, iters, eps, label = 4096, 20, 0.002, "linreg 5 cov, %d obs (streamed), 4096 chains" % n_obs
        prir, pcols = rir, cols
    elif name == "cfg2si":  # the same 5-covariate regression sent as the streamed PRIMAL container: inlined on the device at create
        from oracle.make_bench_models import streamed_primal
        n_obs = 2000 if small else 10000
        prir, pcols = cached("cfg2s_primal", lambda: streamed_primal(configs.linreg(n_obs, covariates=5)))
        rir, cols = None, None
        chains, iters, eps, label = 4096, 200, 0.002, "linreg 5 cov, %d obs, streamed container inlined on the device, 4096 chains" % n_obs
    elif name == "cfg3":
        n_obs, d = (10000, 50) if small else (100000, 50)
        # GPU: primal RIR + the emitter's adjoint gradient (what CudaCompiler sends; 43 MB of columns, L2-resident).
        # CPU oracle: the reference's symbolic-gradient form (its gradient columns quadruple the data to 167 MB).
        prir, pcols = cached("cfg3_primal", lambda: configs.logreg(n_obs, d).compile(False))
        if no_cpu or precompile or not (os.path.exists(os.path.join(ROOT, "build", "models", "cfg3.npz")) or os.path.isdir("/root/reference")):
            rir, cols = None, None  # (the 167 MB symbolic form is not shipped to the GPU box)
        else:
            rir, cols = cached("cfg3", lambda: configs.logreg(n_obs, d).compile(True))
        chains, iters, eps, label = 2048, 2, 0.01, "logreg %d cov, %d obs, 2048 chains (primal RIR, adjoint gradient)" % (d, n_obs)
    elif name == "cfg4":
        rir, cols = cached("cfg4", lambda: configs.eight_schools().compile(True))
        chains, iters, eps, label = 8192, 200, 0.1, "eight schools, 8192 chains"
        prir, pcols = rir, cols
    elif name == "cfg5":
        g, n_obs = (100, 100000) if small else (1000, 1000000)
        prir, pcols = cached("cfg5_primal", lambda: configs.poisson_glm(g, n_obs).compile(False))  # primal only: the symbolic gradient is infeasible
        rir, cols = None, None
        chains, iters, eps, label = 4096, 1, 0.001, "poisson GLM %d groups, %d obs, 4096 chains (primal RIR, adjoint gradient)" % (g, n_obs)
    build_s = time.perf_counter() - t0
    res = {"config": name, "label": label, "model_build_s": round(build_s, 1)}
    mk = static
    if name == "cfg5" and "--static-step" not in sys.argv:
        # a fixed step from a random start is rejected every time on a posterior this narrow (accept_rate 0.0 in round 1's
        # numbers): like bench.py, take the step from 30 warmup iterations of DualAvg(0.8) and time the sampling phase
        def mk(eps, iters, **kw):
            return api.make_config(iterations=iters, warmupIterations=30, sampler=api.HMCSampler(5), stepSizeTuner=api.DualAvgTuner(0.8),
                                   massMatrixTuner=api.IdentityMassMatrixTuner(), launchIterations=iters, **kw)
        res["step_size"] = "DualAvg(0.8), 30 warmup iterations"
    if precompile:
        model = api.CudaModel(prir, pcols, device=-1)
        for mm in (abi.RN_MATH_PARITY, abi.RN_MATH_FAST):
            t1 = time.perf_counter()
            model.emit_cubin(mk(eps, iters, mathMode=mm))
            print(name, "math", mm, "compiled in %.1f s" % (time.perf_counter() - t1), flush=True)
        continue
    model = api.CudaModel(prir, pcols)
    for math_mode, mm in (("parity", abi.RN_MATH_PARITY), ("fast", abi.RN_MATH_FAST)):
        if only_math and math_mode not in only_math:
            continue
        cfg = mk(eps, iters, mathMode=mm)
        try:
            rate, secs, acc = timed_run(model, cfg, chains, iters)
            res[math_mode] = {"steps_x_chains_per_s": rate, "seconds_per_launch": secs, "accept_rate": acc}
        except api.RainierCudaError as e:
            res[math_mode] = {"error": str(e)[:300]}
    res["backend"] = "warp-per-chain" if "#define RN_BACKEND 1" in model.emit_source(mk(eps, iters)) else "thread-per-chain"
    res["op_counts"] = model.op_counts(mk(eps, iters))
    if rir is not None and not no_cpu:
        cpu_iters = {"cfg2": 2000, "cfg2s": 20, "cfg3": 1, "cfg4": 2000}[name]
        r, cores, dt = cpu_rate(rir, cols, static(eps, cpu_iters), cpu_iters)
        res["cpu_oracle"] = {"steps_x_chains_per_s": r, "cores": cores, "seconds": round(dt, 2)}
    print(json.dumps(res), flush=True)
