"""
Debug aid for the CPU box (no GPU here): compiles the CUDA source the emitter produced with g++ under
-DRN_HOST_EMULATION (rn_prelude.cuh maps the few CUDA constructs used to plain C++) and runs the emitted
rn_density() on the host, one "thread" at a time.  This checks the *emitter* (lowering, reverse-mode adjoints,
accumulation order) against the oracle without a device.  It is test infrastructure only -- the product never
compiles or runs this; the product path fails loudly without CUDA.
"""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

_SAMPLER_SHIM = r"""
#include <vector>
#include <cstring>
// Host emulation of rn_sampler_create / warmup / run for the THREAD-PER-CHAIN kernels (one launch per phase, one
// "thread" at a time).  Test infrastructure only: lets the CPU test-suite run the hand-written sampler source
// (rn_sampler.cuh) against the oracle; the product never executes this.
struct EmuCfg {
  int sampler, n_steps, max_steps, min_steps, buf_size, step_tuner;
  double p_count, delta, static_step;
  int mass_tuner, initial_window, skip_first, skip_last;
  double win_expansion;
  int stats_window, warmup, iterations;
  int static_kind;              // StaticMassMatrix: 0 identity, 1 diagonal, 2 dense (mass_tuner == 3)
  const double* static_elements;  // n or n*n
  int chains_per_cta;           // warp-per-chain source: chains of one emulated CTA (1 = one chain at a time)
};
extern "C" int emu_sample(const EmuCfg* c, const long long* seeds, int chains, const double* data, double* samples,
                          double* trace, long long* out_stats, double* mass_out, int* mass_kind_out) {
  const size_t C = (size_t)chains, n = RN_N, W = (size_t)c->stats_window;
  std::vector<double> params((2 * n + 1) * C), grad(n * C), nng(C), da(5 * C), mass(n * n * C + 1), chol(n * (n + 1) / 2 * C + 1),
      emean(n * C + 1), eraw(n * C + 1), ecov(n * n * C + 1), ring((size_t)c->buf_size * C + 1), energy(3 * C), rings(3 * W * C);
  std::vector<long long> seed(C), grads(C), steps(C);
  std::vector<int> have(C), dait(C), ring_i(C), ring_full(C), err(C), iters(C), accepted(C), energy_n(C), sri(3 * C), srf(3 * C);
  RnArgs a;
  std::memset(&a, 0, sizeof(a));
  a.chains = chains;
  a.params = params.data(); a.grad = grad.data(); a.rng_seed = seed.data(); a.rng_nng = nng.data(); a.rng_have = have.data();
  a.da = da.data(); a.da_iter = dait.data(); a.mass = mass.data(); a.chol = chol.data(); a.est_mean = emean.data();
  a.est_raw = eraw.data(); a.est_cov = ecov.data(); a.ring = ring.data(); a.ring_i = ring_i.data(); a.ring_full = ring_full.data();
  a.st_err = err.data(); a.st_grads = grads.data(); a.st_steps = steps.data(); a.st_iters = iters.data();
  a.st_accepted = accepted.data(); a.st_energy = energy.data(); a.st_energy_n = energy_n.data(); a.st_rings = rings.data();
  a.st_ring_i = sri.data(); a.st_ring_full = srf.data(); a.data = data;
  a.sampler = c->sampler; a.n_steps = c->n_steps; a.max_steps = c->max_steps; a.min_steps = c->min_steps; a.buf_size = c->buf_size;
  a.step_tuner = c->step_tuner; a.p_count = c->p_count; a.delta = c->delta; a.static_step = c->static_step;
  a.mass_tuner = c->mass_tuner; a.total_warmup = c->warmup; a.skip_first = c->skip_first; a.skip_last = c->skip_last;
  a.win_expansion = c->win_expansion; a.stats_window = c->stats_window;
  a.chain_begin = 0; a.chain_end = chains;
#if defined(RN_TMA_STAGES) && RN_TMA_STAGES > 0
  a.tma = 1;  // rn_runtime.cpp: run_phase
#endif
  for (size_t k = 0; k < C; k++) seed[k] = (seeds[k] ^ 0x5DEECE66DLL) & ((1LL << 48) - 1);  // new java.util.Random(seed)
  auto launch = [&](void (*kern)(const RnArgs)) {
@LAUNCH@  };
  a.mass_kind = 0;
  launch(rn_k_init);  // LeapFrog.initialize draws with the identity matrix (Driver.scala:22) also under a StaticMassMatrix
  int win_size = c->initial_window, win_i = 0, win_j = 0, est = 0, mass_kind = 0;
  if (c->mass_tuner == 3 && c->static_kind != 0) {  // what rn_sampler_create uploads: the matrix replicated per chain and,
    mass_kind = c->static_kind;                     // for a dense one, choleskyUpperTriangular (MassMatrix.scala:76-117)
    const size_t ne = c->static_kind == 2 ? n * n : n;
    for (size_t e = 0; e < ne; e++)
      for (size_t k = 0; k < C; k++) mass[e * C + k] = c->static_elements[e];
    if (c->static_kind == 2) {
      auto tri = [](size_t k) { return (k * (k + 1)) / 2; };
      std::vector<double> lower(tri(n), 0.0);
      size_t l = 0;
      for (size_t i = 0; i < n; i++)
        for (size_t k = 0; k <= i; k++) {
          double sum = 0.0;
          for (size_t j = 0; j < k; j++) sum += lower[tri(i) + j] * lower[tri(k) + j];
          const double x = c->static_elements[i * n + k] - sum;
          lower[l++] = (i == k) ? std::sqrt(x) : (1.0 / lower[tri(k + 1) - 1] * x);
        }
      l = 0;
      for (size_t i = 0; i < n; i++)
        for (size_t k = 0; k < n - i; k++, l++)
          for (size_t ch = 0; ch < C; ch++) chol[l * C + ch] = lower[tri(k + i) + i];
    }
  }
  if (c->warmup > 0) {
    a.phase = 0; a.n_iter = c->warmup; a.mass_kind = mass_kind; a.win_size = win_size; a.win_i = 0; a.win_j = 0; a.est_samples = 0;
    a.trace = trace;
    launch(RN_K_WARMUP);
    if (c->mass_tuner == 1 || c->mass_tuner == 2)  // host mirror of WindowedMassMatrixTuner.update (rn_runtime.cpp: advance_window)
      for (int k = 0; k < c->warmup; k++) {
        win_j += 1;
        if (win_j < c->skip_first || (c->warmup - win_j) < c->skip_last) continue;
        win_i += 1; est += 1;
        if (win_i == win_size) { win_i = 0; win_size = (int)(win_size * c->win_expansion); mass_kind = c->mass_tuner; }
      }
  }
  // lf.resetStats(), Driver.scala:31
  std::fill(grads.begin(), grads.end(), 0); std::fill(steps.begin(), steps.end(), 0); std::fill(iters.begin(), iters.end(), 0);
  std::fill(accepted.begin(), accepted.end(), 0); std::fill(energy.begin(), energy.end(), 0.0); std::fill(energy_n.begin(), energy_n.end(), 0);
  std::fill(rings.begin(), rings.end(), 0.0); std::fill(sri.begin(), sri.end(), 0); std::fill(srf.begin(), srf.end(), 0);
  if (c->iterations > 0) {
    a.phase = 1; a.n_iter = c->iterations; a.mass_kind = mass_kind; a.samples = samples;
    a.trace = trace ? trace + (size_t)c->warmup * 4 * C : nullptr;
    launch(rn_k_iter);
  }
  for (size_t k = 0; k < C; k++) {
    out_stats[k * 5 + 0] = grads[k]; out_stats[k * 5 + 1] = steps[k]; out_stats[k * 5 + 2] = accepted[k]; out_stats[k * 5 + 3] = seed[k];
    out_stats[k * 5 + 4] = err[k];
  }
  const size_t ne = mass_kind == 2 ? n * n : n;
  for (size_t e = 0; e < ne * C; e++) mass_out[e] = mass_kind == 0 ? 1.0 : mass[e];
  *mass_kind_out = mass_kind;
  return 0;
}
"""

_WPC_RUNNER = r"""
#include <thread>
#include <vector>
#include <functional>
// one emulated chain = RN_G = 32*K host threads: a barrier per warp and one for the group (rn_prelude.cuh,
// RN_HOST_EMULATION && RN_BACKEND == 1).  A CTA holds `slots` chains (`active` of them own a chain; the others leave at
// once, like the idle warps of the last CTA of a launch); with slots > 1 the CTA-level barriers are real.
static void rn_emu_run_cta(int block, int nblocks, int slots, int active, const std::function<void()>& body) {
  std::vector<RnEmuGroup> groups(slots);
  for (auto& g : groups) {
    pthread_barrier_init(&g.bar, nullptr, RN_G);
    for (int k = 0; k < RN_WPC_K; k++) pthread_barrier_init(&g.warp[k].bar, nullptr, 32);
  }
  RnEmuCta cta;
  pthread_barrier_init(&cta.all, nullptr, slots * RN_G);
  pthread_barrier_init(&cta.active, nullptr, active * RN_G);
  rn_emu_cta = slots > 1 ? &cta : nullptr;
  std::vector<std::thread> th;
  for (int tid = 0; tid < slots * RN_G; tid++)
    th.emplace_back([&, tid] {
      rn_emu_group = &groups[tid / RN_G];
      blockDim.x = (unsigned)(slots * RN_G); gridDim.x = (unsigned)nblocks; blockIdx.x = (unsigned)block; threadIdx.x = (unsigned)tid;
      body();
    });
  for (auto& t : th) t.join();
  rn_emu_cta = nullptr;
  pthread_barrier_destroy(&cta.all);
  pthread_barrier_destroy(&cta.active);
  for (auto& g : groups) {
    pthread_barrier_destroy(&g.bar);
    for (int k = 0; k < RN_WPC_K; k++) pthread_barrier_destroy(&g.warp[k].bar);
  }
}
static void rn_emu_run_warp(int block, int nblocks, const std::function<void()>& body) { rn_emu_run_cta(block, nblocks, 1, 1, body); }
"""

_WPC_SHIM = _WPC_RUNNER + r"""
extern "C" void emu_density(const double* q, int chains, double* out, const double* data, int* err) {
  for (int c = 0; c < chains; c++) rn_emu_run_warp(c, chains, [&] { rn_k_density(q, out, data, err, chains); });
}
"""

_SHIM = r"""
extern "C" void emu_density(const double* q, int chains, double* out, const double* data, int* err) {
  blockDim.x = 1; gridDim.x = (unsigned)chains; threadIdx.x = 0;
  for (int c = 0; c < chains; c++) { blockIdx.x = (unsigned)c; rn_k_density(q, out, data, err, chains); }
}
"""


_FUNCTION_SHIM = r"""
// Host emulation of one rn_k_eval launch (rn_function.cuh): `grid` CTAs of 128 "threads", one at a time.
extern "C" void emu_eval(const double* x, double* out, long long count, const long long* lay, int* err, int grid) {
  RnEvalArgs a;
  a.x = x; a.out = out; a.count = count;
  a.in_inner = lay[0]; a.in_outer = lay[1]; a.in_pstride = lay[2]; a.in_estride = lay[3];
  a.out_inner = lay[4]; a.out_outer = lay[5]; a.out_pstride = lay[6]; a.out_estride = lay[7];
  a.err = err;
  blockDim.x = 128; gridDim.x = (unsigned)grid;
  for (int b = 0; b < grid; b++)
    for (int t = 0; t < 128; t++) { blockIdx.x = (unsigned)b; threadIdx.x = (unsigned)t; rn_k_eval(a); }
}
"""


_OPTIMIZER_SHIM = r"""
// Host emulation of one rn_k_lbfgs launch (rn_optimizer.cuh): one "thread" (start) at a time.
extern "C" void emu_lbfgs(const double* x0, double* x, double* f, int* info, int* evals, const double* data, double eps,
                          int starts, int max_evals) {
  RnOptArgs a;
  a.x0 = x0; a.x = x; a.f = f; a.info = info; a.evals = evals; a.data = data; a.eps = eps; a.starts = starts; a.max_evals = max_evals;
  blockDim.x = 1; gridDim.x = (unsigned)starts; threadIdx.x = 0;
  for (int c = 0; c < starts; c++) { blockIdx.x = (unsigned)c; rn_k_lbfgs(a); }
}
"""


_OPTIMIZER_SHIM_WPC = r"""
// one emulated start = one warp of 32 host threads (see _WPC_RUNNER)
extern "C" void emu_lbfgs(const double* x0, double* x, double* f, int* info, int* evals, const double* data, double eps,
                          int starts, int max_evals) {
  RnOptArgs a;
  a.x0 = x0; a.x = x; a.f = f; a.info = info; a.evals = evals; a.data = data; a.eps = eps; a.starts = starts; a.max_evals = max_evals;
  for (int c = 0; c < starts; c++) rn_emu_run_warp(c, starts, [&] { rn_k_lbfgs(a); });
}
"""


def compile_source(src, fast=False, opt="-O1"):
    d = os.path.join(tempfile.gettempdir(), "rn_emul")
    os.makedirs(d, exist_ok=True)
    key = hashlib.sha1((src + str(fast) + opt + os.environ.get("RN_EMU_ASAN", "")).encode()).hexdigest()[:16]
    so = os.path.join(d, key + ".so")
    if not os.path.exists(so):
        cpp = os.path.join(d, key + ".cpp")
        with open(cpp, "w") as f:
            wpc = "#define RN_BACKEND 1" in src
            launch_tpc = ("    blockDim.x = 1; gridDim.x = (unsigned)chains; threadIdx.x = 0;\n"
                          "    for (int k = 0; k < chains; k++) { blockIdx.x = (unsigned)k; kern(a); }\n")
            launch_wpc = ("    const int spc = c->chains_per_cta > 0 ? c->chains_per_cta : 1, nb = (chains + spc - 1) / spc;\n"
                          "    for (int b = 0; b < nb; b++)\n"
                          "      rn_emu_run_cta(b, nb, spc, (chains - b * spc) < spc ? (chains - b * spc) : spc, [&] { kern(a); });\n")
            if src.startswith("// generated by rainier_b200 (CUDA source emitter, function flavour)"):  # rn_function.cuh
                f.write(src + _FUNCTION_SHIM)
            elif src.startswith("// generated by rainier_b200 (CUDA source emitter, optimizer flavour)"):  # rn_optimizer.cuh
                f.write(src + (_WPC_RUNNER + _OPTIMIZER_SHIM_WPC if wpc else _OPTIMIZER_SHIM))
            else:
                f.write(src + (_WPC_SHIM if wpc else _SHIM) + _SAMPLER_SHIM.replace("@LAUNCH@", launch_wpc if wpc else launch_tpc))
        flags = [opt, "-std=c++17", "-fPIC", "-shared", "-DRN_HOST_EMULATION", "-w", "-pthread"] + (["-g", "-fsanitize=address"] if os.environ.get("RN_EMU_ASAN") else [])
        flags.append("-ffp-contract=fast" if fast else "-ffp-contract=off")
        subprocess.run(["g++"] + flags + [cpp, "-o", so], check=True)
    return C.CDLL(so)


def density(src, q, cols, model, fast=False, opt="-O1"):
    """q: [chains][n] -> [chains][n+1] using the emitted code.  `model`: the CudaModel the source came from (its
    rn_model_pack_columns lays the columns out exactly as rn_model_create uploads them: tile-major per target)."""
    L = compile_source(src, fast, opt)
    q = np.ascontiguousarray(q, dtype=np.float64)
    chains, n = q.shape
    qt = np.ascontiguousarray(q.T)
    out = np.zeros((n + 1, chains))
    data = model.pack_columns()
    err = C.c_int(0)
    L.emu_density(C.c_void_p(qt.ctypes.data), chains, C.c_void_p(out.ctypes.data), C.c_void_p(data.ctypes.data), C.byref(err))
    return np.ascontiguousarray(out.T), err.value


class EmuCfg(C.Structure):
    _fields_ = [("sampler", C.c_int), ("n_steps", C.c_int), ("max_steps", C.c_int), ("min_steps", C.c_int), ("buf_size", C.c_int),
                ("step_tuner", C.c_int), ("p_count", C.c_double), ("delta", C.c_double), ("static_step", C.c_double),
                ("mass_tuner", C.c_int), ("initial_window", C.c_int), ("skip_first", C.c_int), ("skip_last", C.c_int),
                ("win_expansion", C.c_double), ("stats_window", C.c_int), ("warmup", C.c_int), ("iterations", C.c_int),
                ("static_kind", C.c_int), ("static_elements", C.POINTER(C.c_double)), ("chains_per_cta", C.c_int)]


def sample(src, cfg, seeds, model, chains_per_cta=1):
    """Runs the emitted thread-per-chain sampler kernels on the host.  cfg: lowered rn_config (abi.Config).
    Returns dict(samples [chains][iters][n], trace [chains][warm+iters][4], stats [chains][5], mass, mass_kind)."""
    L = compile_source(src)
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    chains, n = len(seeds), model.nVars
    e = EmuCfg(cfg.sampler, cfg.n_steps, cfg.max_steps, cfg.min_steps, cfg.buf_size, cfg.step_size_tuner, cfg.p_count, cfg.delta,
               cfg.static_step_size, cfg.mass_tuner, cfg.initial_window_size, cfg.skip_first, cfg.skip_last, cfg.window_expansion,
               cfg.stats_window, cfg.warmup_iterations, cfg.iterations,
               cfg.static_matrix if cfg.mass_tuner == 3 else 0, cfg.static_matrix_elements if cfg.mass_tuner == 3 else None,
               int(chains_per_cta))
    it, tot = cfg.iterations, cfg.warmup_iterations + cfg.iterations
    samples = np.zeros((max(it, 1), n, chains))
    trace = np.zeros((max(tot, 1), 4, chains))
    stats = np.zeros((chains, 5), dtype=np.int64)
    mass = np.zeros((n * n, chains))
    kind = C.c_int(0)
    data = model.pack_columns()
    L.emu_sample.argtypes = [C.POINTER(EmuCfg), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.POINTER(C.c_int)]
    L.emu_sample(C.byref(e), seeds.ctypes.data, chains, data.ctypes.data, samples.ctypes.data, trace.ctypes.data, stats.ctypes.data,
                 mass.ctypes.data, C.byref(kind))
    ne = n * n if kind.value == 2 else n
    return {"samples": np.ascontiguousarray(samples[:it].transpose(2, 0, 1)), "trace": np.ascontiguousarray(trace[:tot].transpose(2, 0, 1)),
            "stats": stats, "mass": np.ascontiguousarray(mass.reshape(-1)[: ne * chains].reshape(ne, chains).T), "mass_kind": kind.value}


def eval_function(src, x, m, layout="rows", iterations=None, chains=None, grid=3, fast=False):
    """Runs the emitted function source (CudaFunction.emit_source) on the host through rn_k_eval's own addressing.
    layout "rows": x [count][n] -> [count][m].  layout "sampler": x [iterations][n][chains] (as rn_sampler_run writes
    draws) -> [chains][iterations][m] (Trace.predict's order), the strides rn_function_eval_device passes."""
    L = compile_source(src, fast)
    x = np.ascontiguousarray(x, dtype=np.float64)
    if layout == "rows":
        count, n = x.shape
        lay = [count, 0, n, 1, count, 0, m, 1]
        out = np.full((count, m), np.nan)
    else:
        n = x.shape[1]
        count = iterations * chains
        lay = [chains, n * chains, 1, chains, chains, m, iterations * m, 1]
        out = np.full((chains, iterations, m), np.nan)
    lay_a = (C.c_longlong * 8)(*[max(int(v), 0) if k not in (0, 4) else max(int(v), 1) for k, v in enumerate(lay)])
    err = C.c_int(0)
    L.emu_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.POINTER(C.c_int), C.c_int]
    L.emu_eval(x.ctypes.data, out.ctypes.data, count, lay_a, C.byref(err), grid)
    return out, err.value


def optimize(src, model, x0=None, starts=1, eps=0.1, max_evals=10000, fast=False):
    """Runs the emitted optimizer source (CudaModel.emit_optimizer_source) on the host.  Returns dict like
    CudaModel.optimize."""
    L = compile_source(src, fast)
    n = model.nVars
    if x0 is not None:
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, n)
        starts = x0.shape[0]
        x0t = np.ascontiguousarray(x0.T)
    x = np.zeros((n, starts))
    f = np.zeros(starts)
    info = np.zeros(starts, dtype=np.int32)
    evals = np.zeros(starts, dtype=np.int32)
    data = model.pack_columns()
    L.emu_lbfgs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int]
    L.emu_lbfgs(x0t.ctypes.data if x0 is not None else None, x.ctypes.data, f.ctypes.data, info.ctypes.data, evals.ctypes.data,
                data.ctypes.data, eps, starts, max_evals)
    return {"x": np.ascontiguousarray(x.T), "f": f, "info": info, "evals": evals}
