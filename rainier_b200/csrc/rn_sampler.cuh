// rn_sampler.cuh -- hand-written batched HMC/EHMC integrator, appended after the emitted rn_density().
//
// One CUDA thread owns one chain (RN_BACKEND == 0, "thread per chain") and runs whole iterations --
// momentum draw, leapfrog steps, Metropolis test, adaptation, sample write -- inside one launch; chain state
// lives in registers across the L steps of an iteration and in chain-fastest (coalesced) SoA arrays between
// iterations.  Control flow follows the reference line by line so that chain c reproduces a single-chain
// reference run seeded with ScalaRNG(seeds[c]):
//   LeapFrog      rainier-sampler/src/main/scala/com/stripe/rainier/sampler/LeapFrog.scala:3-252
//   HMCSampler    .../sampler/HMC.scala:3-24         EHMCSampler  .../sampler/EHMC.scala:3-62
//   DualAvgTuner  .../sampler/DualAvg.scala:3-90     mass tuners  .../sampler/MassMatrix.scala:120-181
//   estimators    .../sampler/MassMatrixEstimator.scala:9-112
//   Driver        .../sampler/Driver.scala:7-119     Stats/RingBuffer .../sampler/Stats.scala:3-59
// The reference re-evaluates the density in every fullPs() at the position the previous update already
// evaluated (LeapFrog.scala:168-176 vs :161-166); update() is a pure function of q, so this kernel keeps the
// gradient of the last evaluation instead (l+1 evaluations per takeSteps(l) instead of 2l+1) while still
// counting gradientEvaluations the reference's way.
//
// Compile-time switches (set by the emitter): RN_N, RN_NSLOTS, RN_MASS_MAX (0 identity only, 1 +diagonal,
// 2 +dense), RN_ENABLE_EHMC.
#ifndef RN_SAMPLER_CUH
#define RN_SAMPLER_CUH

// struct RnArgs: see rn_args.h (shared verbatim with the host runtime)

#define RN_LN2 0.6931471805599453
#define RN_AT(ptr, field, c) (ptr)[(size_t)(field) * (size_t)A.chains + (size_t)(c)]

struct RnStats {
  rn_i64 grads, steps;
  int iters, accepted, err;
  double e_mean, e_raw, trans2;
  int e_n;
  int ring_i[3], ring_full[3];
};

RN_DEVICE void rn_ring_add(const RnArgs& A, int c, RnStats& S, int which, double value) {  // Stats.scala:24-30
  int i = S.ring_i[which] + 1;
  if (i == A.stats_window) S.ring_full[which] = 1;
  i = i % A.stats_window;
  S.ring_i[which] = i;
  RN_AT(A.st_rings, which * A.stats_window + i, c) = value;
}

struct RnMass {
  int kind;
#if RN_MASS_MAX >= 1
  double m[RN_N];  // diagonal elements (variances)
#endif
};

// velocity = M^-1 p  (LeapFrog.scala:205-219)
RN_DEVICE void rn_velocity(const RnArgs& A, int c, const RnMass& M, const double (&in)[RN_N], double (&out)[RN_N]) {
  (void)A;
  (void)c;
#if RN_MASS_MAX >= 2
  if (M.kind == 2) {  // DenseMassMatrix.squareMultiply, MassMatrix.scala:35-51
    for (int i = 0; i < RN_N; i++) {
      double y = 0.0;
      for (int j = 0; j < RN_N; j++) y += in[j] * RN_AT(A.mass, i * RN_N + j, c);
      out[i] = y;
    }
    return;
  }
#endif
#if RN_MASS_MAX >= 1
  if (M.kind == 1) {
    RN_UNROLL
    for (int i = 0; i < RN_N; i++) out[i] = in[i] * M.m[i];
    return;
  }
#endif
  RN_UNROLL
  for (int i = 0; i < RN_N; i++) out[i] = in[i];
}

// energy = potential + dot(velocity, p)/2  (LeapFrog.scala:134-139,221-231)
RN_DEVICE double rn_energy(const RnArgs& A, int c, const RnMass& M, const double (&p)[RN_N], double U) {
  double v[RN_N];
  rn_velocity(A, c, M, p, v);
  double k = 0.0;
  RN_UNROLL
  for (int i = 0; i < RN_N; i++) k += (v[i] * p[i]);
  return U + k / 2.0;
}

RN_DEVICE double rn_log_accept(double deltaH) {  // LeapFrog.scala:141-145
  if (deltaH != deltaH) return -RN_INF;
  return rn_jmin0(-deltaH);
}

struct RnPQ {  // pqBuf + the gradient at pqBuf.q
  double p[RN_N], q[RN_N], g[RN_N];
  double U;
};

RN_DEVICE void rn_update(const RnArgs& A, RnPQ& s, RnStats& S) {  // copyQsAndUpdateDensity + potential
  double dens;
  rn_density(s.q, dens, s.g, A.data, S.err);
  s.U = dens * -1;
  S.grads += 1;
}
RN_DEVICE void rn_full_ps(RnPQ& s, double stepSize, RnStats& S) {  // LeapFrog.scala:168-176 (gradient reused)
  S.grads += 1;
  RN_UNROLL
  for (int i = 0; i < RN_N; i++) s.p[i] += stepSize * s.g[i];
}
RN_DEVICE void rn_new_qs(const RnArgs& A, int c, const RnMass& M, RnPQ& s, double stepSize) {  // :147-154
  double v[RN_N];
  rn_velocity(A, c, M, s.p, v);
  RN_UNROLL
  for (int i = 0; i < RN_N; i++) s.q[i] += (stepSize * v[i]);
}
// initialHalfThenFullStep + (l-1) twoFullSteps + finalHalfStep, LeapFrog.scala:24-33,156-191.
// `g` must hold the gradient at s.q on entry (true for params and for every state this kernel produces).
RN_DEVICE void rn_leapfrog(const RnArgs& A, int c, const RnMass& M, RnPQ& s, int l, double stepSize, RnStats& S) {
  // (one rn_update call site: the emitted density is inlined exactly once per use of rn_leapfrog)
  rn_full_ps(s, stepSize / 2.0, S);
  for (int i = 0;;) {
    rn_new_qs(A, c, M, s, stepSize);
    rn_update(A, s, S);
    if (++i >= l) break;
    rn_full_ps(s, stepSize, S);
  }
  rn_full_ps(s, stepSize / 2.0, S);
}
RN_DEVICE void rn_take_steps(const RnArgs& A, int c, const RnMass& M, RnPQ& s, int l, double stepSize, RnStats& S) {
  rn_ring_add(A, c, S, 0, stepSize);  // stats.stepSizes.add, LeapFrog.scala:25
  rn_leapfrog(A, c, M, s, l, stepSize, S);
  S.steps += l;
}

// momentum draw, LeapFrog.scala:233-255
RN_DEVICE void rn_initialize_ps(const RnArgs& A, int c, const RnMass& M, RnRng& rng, double (&p)[RN_N]) {
  (void)A;
  (void)c;
  double z[RN_N];
  for (int i = 0; i < RN_N; i++) z[i] = rn_normal(rng);
#if RN_MASS_MAX >= 2
  if (M.kind == 2) {  // DenseMassMatrix.upperTriangularSolve, MassMatrix.scala:55-72
    int i = RN_N - 1;
    int m = ((i + 1) * (i + 2)) / 2 - 1;
    while (i >= 0) {
      int j = RN_N - 1;
      double dot = 0.0;
      while (j > i) {
        dot += p[j] * RN_AT(A.chol, m, c);
        j -= 1;
        m -= 1;
      }
      p[i] = (z[i] - dot) / RN_AT(A.chol, m, c);
      i -= 1;
      m -= 1;
    }
    return;
  }
#endif
#if RN_MASS_MAX >= 1
  if (M.kind == 1) {
    RN_UNROLL
    for (int i = 0; i < RN_N; i++) p[i] = z[i] / sqrt(M.m[i]);  // buf(i) / stdDevs(i), stdDevs = sqrt(elements)
    return;
  }
#endif
  RN_UNROLL
  for (int i = 0; i < RN_N; i++) p[i] = z[i];
}

RN_DEVICE void rn_load_mass(const RnArgs& A, int c, RnMass& M) {
  (void)A;
  (void)c;
#if RN_MASS_MAX >= 1
  if (M.kind == 1) {
    RN_UNROLL
    for (int i = 0; i < RN_N; i++) M.m[i] = RN_AT(A.mass, i, c);
  }
#endif
}

RN_DEVICE void rn_load_stats(const RnArgs& A, int c, RnStats& S) {
  S.grads = A.st_grads[c];
  S.steps = A.st_steps[c];
  S.iters = A.st_iters[c];
  S.accepted = A.st_accepted[c];
  S.err = A.st_err[c];
  S.e_mean = RN_AT(A.st_energy, 0, c);
  S.e_raw = RN_AT(A.st_energy, 1, c);
  S.trans2 = RN_AT(A.st_energy, 2, c);
  S.e_n = A.st_energy_n[c];
  for (int r = 0; r < 3; r++) {
    S.ring_i[r] = RN_AT(A.st_ring_i, r, c);
    S.ring_full[r] = RN_AT(A.st_ring_full, r, c);
  }
}
RN_DEVICE void rn_store_stats(const RnArgs& A, int c, const RnStats& S) {
  A.st_grads[c] = S.grads;
  A.st_steps[c] = S.steps;
  A.st_iters[c] = S.iters;
  A.st_accepted[c] = S.accepted;
  A.st_err[c] = S.err;
  RN_AT(A.st_energy, 0, c) = S.e_mean;
  RN_AT(A.st_energy, 1, c) = S.e_raw;
  RN_AT(A.st_energy, 2, c) = S.trans2;
  A.st_energy_n[c] = S.e_n;
  for (int r = 0; r < 3; r++) {
    RN_AT(A.st_ring_i, r, c) = S.ring_i[r];
    RN_AT(A.st_ring_full, r, c) = S.ring_full[r];
  }
}

// =============================================================================================================
// rn_k_init: LeapFrog.initialize(IdentityMassMatrix) (Driver.scala:22) + stepSizeTuner.initialize (Driver.scala:60)
// =============================================================================================================
RN_GLOBAL void rn_k_init(const RnArgs A) {
  const int c = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (c >= A.chains) return;
  RnRng rng;
  rng.seed = A.rng_seed[c];
  rng.nng = A.rng_nng[c];
  rng.have = A.rng_have[c];
  RnStats S;
  rn_load_stats(A, c, S);
  RnMass M;
  M.kind = 0;

  // LeapFrog.initialize, LeapFrog.scala:102-116
  RnPQ s;
  for (int i = 0; i < RN_N; i++) {
    s.p[i] = 0.0;
    s.q[i] = rn_normal(rng);
  }
  rn_update(A, s, S);
  double cq[RN_N], cg[RN_N], cp[RN_N];
  double cU = s.U;
  RN_UNROLL
  for (int i = 0; i < RN_N; i++) {
    cq[i] = s.q[i];
    cg[i] = s.g[i];
  }
  rn_initialize_ps(A, c, M, rng, cp);

  // stepSizeTuner.initialize
  double stepSize;
  if (A.step_tuner == 0) {  // DualAvgTuner.findReasonableStepSize, DualAvg.scala:27-41 (IdentityMassMatrix)
    const double H0 = rn_energy(A, c, M, cp, cU);
    stepSize = 1.0;
    double lap;
    {
      RN_UNROLL
      for (int i = 0; i < RN_N; i++) { s.p[i] = cp[i]; s.q[i] = cq[i]; s.g[i] = cg[i]; }
      s.U = cU;
      rn_leapfrog(A, c, M, s, 1, stepSize, S);  // tryStepping, LeapFrog.scala:14-22
      lap = rn_log_accept(rn_energy(A, c, M, s.p, s.U) - H0);
    }
    const double exponent = (lap > -RN_LN2) ? 1.0 : -1.0;
    const double doubleOrHalf = (exponent > 0) ? 2.0 : 0.5;
    while (stepSize != 0.0 && (exponent * lap > -exponent * RN_LN2)) {
      stepSize *= doubleOrHalf;
      RN_UNROLL
      for (int i = 0; i < RN_N; i++) { s.p[i] = cp[i]; s.q[i] = cq[i]; s.g[i] = cg[i]; }
      s.U = cU;
      rn_leapfrog(A, c, M, s, 1, stepSize, S);
      lap = rn_log_accept(rn_energy(A, c, M, s.p, s.U) - H0);
    }
    // DualAvg.apply, DualAvg.scala:80-90
    RN_AT(A.da, 1, c) = rn_log(stepSize);
    RN_AT(A.da, 2, c) = 0.0;
    RN_AT(A.da, 3, c) = 0.0;
    RN_AT(A.da, 4, c) = rn_log(10 * stepSize);
    A.da_iter[c] = 0;
  } else {
    stepSize = A.static_step;
  }
  RN_AT(A.da, 0, c) = stepSize;

  RN_UNROLL
  for (int i = 0; i < RN_N; i++) {
    RN_AT(A.params, i, c) = cp[i];
    RN_AT(A.params, RN_N + i, c) = cq[i];
    RN_AT(A.grad, i, c) = cg[i];
  }
  RN_AT(A.params, 2 * RN_N, c) = cU;
  A.rng_seed[c] = rng.seed;
  A.rng_nng[c] = rng.nng;
  A.rng_have[c] = rng.have;
  rn_store_stats(A, c, S);
}

// =============================================================================================================
// rn_k_iter: A.n_iter iterations of Driver.warmup's loop (phase 0, Driver.scala:67-88) or of
// Driver.collectSamples (phase 1, Driver.scala:102-117)
// =============================================================================================================
RN_GLOBAL void rn_k_iter(const RnArgs A) {
  const int c = A.chain_begin + (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (c >= A.chain_end) return;
  RnRng rng;
  rng.seed = A.rng_seed[c];
  rng.nng = A.rng_nng[c];
  rng.have = A.rng_have[c];
  RnStats S;
  rn_load_stats(A, c, S);
  RnMass M;
  M.kind = A.mass_kind;
  rn_load_mass(A, c, M);

  // step size in force: warmup uses the tuner's running value; sampling uses stepSizeTuner.stepSize
  // (= rn_exp(logStepSizeBar) for DualAvg, Driver.scala:37 / DualAvg.scala:23-25)
  double stepSize = RN_AT(A.da, 0, c);
  double logStepSize = 0, logStepSizeBar = 0, avgError = 0, shrinkageTarget = 0;
  int daIter = 0;
  if (A.step_tuner == 0) {
    logStepSize = RN_AT(A.da, 1, c);
    logStepSizeBar = RN_AT(A.da, 2, c);
    avgError = RN_AT(A.da, 3, c);
    shrinkageTarget = RN_AT(A.da, 4, c);
    daIter = A.da_iter[c];
    if (A.phase == 1) stepSize = rn_exp(logStepSizeBar);
  }
  int win_size = A.win_size, win_i = A.win_i, win_j = A.win_j, est_samples = A.est_samples;
#if RN_ENABLE_EHMC
  int ring_i = 0, ring_full = 0;
  if (A.sampler == 1) {
    ring_i = A.ring_i[c];
    ring_full = A.ring_full[c];
  }
#endif

  for (int it = 0; it < A.n_iter; it++) {
    // ---------------- lf.startIteration, LeapFrog.scala:52-59 ----------------
    RnPQ s;
    RN_UNROLL
    for (int i = 0; i < RN_N; i++) s.p[i] = RN_AT(A.params, i, c);  // old momentum, for prevH
    const double cU = RN_AT(A.params, 2 * RN_N, c);
    const double prevH = rn_energy(A, c, M, s.p, cU);
    rn_initialize_ps(A, c, M, rng, s.p);
    RN_UNROLL
    for (int i = 0; i < RN_N; i++) {
      RN_AT(A.params, i, c) = s.p[i];  // initializePs writes into params (LeapFrog.scala:55); kept on reject
      s.q[i] = RN_AT(A.params, RN_N + i, c);
      s.g[i] = RN_AT(A.grad, i, c);
    }
    s.U = cU;
    const double startH = rn_energy(A, c, M, s.p, cU);  // finishIteration's energy(params), :62
    const rn_i64 iterationStartGrads = S.grads;
    const rn_i64 steps0 = S.steps;
    const double usedStep = stepSize;

    // ---------------- sampler.warmup / sampler.run ----------------
    if (A.sampler == 0) {  // HMCSampler, HMC.scala:6-23
      rn_take_steps(A, c, M, s, A.n_steps, stepSize, S);
    }
#if RN_ENABLE_EHMC
    else {  // EHMCSampler, EHMC.scala:15-61
      bool count = false;
      if (A.phase == 0) count = (!ring_full) || (rn_uniform(rng) < A.p_count);  // shouldCountSteps, :29-30
      if (count) {  // countSteps, :32-50
        RnPQ snap;
        int l = 0;
        for (;;) {
          double out = 0.0;  // lf.isUTurn(params), LeapFrog.scala:35-47
          RN_UNROLL
          for (int i = 0; i < RN_N; i++) out += (s.q[i] - RN_AT(A.params, RN_N + i, c)) * s.p[i];
          const bool uturn = (out != out) ? true : (out < 0);
          if (uturn || !(l < A.max_steps)) break;
          l += 1;
          rn_take_steps(A, c, M, s, 1, stepSize, S);
          if (l == A.min_steps) snap = s;
        }
        if (l < A.min_steps) {
          rn_take_steps(A, c, M, s, A.min_steps - l, stepSize, S);
        } else {
          s = snap;
        }
        // steps.add(l), Stats.scala:24-30
        ring_i += 1;
        if (ring_i == A.buf_size) ring_full = 1;
        ring_i = ring_i % A.buf_size;
        RN_AT(A.ring, ring_i, c) = (double)l;
      } else {  // steps.sample().toInt, Stats.scala:40-45
        const int idx = ring_full ? rn_rng_int(rng, A.buf_size) : rn_rng_int(rng, ring_i + 1);
        const int nsteps = rn_d2i(RN_AT(A.ring, idx, c));
        rn_take_steps(A, c, M, s, nsteps, stepSize, S);
      }
    }
#endif

    // ---------------- lf.finishIteration, LeapFrog.scala:61-82 ----------------
    const double endH = rn_energy(A, c, M, s.p, s.U);
    const double deltaH = endH - startH;
    const double a = rn_log_accept(deltaH);
    const bool accept = a > rn_log(rn_uniform(rng));
    double eH;
    if (accept) {
      RN_UNROLL
      for (int i = 0; i < RN_N; i++) {
        RN_AT(A.params, i, c) = s.p[i];
        RN_AT(A.params, RN_N + i, c) = s.q[i];
        RN_AT(A.grad, i, c) = s.g[i];
      }
      RN_AT(A.params, 2 * RN_N, c) = s.U;
      eH = endH;
      S.accepted += 1;
    } else {
      RN_UNROLL
      for (int i = 0; i < RN_N; i++) s.q[i] = RN_AT(A.params, RN_N + i, c);  // s.q := current position either way
      eH = startH;
    }
    {  // stats.energyVariance.update(eH); energyTransitions2 += pow(eH - prevH, 2)
      S.e_n += 1;
      const double oldDiff = eH - S.e_mean;
      S.e_mean += (oldDiff / (double)S.e_n);
      const double newDiff = eH - S.e_mean;
      S.e_raw += oldDiff * newDiff;
      const double d = eH - prevH;
      S.trans2 += d * d;
    }
    S.iters += 1;
    rn_ring_add(A, c, S, 1, rn_exp(a));
    rn_ring_add(A, c, S, 2, (double)(S.grads - iterationStartGrads));

    if (A.trace) {
      double* tr = A.trace + (size_t)it * 4 * (size_t)A.chains;
      tr[0 * (size_t)A.chains + c] = a;
      tr[1 * (size_t)A.chains + c] = accept ? 1.0 : 0.0;
      tr[2 * (size_t)A.chains + c] = usedStep;
      tr[3 * (size_t)A.chains + c] = (double)(S.steps - steps0);
    }

    if (A.phase == 0) {
      // ---------------- stepSizeTuner.update, Driver.scala:69 / DualAvg.scala:58-77 ----------------
      if (A.step_tuner == 0) {
        const double newAcceptanceProb = rn_exp(a);
        daIter = daIter + 1;
        const double avgErrorMultiplier = 1.0 / ((double)daIter + 10);
        const double stepSizeMultiplier = rn_pow((double)daIter, -0.75);
        avgError = ((1.0 - avgErrorMultiplier) * avgError + (avgErrorMultiplier * (A.delta - newAcceptanceProb)));
        logStepSize = (shrinkageTarget - (avgError * sqrt((double)daIter) / 0.05));
        logStepSizeBar = (stepSizeMultiplier * logStepSize + (1.0 - stepSizeMultiplier) * logStepSizeBar);
        stepSize = rn_exp(logStepSize);
      }
      // ---------------- massMatrixTuner.update(sample), Driver.scala:74-80 / MassMatrix.scala:147-164 -------
#if RN_MASS_MAX >= 1
      if (A.mass_tuner == 1 || A.mass_tuner == 2) {
        win_j += 1;
        if (A.adaptation == 1) {
          // pooled extension: plain window sums per chain; the host reduces them over chains (and ranks) at the
          // window end (rn_k_pool_reduce / rn_k_pool_apply) -- launches are cut at window ends in this mode
          if (!(win_j < A.skip_first || (A.total_warmup - win_j) < A.skip_last)) {
            win_i += 1;
            RN_UNROLL
            for (int i = 0; i < RN_N; i++) {
              RN_AT(A.est_mean, i, c) += s.q[i];
              RN_AT(A.est_raw, i, c) += s.q[i] * s.q[i];
            }
            if (win_i == win_size) {
              win_i = 0;
              win_size = rn_d2i(win_size * A.win_expansion);
            }
          }
        } else if (!(win_j < A.skip_first || (A.total_warmup - win_j) < A.skip_last)) {
          win_i += 1;
          est_samples += 1;  // VarianceEstimator.update, MassMatrixEstimator.scala:69-83
          double oldDiff[RN_N], newDiff[RN_N];
          RN_UNROLL
          for (int i = 0; i < RN_N; i++) {
            double mean = RN_AT(A.est_mean, i, c);
            oldDiff[i] = s.q[i] - mean;
            mean += (oldDiff[i] / (double)est_samples);
            newDiff[i] = s.q[i] - mean;
            RN_AT(A.est_mean, i, c) = mean;
            RN_AT(A.est_raw, i, c) += oldDiff[i] * newDiff[i];
          }
#if RN_MASS_MAX >= 2
          if (A.mass_tuner == 2) {  // CovarianceEstimator.update, :28-41
            for (int j = 0; j < RN_N; j++)
              for (int k = 0; k < RN_N; k++) RN_AT(A.est_cov, j * RN_N + k, c) += newDiff[j] * oldDiff[k];
          }
#endif
          if (win_i == win_size) {
            win_i = 0;
            win_size = rn_d2i(win_size * A.win_expansion);
            if (A.mass_tuner == 1) {  // DiagonalMassMatrix(variance()), :92-103
              M.kind = 1;
              RN_UNROLL
              for (int i = 0; i < RN_N; i++) {
                const double v = RN_AT(A.est_raw, i, c) / (double)est_samples;
                if (v == 0.0) S.err |= 2;  // require(!elements.contains(0.0)), MassMatrix.scala:8
                M.m[i] = v;
                RN_AT(A.mass, i, c) = v;
                RN_AT(A.est_mean, i, c) = 0.0;  // reset(): mean/raw only, NOT samples (:60-67)
                RN_AT(A.est_raw, i, c) = 0.0;
              }
            }
#if RN_MASS_MAX >= 2
            else {  // DenseMassMatrix(covariance()), :43-50 + Cholesky MassMatrix.scala:76-117
              M.kind = 2;
              const double z = (double)(est_samples - 1);
              for (int i = 0; i < RN_N * RN_N; i++) {
                const double v = RN_AT(A.est_cov, i, c) / z;
                if (v == 0.0) S.err |= 2;
                RN_AT(A.mass, i, c) = v;
                RN_AT(A.est_cov, i, c) = 0.0;
              }
              for (int i = 0; i < RN_N; i++) {
                RN_AT(A.est_mean, i, c) = 0.0;
                RN_AT(A.est_raw, i, c) = 0.0;
              }
              double lower[(RN_N * (RN_N + 1)) / 2];
              int l = 0;
              for (int i = 0; i < RN_N; i++)
                for (int k = 0; k <= i; k++) {
                  double sum = 0.0;
                  for (int j = 0; j < k; j++) sum += lower[(i * (i + 1)) / 2 + j] * lower[(k * (k + 1)) / 2 + j];
                  const double x = RN_AT(A.mass, i * RN_N + k, c) - sum;
                  if (i == k)
                    lower[l] = sqrt(x);
                  else {
                    const double diag = lower[((k + 1) * (k + 2)) / 2 - 1];
                    lower[l] = (1.0 / diag * x);
                  }
                  l += 1;
                }
              l = 0;
              for (int i = 0; i < RN_N; i++)
                for (int k = 0; k < (RN_N - i); k++) {
                  RN_AT(A.chol, l, c) = lower[((k + i) * (k + i + 1)) / 2 + i];
                  l += 1;
                }
            }
#endif
            // stepSize = stepSizeTuner.reset(), Driver.scala:78 / DualAvg.scala:17-21
            if (A.step_tuner == 0) {
              const double ss = rn_exp(logStepSizeBar);
              logStepSize = rn_log(ss);
              logStepSizeBar = 0.0;
              avgError = 0.0;
              daIter = 0;
              shrinkageTarget = rn_log(10 * ss);
              stepSize = ss;
            }
          }
        }
      }
#endif
    } else if (A.samples) {  // lf.variables(params, output), Driver.scala:105-107
      double* out = A.samples + (size_t)it * RN_N * (size_t)A.chains;
      RN_UNROLL
      for (int i = 0; i < RN_N; i++) out[(size_t)i * (size_t)A.chains + c] = s.q[i];
    }
  }

  if (A.phase == 0) {
    RN_AT(A.da, 0, c) = stepSize;
    if (A.step_tuner == 0) {
      RN_AT(A.da, 1, c) = logStepSize;
      RN_AT(A.da, 2, c) = logStepSizeBar;
      RN_AT(A.da, 3, c) = avgError;
      RN_AT(A.da, 4, c) = shrinkageTarget;
      A.da_iter[c] = daIter;
    }
  }
#if RN_ENABLE_EHMC
  if (A.sampler == 1) {
    A.ring_i[c] = ring_i;
    A.ring_full[c] = ring_full;
  }
#endif
  A.rng_seed[c] = rng.seed;
  A.rng_nng[c] = rng.nng;
  A.rng_have[c] = rng.have;
  rn_store_stats(A, c, S);
}

// =============================================================================================================
// rn_k_density: DensityFunction.update/density/gradient for a batch of positions (Model.scala:38-50).
// q: [N][chains] ; out: [N+1][chains] = density, gradient
// =============================================================================================================
RN_GLOBAL void rn_k_density(const double* RN_RESTRICT qin, double* RN_RESTRICT out, const double* data, int* err,
                            int chains) {
  const int c = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (c >= chains) return;
  double q[RN_N], g[RN_N], dens;
  RN_UNROLL
  for (int i = 0; i < RN_N; i++) q[i] = qin[(size_t)i * chains + c];
  int e = 0;
  rn_density(q, dens, g, data, e);
  out[c] = dens;
  RN_UNROLL
  for (int i = 0; i < RN_N; i++) out[(size_t)(i + 1) * chains + c] = g[i];
#ifdef RN_HOST_EMULATION
  if (e) *err |= e;
#else
  if (e) atomicOr(err, e);
#endif
}

// =============================================================================================================
// rn_k_transpose: [rows][cols] -> [cols][rows] (sample chunks [iter][n][chain] -> [chain][iter][n] before the
// device->host copy of rn_sample).  32x32 tiles through shared memory, both sides coalesced.
// =============================================================================================================
#ifndef RN_HOST_EMULATION
RN_GLOBAL void rn_k_transpose(const double* RN_RESTRICT src, double* RN_RESTRICT dst, int rows, int cols,
                              long long src_ld, long long dst_ld, long long dst_off) {
  // dst[c * dst_ld + dst_off + r] = src[r * src_ld + c]   (a block of `cols` chains out of src_ld)
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[(size_t)r * (size_t)src_ld + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[(size_t)c * (size_t)dst_ld + (size_t)dst_off + r] = tile[threadIdx.x][j];
  }
}
#endif

// =============================================================================================================
// Pooled mass-matrix adaptation (RN_ADAPT_POOLED; an extension, not reference semantics): at a window end the
// per-chain window sums {sum q_i, sum q_i^2} are reduced over all chains of this GPU into pool[1..2n] (pool[0] =
// number of draws), all-reduced over ranks by the host (NCCL) and applied to every chain: one shared diagonal
// mass matrix, sums cleared, DualAvg restarted from each chain's averaged step size (Driver.scala:75-80).
// =============================================================================================================
#ifndef RN_HOST_EMULATION
RN_GLOBAL void rn_k_pool_reduce(const RnArgs A, double* pool, int window_len) {
  __shared__ double red[256];
  for (int i = 0; i < 2 * RN_N; i++) {
    const double* src = (i < RN_N) ? (A.est_mean + (size_t)i * A.chains) : (A.est_raw + (size_t)(i - RN_N) * A.chains);
    double acc = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < A.chains; c += gridDim.x * blockDim.x) acc += src[c];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(&pool[1 + i], red[0]);
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&pool[0], (double)A.chains * (double)window_len);
}
RN_GLOBAL void rn_k_pool_apply(const RnArgs A, const double* pool) {
  const int c = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (c >= A.chains) return;
  const double cnt = pool[0];
  for (int i = 0; i < RN_N; i++) {
    const double mean = pool[1 + i] / cnt;
    const double var = pool[1 + RN_N + i] / cnt - mean * mean;
    if (!(var > 0.0)) A.st_err[c] |= 2;
    RN_AT(A.mass, i, c) = var;
    RN_AT(A.est_mean, i, c) = 0.0;
    RN_AT(A.est_raw, i, c) = 0.0;
  }
  if (A.step_tuner == 0) {  // stepSizeTuner.reset(), DualAvg.scala:17-21
    const double ss = rn_exp(RN_AT(A.da, 2, c));
    RN_AT(A.da, 0, c) = ss;
    RN_AT(A.da, 1, c) = rn_log(ss);
    RN_AT(A.da, 2, c) = 0.0;
    RN_AT(A.da, 3, c) = 0.0;
    RN_AT(A.da, 4, c) = rn_log(10 * ss);
    A.da_iter[c] = 0;
  }
}
#endif

#endif  // RN_SAMPLER_CUH
