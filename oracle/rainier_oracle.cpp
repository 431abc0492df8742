/*
 * oracle/rainier_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement of Rainier's HMC hot path, kept structurally line-for-line with the reference so it
 * can serve as the parity oracle for the CUDA path and as the labelled "port" CPU baseline in bench.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the
 * resulting librainier_oracle.so; the product (rainier_b200/csrc) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_goldsets.py drives this file, through the Python DAG restatement in
 * oracle/rainier_py, against the reference's own golden vectors
 * (rainier-test/src/main/scala/com/stripe/rainier/core/SBCModel.scala:46-267, tolerance 1e-10 as in
 * rainier-test/src/test/scala/com/stripe/rainier/core/SBCTest.scala:7-15).
 *
 * Each block cites the reference file:line it follows; S/ = rainier-sampler/src/main/scala/com/stripe/rainier/sampler/,
 * IR/ = rainier-compute/src/main/scala/com/stripe/rainier/ir/, K/ = rainier-core/src/main/scala/com/stripe/rainier/core/.
 *
 * Built with -ffp-contract=off: the JVM never contracts a*b+c into an FMA.
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../include/rainier_cuda.h"
#include "../include/rainier_rir.h"
#include "jmath.h"

namespace rno {

static thread_local std::string g_last_error;
static int g_threads = 0;

static int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

/* =====================================================================================================
 * RNG  (S/RNG.scala:6-26)
 * ===================================================================================================== */
struct RNG {
  JRandom rand;
  double standardUniform() { return rand.next_double(); } /* S/RNG.scala:24 */
  double standardNormal() { return rand.next_gaussian(); } /* S/RNG.scala:25 */
  int int_(int until) {                                    /* S/RNG.scala:9-10 */
    int v = jd2i(standardUniform() * until);
    return std::min(v, until - 1);
  }
};

/* =====================================================================================================
 * The frozen DAG and DataFunction  (IR/IR.scala:3-23, IR/DataFunction.scala:13-85, K/Model.scala:38-50)
 * ===================================================================================================== */
struct Target {
  uint64_t n_rows;
  uint32_t first_input, n_cols;
  std::vector<uint32_t> outputs;
  std::vector<int32_t> row_nodes; /* nodes that depend on this target's columns, topological order */
};

struct Model {
  rir_header h;
  std::vector<rir_node> nodes;
  std::vector<int32_t> lookup_refs;
  std::vector<Target> targets;
  std::vector<std::vector<double>> cols; /* indexed by input - n_params */
  std::vector<int32_t> once_nodes;       /* column-independent nodes, topological order */
  int n() const { return (int)h.n_params; }
};

static int parse_rir(const void* data, size_t len, Model& m) {
  const uint8_t* p = (const uint8_t*)data;
  const uint8_t* end = p + len;
  if (len < sizeof(rir_header)) return fail(RN_E_INVALID, "RIR: truncated header");
  std::memcpy(&m.h, p, sizeof(rir_header));
  p += sizeof(rir_header);
  if (m.h.magic != RIR_MAGIC || m.h.version != RIR_VERSION) return fail(RN_E_INVALID, "RIR: bad magic/version");
  if ((size_t)(end - p) < (size_t)m.h.n_nodes * sizeof(rir_node)) return fail(RN_E_INVALID, "RIR: truncated nodes");
  m.nodes.resize(m.h.n_nodes);
  std::memcpy(m.nodes.data(), p, (size_t)m.h.n_nodes * sizeof(rir_node));
  p += (size_t)m.h.n_nodes * sizeof(rir_node);
  size_t lr_bytes = ((size_t)m.h.n_lookup_refs * 4 + 7) & ~(size_t)7;
  if ((size_t)(end - p) < lr_bytes) return fail(RN_E_INVALID, "RIR: truncated lookup refs");
  m.lookup_refs.resize(m.h.n_lookup_refs);
  std::memcpy(m.lookup_refs.data(), p, (size_t)m.h.n_lookup_refs * 4);
  p += lr_bytes;
  m.targets.resize(m.h.n_targets);
  for (uint32_t t = 0; t < m.h.n_targets; t++) {
    rir_target rt;
    if ((size_t)(end - p) < sizeof(rt)) return fail(RN_E_INVALID, "RIR: truncated target");
    std::memcpy(&rt, p, sizeof(rt));
    p += sizeof(rt);
    size_t ob = ((size_t)rt.n_outputs * 4 + 7) & ~(size_t)7;
    if ((size_t)(end - p) < ob) return fail(RN_E_INVALID, "RIR: truncated outputs");
    Target& T = m.targets[t];
    T.n_rows = rt.n_rows;
    T.first_input = rt.first_input;
    T.n_cols = rt.n_cols;
    T.outputs.resize(rt.n_outputs);
    std::memcpy(T.outputs.data(), p, (size_t)rt.n_outputs * 4);
    p += ob;
    for (uint32_t o : T.outputs)
      if (o >= m.h.n_nodes) return fail(RN_E_INVALID, "RIR: output id out of range");
  }
  /* validate topological order + operand ranges */
  for (uint32_t i = 0; i < m.h.n_nodes; i++) {
    const rir_node& nd = m.nodes[i];
    auto ok = [&](int32_t x) { return x >= 0 && (uint32_t)x < i; };
    switch (nd.kind) {
      case RIR_INPUT:
        if (nd.a < 0 || (uint32_t)nd.a >= m.h.n_inputs) return fail(RN_E_INVALID, "RIR: input index out of range");
        break;
      case RIR_CONST: break;
      case RIR_UNARY:
        if (!ok(nd.a)) return fail(RN_E_INVALID, "RIR: unary operand not defined before use");
        break;
      case RIR_BINARY:
        if (!ok(nd.a) || !ok(nd.b)) return fail(RN_E_INVALID, "RIR: binary operand not defined before use");
        break;
      case RIR_LOOKUP:
        if (!ok(nd.a) || nd.b < 0 || nd.c <= 0 || (uint32_t)(nd.b + nd.c) > m.h.n_lookup_refs)
          return fail(RN_E_INVALID, "RIR: bad lookup");
        for (int k = 0; k < nd.c; k++)
          if (!ok(m.lookup_refs[nd.b + k])) return fail(RN_E_INVALID, "RIR: lookup ref not defined before use");
        break;
      default: return fail(RN_E_INVALID, "RIR: unknown node kind");
    }
  }
  return RN_OK;
}

/* Classify nodes: which target's columns (if any) each node depends on.  Values do not depend on this --
 * DataFunction.compute (IR/DataFunction.scala:48-84) re-evaluates everything per row -- it only avoids
 * recomputing row-invariant subexpressions, which are pure functions of the parameters. */
static int classify(Model& m) {
  const int N = (int)m.h.n_nodes;
  std::vector<int> dep(N, -1); /* -1: parameters/constants only; t: target t's columns */
  auto target_of_input = [&](int inp) -> int {
    if (inp < (int)m.h.n_params) return -1;
    for (size_t t = 0; t < m.targets.size(); t++)
      if ((uint32_t)inp >= m.targets[t].first_input && (uint32_t)inp < m.targets[t].first_input + m.targets[t].n_cols)
        return (int)t;
    return -2;
  };
  for (int i = 0; i < N; i++) {
    const rir_node& nd = m.nodes[i];
    int d = -1;
    auto merge = [&](int x) {
      if (x == -1) return true;
      if (d == -1 || d == x) {
        d = x;
        return true;
      }
      return false;
    };
    bool ok = true;
    switch (nd.kind) {
      case RIR_INPUT: {
        int t = target_of_input(nd.a);
        if (t == -2) return fail(RN_E_INVALID, "RIR: column input not owned by any target");
        d = t;
        break;
      }
      case RIR_CONST: break;
      case RIR_UNARY: ok = merge(dep[nd.a]); break;
      case RIR_BINARY: ok = merge(dep[nd.a]) && merge(dep[nd.b]); break;
      case RIR_LOOKUP:
        ok = merge(dep[nd.a]);
        for (int k = 0; k < nd.c && ok; k++) ok = merge(dep[m.lookup_refs[nd.b + k]]);
        break;
    }
    if (!ok) return fail(RN_E_INVALID, "RIR: node mixes columns of two targets");
    dep[i] = d;
  }
  /* reachability per target */
  for (size_t t = 0; t < m.targets.size(); t++) {
    std::vector<char> need(N, 0);
    for (uint32_t o : m.targets[t].outputs) need[o] = 1;
    for (int i = N - 1; i >= 0; i--) {
      if (!need[i]) continue;
      const rir_node& nd = m.nodes[i];
      switch (nd.kind) {
        case RIR_UNARY: need[nd.a] = 1; break;
        case RIR_BINARY: need[nd.a] = need[nd.b] = 1; break;
        case RIR_LOOKUP:
          need[nd.a] = 1;
          for (int k = 0; k < nd.c; k++) need[m.lookup_refs[nd.b + k]] = 1;
          break;
        default: break;
      }
    }
    for (int i = 0; i < N; i++)
      if (need[i] && dep[i] >= 0) {
        if (dep[i] != (int)t) return fail(RN_E_INVALID, "RIR: target reads another target's columns");
        m.targets[t].row_nodes.push_back(i);
      }
  }
  for (int i = 0; i < N; i++)
    if (dep[i] == -1) m.once_nodes.push_back(i);
  return RN_OK;
}

/* S/DensityFunction.scala:3-8 */
struct DensityFunction {
  virtual ~DensityFunction() {}
  virtual int nVars() const = 0;
  virtual void update(const double* vars) = 0;
  virtual double density() const = 0;
  virtual double gradient(int index) const = 0;
};

/* K/Model.scala:38-50 over IR/DataFunction.scala:32-84, evaluating the flat SSA form with the op semantics of
 * IR/MethodGenerator.scala:56-94,130-167 and IR/ExprMethodGenerator.scala:39-63. */
struct RirDensity : DensityFunction {
  const Model& m;
  std::vector<double> vals;    /* one slot per node ("globals" + locals + stack of the generated code) */
  std::vector<double> inputs;  /* K/Model.scala:41 */
  std::vector<double> outputs; /* K/Model.scala:43: numOutputs = nVars+1 (or 1 for primal-only RIR) */
  int lookup_error = 0;

  explicit RirDensity(const Model& mm) : m(mm) {
    vals.assign(m.h.n_nodes, 0.0);
    inputs.assign(m.h.n_inputs, 0.0);
    outputs.assign((m.h.flags & RIR_FLAG_GRADIENT) ? m.h.n_params + 1 : 1, 0.0);
  }
  int nVars() const override { return m.n(); }

  inline void eval_node(int i) {
    const rir_node& nd = m.nodes[i];
    double r = 0.0;
    switch (nd.kind) {
      case RIR_INPUT: r = inputs[nd.a]; break; /* loadParameter, IR/MethodGenerator.scala:50-54 */
      case RIR_CONST: r = nd.value; break;
      case RIR_UNARY: {
        double x = vals[nd.a];
        switch (nd.op) { /* IR/MethodGenerator.scala:74-94 : java.lang.Math.* */
          case RIR_U_EXP: r = jexp(x); break;
          case RIR_U_LOG: r = jlog(x); break;
          case RIR_U_ABS: r = std::fabs(x); break;
          case RIR_U_NOOP: r = x; break;
          case RIR_U_SIN: r = std::sin(x); break;
          case RIR_U_COS: r = std::cos(x); break;
          case RIR_U_TAN: r = std::tan(x); break;
          case RIR_U_ASIN: r = std::asin(x); break;
          case RIR_U_ACOS: r = std::acos(x); break;
          case RIR_U_ATAN: r = std::atan(x); break;
        }
        break;
      }
      case RIR_BINARY: {
        double x = vals[nd.a], y = vals[nd.b];
        switch (nd.op) { /* IR/MethodGenerator.scala:56-72 */
          case RIR_B_ADD: r = x + y; break;
          case RIR_B_MUL: r = x * y; break;
          case RIR_B_SUB: r = x - y; break;
          case RIR_B_DIV: r = x / y; break;
          case RIR_B_POW: r = jpow(x, y); break;
          case RIR_B_COMPARE: r = jcompare(x, y); break;
        }
        break;
      }
      case RIR_LOOKUP: { /* D2I ; tableswitch low..low+len-1 ; default -> throw (IR/ExprMethodGenerator.scala:50-56) */
        int idx = jd2i(vals[nd.a]) - nd.d;
        if (idx < 0 || idx >= nd.c) {
          lookup_error = 1;
          r = std::numeric_limits<double>::quiet_NaN();
        } else {
          r = vals[m.lookup_refs[nd.b + idx]];
        }
        break;
      }
    }
    vals[i] = r;
  }

  void update(const double* vars) override {
    const int n = m.n();
    std::memcpy(inputs.data(), vars, sizeof(double) * n); /* K/Model.scala:45 */
    const int numOutputs = (int)outputs.size();
    for (int k = 0; k < numOutputs; k++) outputs[k] = 0.0; /* IR/DataFunction.scala:35-39 */
    for (int i : m.once_nodes) eval_node(i);
    for (size_t t = 0; t < m.targets.size(); t++) { /* IR/DataFunction.scala:41-45 */
      const Target& T = m.targets[t];
      if (T.n_cols > 0) { /* IR/DataFunction.scala:56-73 */
        for (uint64_t k = 0; k < T.n_rows; k++) {
          for (uint32_t j = 0; j < T.n_cols; j++)
            inputs[T.first_input + j] = m.cols[T.first_input - n + j][k]; /* :61-63 */
          for (int i : T.row_nodes) eval_node(i);
          for (int o = 0; o < numOutputs; o++) outputs[o] += vals[T.outputs[o]]; /* :65-71 */
        }
      } else { /* IR/DataFunction.scala:74-83 */
        for (int o = 0; o < numOutputs; o++) outputs[o] += vals[T.outputs[o]];
      }
    }
  }
  double density() const override { return outputs[0]; }                   /* K/Model.scala:48 */
  double gradient(int index) const override { return outputs[index + 1]; } /* K/Model.scala:49 */
};

/* =====================================================================================================
 * Stats / RingBuffer / estimators  (S/Stats.scala:3-59, S/MassMatrixEstimator.scala:9-112)
 * ===================================================================================================== */
struct RingBuffer { /* S/Stats.scala:19-59 */
  bool full = false;
  int i = 0;
  int size;
  std::vector<double> buf;
  explicit RingBuffer(int sz) : size(sz), buf(sz, 0.0) {}
  void add(double value) { /* :24-30 */
    i += 1;
    if (i == size) full = true;
    i = i % size;
    buf[i] = value;
  }
  double sample(RNG& rng) { /* :40-45 */
    if (full) return buf[rng.int_(size)];
    return buf[rng.int_(i + 1)];
  }
  double mean() const { /* :47-58 */
    double sum = 0.0;
    for (int j = 0; j < size; j++) sum += buf[j];
    if (full) return sum / (double)size;
    return sum / (double)i;
  }
};

struct MassMatrix { /* S/MassMatrix.scala:3-32 */
  int kind = RN_MATRIX_IDENTITY;
  std::vector<double> elements;
  std::vector<double> stdDevs;                 /* Diagonal :10-12 */
  std::vector<double> choleskyUpperTriangular; /* Dense :18-19 */
  bool invalid = false;                        /* require(!elements.contains(0.0)) failed */
};

namespace dense { /* S/MassMatrix.scala:34-117 */
static void squareMultiply(const std::vector<double>& matrix, const double* vector, double* out, int n) {
  for (int i = 0; i < n; i++) {
    double y = 0.0;
    for (int j = 0; j < n; j++) y += vector[j] * matrix[(i * n) + j];
    out[i] = y;
  }
}
static int triangleNumber(int k) { return (k * (k + 1)) / 2; }
static void upperTriangularSolve(const std::vector<double>& packed, const double* vector, double* out, int size) {
  int i = size - 1;
  int m = triangleNumber(i + 1) - 1;
  while (i >= 0) {
    int j = size - 1;
    double dot = 0.0;
    while (j > i) {
      dot += out[j] * packed[m];
      j -= 1;
      m -= 1;
    }
    out[i] = (vector[i] - dot) / packed[m];
    i -= 1;
    m -= 1;
  }
}
static int matrixSize(size_t elements) { return (int)std::floor(std::sqrt((double)elements)); }
static std::vector<double> choleskyUpperTriangular(const std::vector<double>& matrix) {
  int n = matrixSize(matrix.size());
  std::vector<double> lower(triangleNumber(n), 0.0);
  int i = 0, l = 0;
  while (i < n) {
    int k = 0;
    while (k <= i) {
      double sum = 0.0;
      int j = 0;
      while (j < k) {
        sum += lower[triangleNumber(i) + j] * lower[triangleNumber(k) + j];
        j += 1;
      }
      double x = matrix[(i * n) + k] - sum;
      if (i == k)
        lower[l] = std::sqrt(x);
      else {
        double diag = lower[triangleNumber(k + 1) - 1];
        lower[l] = (1.0 / diag * x);
      }
      k += 1;
      l += 1;
    }
    i += 1;
  }
  std::vector<double> upper(lower.size(), 0.0);
  i = 0;
  l = 0;
  while (i < n) {
    int k = 0;
    while (k < (n - i)) {
      upper[l] = lower[triangleNumber(k + i) + i];
      k += 1;
      l += 1;
    }
    i += 1;
  }
  return upper;
}
} /* namespace dense */

static MassMatrix DiagonalMassMatrix(const std::vector<double>& elements) { /* S/MassMatrix.scala:7-13 */
  MassMatrix m;
  m.kind = RN_MATRIX_DIAGONAL;
  m.elements = elements;
  for (double x : elements) {
    if (x == 0.0) m.invalid = true;
    m.stdDevs.push_back(std::sqrt(x));
  }
  return m;
}
static MassMatrix DenseMassMatrix(const std::vector<double>& elements) { /* S/MassMatrix.scala:15-32 */
  MassMatrix m;
  m.kind = RN_MATRIX_DENSE;
  m.elements = elements;
  for (double x : elements)
    if (x == 0.0) m.invalid = true;
  m.choleskyUpperTriangular = dense::choleskyUpperTriangular(elements);
  return m;
}

struct VarianceEstimator { /* S/MassMatrixEstimator.scala:52-112 */
  int size;
  int samples = 0;
  std::vector<double> mean, raw, oldDiff, newDiff;
  explicit VarianceEstimator(int sz) : size(sz), mean(sz, 0.0), raw(sz, 0.0), oldDiff(sz, 0.0), newDiff(sz, 0.0) {}
  void reset() { /* :60-67  (does NOT reset `samples`) */
    for (int i = 0; i < size; i++) {
      mean[i] = 0.0;
      raw[i] = 0.0;
    }
  }
  void diff(const double* sample, std::vector<double>& buf) { /* :105-111 */
    for (int i = 0; i < size; i++) buf[i] = sample[i] - mean[i];
  }
  void update(const double* sample) { /* :69-83 */
    samples += 1;
    diff(sample, oldDiff);
    for (int i = 0; i < size; i++) mean[i] += (oldDiff[i] / (double)samples);
    diff(sample, newDiff);
    for (int j = 0; j < size; j++) raw[j] += oldDiff[j] * newDiff[j];
  }
  void update1(double sample) { update(&sample); } /* :86-90 */
  std::vector<double> variance() const {         /* :92-100 */
    std::vector<double> e(size);
    for (int i = 0; i < size; i++) e[i] = raw[i] / (double)samples;
    return e;
  }
};

struct CovarianceEstimator { /* S/MassMatrixEstimator.scala:9-50 */
  int size;
  VarianceEstimator variance;
  std::vector<double> cov;
  explicit CovarianceEstimator(int sz) : size(sz), variance(sz), cov((size_t)sz * sz, 0.0) {}
  void reset() {
    variance.reset();
    for (auto& c : cov) c = 0.0;
  }
  void update(const double* sample) {
    variance.update(sample);
    for (int j = 0; j < size; j++)
      for (int k = 0; k < size; k++) cov[j * size + k] += variance.newDiff[j] * variance.oldDiff[k];
  }
  std::vector<double> covariance() const {
    std::vector<double> e(cov.size());
    double z = (double)(variance.samples - 1);
    for (size_t i = 0; i < cov.size(); i++) e[i] = cov[i] / z;
    return e;
  }
};

struct Stats { /* S/Stats.scala:3-17 */
  int64_t gradientEvaluations = 0;
  int iterations = 0;
  int divergences = 0;
  RingBuffer stepSizes, acceptanceRates, gradsPerIteration;
  VarianceEstimator energyVariance;
  double energyTransitions2 = 0.0;
  /* extras for the ABI (not in the reference) */
  int64_t leapfrogSteps = 0;
  int accepted = 0;
  explicit Stats(int n) : stepSizes(n), acceptanceRates(n), gradsPerIteration(n), energyVariance(1) {}
};

/* =====================================================================================================
 * LeapFrog  (S/LeapFrog.scala:3-252)
 * ===================================================================================================== */
struct LeapFrog {
  DensityFunction& density;
  int statsWindow;
  Stats stats;
  int nVars, potentialIndex, inputOutputSize; /* :127-129 */
  std::vector<double> pqBuf, buf;             /* :131-132 */
  int64_t iterationStartGrads = 0;
  double prevH = 0.0;
  double lastLogAcceptanceProb = 0.0; /* test instrumentation */

  LeapFrog(DensityFunction& d, int sw)
      : density(d), statsWindow(sw), stats(sw), nVars(d.nVars()), potentialIndex(nVars * 2),
        inputOutputSize(nVars * 2 + 1), pqBuf(nVars * 2 + 1, 0.0), buf(nVars, 0.0) {}

  void resetStats() { stats = Stats(statsWindow); } /* :6-10 */

  double tryStepping(const std::vector<double>& params, double stepSize, const MassMatrix& mass) { /* :14-22 */
    copy(params, pqBuf);
    initialHalfThenFullStep(stepSize, mass);
    finalHalfStep(stepSize);
    double deltaH = energy(pqBuf, mass) - energy(params, mass);
    return logAcceptanceProb(deltaH);
  }

  void takeSteps(int l, double stepSize, const MassMatrix& mass) { /* :24-33 */
    stats.stepSizes.add(stepSize);
    initialHalfThenFullStep(stepSize, mass);
    int i = 1;
    while (i < l) {
      twoFullSteps(stepSize, mass);
      i += 1;
    }
    finalHalfStep(stepSize);
    stats.leapfrogSteps += l;
  }

  bool isUTurn(const std::vector<double>& params) { /* :35-47 */
    double out = 0.0;
    for (int i = 0; i < nVars; i++) out += (pqBuf[i + nVars] - params[i + nVars]) * pqBuf[i];
    if (out != out) return true;
    return out < 0;
  }

  void startIteration(std::vector<double>& params, const MassMatrix& mass, RNG& rng) { /* :52-59 */
    prevH = energy(params, mass);
    initializePs(params, mass, rng);
    copy(params, pqBuf);
    iterationStartGrads = stats.gradientEvaluations;
  }

  double finishIteration(std::vector<double>& params, const MassMatrix& mass, RNG& rng) { /* :61-82 */
    double startH = energy(params, mass);
    double endH = energy(pqBuf, mass);
    double deltaH = endH - startH;
    double a = logAcceptanceProb(deltaH);
    lastLogAcceptanceProb = a;
    if (a > jlog(rng.standardUniform())) {
      copy(pqBuf, params);
      stats.energyVariance.update1(endH);
      stats.energyTransitions2 += jpow(endH - prevH, 2);
      stats.accepted += 1;
    } else {
      stats.energyVariance.update1(startH);
      stats.energyTransitions2 += jpow(startH - prevH, 2);
    }
    stats.iterations += 1;
    stats.acceptanceRates.add(jexp(a));
    stats.gradsPerIteration.add((double)(stats.gradientEvaluations - iterationStartGrads));
    return a;
  }

  void snapshot(std::vector<double>& out) { copy(pqBuf, out); }     /* :84-85 */
  void restore(const std::vector<double>& in) { copy(in, pqBuf); } /* :87-88 */

  void variables(const std::vector<double>& params, double* out) { /* :91-97 */
    for (int i = 0; i < nVars; i++) out[i] = params[i + nVars];
  }

  std::vector<double> initialize(const MassMatrix& mass, RNG& rng) { /* :102-116 */
    std::vector<double> params(inputOutputSize, 0.0);
    std::fill(pqBuf.begin(), pqBuf.end(), 0.0);
    int i = nVars;
    int j = nVars * 2;
    while (i < j) {
      pqBuf[i] = rng.standardNormal();
      i += 1;
    }
    copyQsAndUpdateDensity();
    pqBuf[potentialIndex] = density.density() * -1;
    copy(pqBuf, params);
    initializePs(params, mass, rng);
    return params;
  }

  double energy(const std::vector<double>& params, const MassMatrix& mass) { /* :134-139 */
    double potential = params[potentialIndex];
    velocity(params, buf, mass);
    double kinetic = dot(buf, params) / 2.0;
    return potential + kinetic;
  }

  static double logAcceptanceProb(double deltaH) { /* :141-145 */
    if (deltaH != deltaH) return jlog(0.0);
    return jmin(-deltaH, 0.0);
  }

  void newQs(double stepSize, const MassMatrix& mass) { /* :147-154 */
    velocity(pqBuf, buf, mass);
    for (int i = 0; i < nVars; i++) pqBuf[i + nVars] += (stepSize * buf[i]);
  }
  void halfPsNewQs(double stepSize, const MassMatrix& mass) { /* :156-159 */
    fullPs(stepSize / 2.0);
    newQs(stepSize, mass);
  }
  void initialHalfThenFullStep(double stepSize, const MassMatrix& mass) { /* :161-166 */
    halfPsNewQs(stepSize, mass);
    copyQsAndUpdateDensity();
    pqBuf[potentialIndex] = density.density() * -1;
  }
  void fullPs(double stepSize) { /* :168-176 */
    copyQsAndUpdateDensity();
    for (int i = 0; i < nVars; i++) pqBuf[i] += stepSize * density.gradient(i);
  }
  void fullPsNewQs(double stepSize, const MassMatrix& mass) { /* :178-181 */
    fullPs(stepSize);
    newQs(stepSize, mass);
  }
  void twoFullSteps(double stepSize, const MassMatrix& mass) { /* :183-187 */
    fullPsNewQs(stepSize, mass);
    copyQsAndUpdateDensity();
    pqBuf[potentialIndex] = density.density() * -1;
  }
  void finalHalfStep(double stepSize) { fullPs(stepSize / 2.0); } /* :189-191 */

  void copy(const std::vector<double>& src, std::vector<double>& dst) { /* :193-195 */
    std::memcpy(dst.data(), src.data(), sizeof(double) * inputOutputSize);
  }
  void copyQsAndUpdateDensity() { /* :197-203 */
    std::memcpy(buf.data(), pqBuf.data() + nVars, sizeof(double) * nVars);
    density.update(buf.data());
    stats.gradientEvaluations += 1;
  }
  void velocity(const std::vector<double>& in, std::vector<double>& out, const MassMatrix& mass) { /* :205-219 */
    switch (mass.kind) {
      case RN_MATRIX_IDENTITY: std::memcpy(out.data(), in.data(), sizeof(double) * nVars); break;
      case RN_MATRIX_DIAGONAL:
        for (int i = 0; i < nVars; i++) out[i] = in[i] * mass.elements[i];
        break;
      case RN_MATRIX_DENSE: dense::squareMultiply(mass.elements, in.data(), out.data(), nVars); break;
    }
  }
  double dot(const std::vector<double>& x, const std::vector<double>& y) { /* :221-231 */
    double k = 0.0;
    int n = (int)x.size();
    for (int i = 0; i < n; i++) k += (x[i] * y[i]);
    return k;
  }
  void initializePs(std::vector<double>& params, const MassMatrix& mass, RNG& rng) { /* :233-255 */
    for (int i = 0; i < nVars; i++) buf[i] = rng.standardNormal();
    switch (mass.kind) {
      case RN_MATRIX_IDENTITY: std::memcpy(params.data(), buf.data(), sizeof(double) * nVars); break;
      case RN_MATRIX_DIAGONAL:
        for (int i = 0; i < nVars; i++) params[i] = buf[i] / mass.stdDevs[i];
        break;
      case RN_MATRIX_DENSE:
        dense::upperTriangularSolve(mass.choleskyUpperTriangular, buf.data(), params.data(), nVars);
        break;
    }
  }
};

/* =====================================================================================================
 * Samplers  (S/Sampler.scala:52-62, S/HMC.scala:3-24, S/EHMC.scala:3-62)
 * ===================================================================================================== */
struct Sampler {
  virtual ~Sampler() {}
  virtual void initialize(std::vector<double>& params, LeapFrog& lf, RNG& rng) = 0;
  virtual double warmup(std::vector<double>& params, LeapFrog& lf, double stepSize, const MassMatrix& mass, RNG& rng) = 0;
  virtual void run(std::vector<double>& params, LeapFrog& lf, double stepSize, const MassMatrix& mass, RNG& rng) = 0;
};

struct HMCSampler : Sampler { /* S/HMC.scala:3-24 */
  int nSteps;
  explicit HMCSampler(int n) : nSteps(n) {}
  void initialize(std::vector<double>&, LeapFrog&, RNG&) override {}
  double warmup(std::vector<double>& params, LeapFrog& lf, double stepSize, const MassMatrix& mass, RNG& rng) override {
    lf.startIteration(params, mass, rng);
    lf.takeSteps(nSteps, stepSize, mass);
    return lf.finishIteration(params, mass, rng);
  }
  void run(std::vector<double>& params, LeapFrog& lf, double stepSize, const MassMatrix& mass, RNG& rng) override {
    lf.startIteration(params, mass, rng);
    lf.takeSteps(nSteps, stepSize, mass);
    lf.finishIteration(params, mass, rng);
  }
};

struct EHMCSampler : Sampler { /* S/EHMC.scala:3-62 */
  int maxSteps, minSteps, bufSize;
  double pCount;
  RingBuffer steps;
  std::vector<double> buf;
  EHMCSampler(int maxS, int minS, int bufS, double pC) : maxSteps(maxS), minSteps(minS), bufSize(bufS), pCount(pC), steps(bufS) {}
  void initialize(std::vector<double>&, LeapFrog& lf, RNG&) override { buf.assign(lf.inputOutputSize, 0.0); } /* :11-13 */
  double warmup(std::vector<double>& params, LeapFrog& lf, double stepSize, const MassMatrix& mass, RNG& rng) override { /* :15-27 */
    lf.startIteration(params, mass, rng);
    if (shouldCountSteps(rng)) {
      countSteps(params, lf, stepSize, mass);
    } else {
      int n = jd2i(steps.sample(rng));
      lf.takeSteps(n, stepSize, mass);
    }
    return lf.finishIteration(params, mass, rng);
  }
  bool shouldCountSteps(RNG& rng) { return !steps.full || rng.standardUniform() < pCount; } /* :29-30 */
  void countSteps(std::vector<double>& params, LeapFrog& lf, double stepSize, const MassMatrix& mass) { /* :32-50 */
    int l = 0;
    while (!lf.isUTurn(params) && l < maxSteps) {
      l += 1;
      lf.takeSteps(1, stepSize, mass);
      if (l == minSteps) lf.snapshot(buf);
    }
    if (l < minSteps) {
      lf.takeSteps(minSteps - l, stepSize, mass);
    } else {
      lf.restore(buf);
    }
    steps.add((double)l);
  }
  void run(std::vector<double>& params, LeapFrog& lf, double stepSize, const MassMatrix& mass, RNG& rng) override { /* :52-61 */
    lf.startIteration(params, mass, rng);
    int n = jd2i(steps.sample(rng));
    lf.takeSteps(n, stepSize, mass);
    lf.finishIteration(params, mass, rng);
  }
};

/* =====================================================================================================
 * Step-size tuners  (S/Sampler.scala:29-40, S/DualAvg.scala:3-90)
 * ===================================================================================================== */
struct StepSizeTuner {
  virtual ~StepSizeTuner() {}
  virtual double initialize(std::vector<double>& params, LeapFrog& lf) = 0;
  virtual double update(double logAcceptanceProb) = 0;
  virtual double reset() = 0;
  virtual double stepSize() = 0;
};

struct StaticStepSize : StepSizeTuner { /* S/Sampler.scala:36-40 */
  double ss;
  explicit StaticStepSize(double s) : ss(s) {}
  double initialize(std::vector<double>&, LeapFrog&) override { return ss; }
  double update(double) override { return ss; }
  double reset() override { return ss; }
  double stepSize() override { return ss; }
};

struct DualAvg { /* S/DualAvg.scala:44-77 */
  double delta, logStepSize, logStepSizeBar, avgError;
  int iteration;
  double shrinkageTarget;
  double stepSizeUpdateDenom = 0.05;
  int acceptanceProbUpdateDenom = 10;
  double decayRate = 0.75;
  double stepSize() const { return jexp(logStepSize); }
  double finalStepSize() const { return jexp(logStepSizeBar); }
  void update(double logAcceptanceProb) { /* :58-77 */
    double newAcceptanceProb = jexp(logAcceptanceProb);
    iteration = iteration + 1;
    double avgErrorMultiplier = 1.0 / ((double)iteration + acceptanceProbUpdateDenom);
    double stepSizeMultiplier = jpow((double)iteration, -decayRate);
    avgError = ((1.0 - avgErrorMultiplier) * avgError + (avgErrorMultiplier * (delta - newAcceptanceProb)));
    logStepSize = (shrinkageTarget - (avgError * std::sqrt((double)iteration) / stepSizeUpdateDenom));
    logStepSizeBar = (stepSizeMultiplier * logStepSize + (1.0 - stepSizeMultiplier) * logStepSizeBar);
  }
  static DualAvg apply(double delta, double stepSize) { /* :80-90 */
    DualAvg d;
    d.delta = delta;
    d.logStepSize = jlog(stepSize);
    d.logStepSizeBar = 0.0;
    d.avgError = 0.0;
    d.iteration = 0;
    d.shrinkageTarget = jlog(10 * stepSize);
    return d;
  }
};

struct DualAvgTuner : StepSizeTuner { /* S/DualAvg.scala:3-42 */
  double delta;
  DualAvg da;
  explicit DualAvgTuner(double d) : delta(d) {}
  double initialize(std::vector<double>& params, LeapFrog& lf) override { /* :6-10 */
    MassMatrix identity;
    double stepSize0 = findReasonableStepSize(params, lf, identity);
    da = DualAvg::apply(delta, stepSize0);
    return stepSize0;
  }
  double update(double logAcceptanceProb) override { /* :12-15 */
    da.update(logAcceptanceProb);
    return da.stepSize();
  }
  double reset() override { /* :17-21 */
    double ss = stepSize();
    da = DualAvg::apply(delta, ss);
    return ss;
  }
  double stepSize() override { return da.finalStepSize(); } /* :23-25 */
  double findReasonableStepSize(std::vector<double>& params, LeapFrog& lf, const MassMatrix& mass) { /* :27-41 */
    double stepSize = 1.0;
    double logAcceptanceProb = lf.tryStepping(params, stepSize, mass);
    double exponent = (logAcceptanceProb > jlog(0.5)) ? 1.0 : -1.0;
    double doubleOrHalf = jpow(2, exponent);
    while (stepSize != 0.0 && (exponent * logAcceptanceProb > -exponent * jlog(2))) {
      stepSize *= doubleOrHalf;
      logAcceptanceProb = lf.tryStepping(params, stepSize, mass);
    }
    return stepSize;
  }
};

/* =====================================================================================================
 * Mass-matrix tuners  (S/Sampler.scala:42-50, S/MassMatrix.scala:120-181)
 * ===================================================================================================== */
struct MassMatrixTuner {
  virtual ~MassMatrixTuner() {}
  virtual MassMatrix initialize(LeapFrog& lf, int iterations) = 0;
  virtual bool update(const double* sample, MassMatrix& out) = 0; /* Option[MassMatrix] */
};
struct IdentityMassMatrixTuner : MassMatrixTuner { /* S/MassMatrix.scala:120-124 */
  MassMatrix initialize(LeapFrog&, int) override { return MassMatrix(); }
  bool update(const double*, MassMatrix&) override { return false; }
};
struct StaticMassMatrix : MassMatrixTuner { /* S/Sampler.scala:47-50 */
  MassMatrix mass;
  explicit StaticMassMatrix(const MassMatrix& m) : mass(m) {}
  MassMatrix initialize(LeapFrog&, int) override { return mass; }
  bool update(const double*, MassMatrix&) override { return false; }
};
struct WindowedMassMatrixTuner : MassMatrixTuner { /* S/MassMatrix.scala:126-165 */
  int initialWindowSize;
  double windowExpansion;
  int skipFirst, skipLast;
  bool denseKind;
  std::unique_ptr<VarianceEstimator> var;
  std::unique_ptr<CovarianceEstimator> cov;
  int windowSize;
  int i = 0, j = 0, totalIterations = 0;
  WindowedMassMatrixTuner(int iws, double we, int sf, int sl, bool dense_)
      : initialWindowSize(iws), windowExpansion(we), skipFirst(sf), skipLast(sl), denseKind(dense_), windowSize(iws) {}
  MassMatrix initialize(LeapFrog& lf, int iterations) override { /* :139-143 */
    if (denseKind)
      cov.reset(new CovarianceEstimator(lf.nVars));
    else
      var.reset(new VarianceEstimator(lf.nVars));
    totalIterations = iterations;
    return MassMatrix();
  }
  bool update(const double* sample, MassMatrix& out) override { /* :147-164 */
    j += 1;
    if (j < skipFirst || (totalIterations - j) < skipLast) return false;
    i += 1;
    if (denseKind)
      cov->update(sample);
    else
      var->update(sample);
    if (i == windowSize) {
      i = 0;
      windowSize = jd2i(windowSize * windowExpansion);
      if (denseKind) {
        out = DenseMassMatrix(cov->covariance());
        cov->reset();
      } else {
        out = DiagonalMassMatrix(var->variance());
        var->reset();
      }
      return true;
    }
    return false;
  }
};

/* =====================================================================================================
 * Driver  (S/Driver.scala:7-119)
 * ===================================================================================================== */
struct ChainResult {
  MassMatrix mass;
  int error_flags = 0;
};

static std::unique_ptr<Sampler> make_sampler(const rn_config& c) {
  if (c.sampler == RN_SAMPLER_EHMC)
    return std::unique_ptr<Sampler>(new EHMCSampler(c.max_steps, c.min_steps, c.buf_size, c.p_count));
  return std::unique_ptr<Sampler>(new HMCSampler(c.n_steps));
}
static std::unique_ptr<StepSizeTuner> make_step_tuner(const rn_config& c) {
  if (c.step_size_tuner == RN_STEP_STATIC) return std::unique_ptr<StepSizeTuner>(new StaticStepSize(c.static_step_size));
  return std::unique_ptr<StepSizeTuner>(new DualAvgTuner(c.delta));
}
static std::unique_ptr<MassMatrixTuner> make_mass_tuner(const rn_config& c, int n) {
  switch (c.mass_tuner) {
    case RN_MASS_DIAGONAL:
      return std::unique_ptr<MassMatrixTuner>(
          new WindowedMassMatrixTuner(c.initial_window_size, c.window_expansion, c.skip_first, c.skip_last, false));
    case RN_MASS_DENSE:
      return std::unique_ptr<MassMatrixTuner>(
          new WindowedMassMatrixTuner(c.initial_window_size, c.window_expansion, c.skip_first, c.skip_last, true));
    case RN_MASS_STATIC: {
      MassMatrix m;
      if (c.static_matrix == RN_MATRIX_DIAGONAL)
        m = DiagonalMassMatrix(std::vector<double>(c.static_matrix_elements, c.static_matrix_elements + n));
      else if (c.static_matrix == RN_MATRIX_DENSE)
        m = DenseMassMatrix(std::vector<double>(c.static_matrix_elements, c.static_matrix_elements + (size_t)n * n));
      return std::unique_ptr<MassMatrixTuner>(new StaticMassMatrix(m));
    }
    default: return std::unique_ptr<MassMatrixTuner>(new IdentityMassMatrixTuner());
  }
}

/* Driver.sample for one chain.  `trace` (optional) receives one record per iteration (warmup and sampling):
 * [logAcceptProb, accepted(0/1), stepSizeUsed, leapfrogStepsThisIteration] -- test instrumentation only. */
static void driver_sample(const Model& model, const rn_config& cfg, RNG& rng, double* samples /*[iterations][n]*/,
                          ChainResult& res, rn_chain_stats* st, double* rings, double* trace) {
  RirDensity density(model);
  auto sampler = make_sampler(cfg);                        /* S/Driver.scala:13 */
  auto stepSizeTuner = make_step_tuner(cfg);               /* :14 */
  auto massMatrixTuner = make_mass_tuner(cfg, model.n());  /* :15 */
  LeapFrog lf(density, cfg.stats_window);                  /* :17 */
  const int n = lf.nVars;

  MassMatrix identity;
  std::vector<double> params = lf.initialize(identity, rng); /* :22 */

  /* ---- warmup, S/Driver.scala:48-90 ---- */
  sampler->initialize(params, lf, rng);                      /* :59 */
  double stepSize = stepSizeTuner->initialize(params, lf);   /* :60 */
  MassMatrix mass = massMatrixTuner->initialize(lf, cfg.warmup_iterations); /* :61 */
  std::vector<double> sample(n);
  int tr = 0;
  for (int i = 0; i < cfg.warmup_iterations; i++) { /* :67-88 */
    int64_t steps0 = lf.stats.leapfrogSteps;
    int acc0 = lf.stats.accepted;
    double used = stepSize;
    double logAcceptProb = sampler->warmup(params, lf, stepSize, mass, rng); /* :68 */
    stepSize = stepSizeTuner->update(logAcceptProb);                          /* :69 */
    lf.variables(params, sample.data());                                      /* :74 */
    MassMatrix m;
    if (massMatrixTuner->update(sample.data(), m)) { /* :75-80 */
      mass = m;
      if (m.invalid) res.error_flags |= 2;
      stepSize = stepSizeTuner->reset();
    }
    if (trace) {
      trace[4 * tr + 0] = logAcceptProb;
      trace[4 * tr + 1] = (double)(lf.stats.accepted - acc0);
      trace[4 * tr + 2] = used;
      trace[4 * tr + 3] = (double)(lf.stats.leapfrogSteps - steps0);
      tr++;
    }
  }
  lf.resetStats(); /* :31 */

  /* ---- sampling, S/Driver.scala:92-119 ---- */
  const double finalStep = stepSizeTuner->stepSize(); /* :37 */
  for (int i = 0; i < cfg.iterations; i++) {
    int64_t steps0 = lf.stats.leapfrogSteps;
    int acc0 = lf.stats.accepted;
    sampler->run(params, lf, finalStep, mass, rng); /* :104 */
    lf.variables(params, samples + (size_t)i * n);  /* :105-107 */
    if (trace) {
      trace[4 * tr + 0] = lf.lastLogAcceptanceProb;
      trace[4 * tr + 1] = (double)(lf.stats.accepted - acc0);
      trace[4 * tr + 2] = finalStep;
      trace[4 * tr + 3] = (double)(lf.stats.leapfrogSteps - steps0);
      tr++;
    }
  }

  res.mass = mass;
  if (density.lookup_error) res.error_flags |= 1;
  if (st) {
    const Stats& s = lf.stats;
    std::memset(st, 0, sizeof(*st));
    st->gradient_evaluations = s.gradientEvaluations;
    st->leapfrog_steps = s.leapfrogSteps;
    st->iterations = s.iterations;
    st->divergences = s.divergences;
    st->accepted = s.accepted;
    st->error_flags = res.error_flags;
    st->step_size = finalStep;
    st->energy_mean = s.energyVariance.mean[0];
    st->energy_raw = s.energyVariance.raw[0];
    st->energy_transitions2 = s.energyTransitions2;
    st->energy_samples = s.energyVariance.samples;
    st->ring_pos[0] = s.stepSizes.i;
    st->ring_pos[1] = s.acceptanceRates.i;
    st->ring_pos[2] = s.gradsPerIteration.i;
    st->ring_full[0] = s.stepSizes.full ? 1 : 0;
    st->ring_full[1] = s.acceptanceRates.full ? 1 : 0;
    st->ring_full[2] = s.gradsPerIteration.full ? 1 : 0;
    st->step_sizes_mean = s.stepSizes.mean();
    st->acceptance_rates_mean = s.acceptanceRates.mean();
    st->grads_per_iteration_mean = s.gradsPerIteration.mean();
    st->rng.seed48 = rng.rand.seed;
    st->rng.next_gaussian = rng.rand.next_next_gaussian;
    st->rng.have_next = rng.rand.have_next_next_gaussian ? 1 : 0;
  }
  if (rings) {
    const int w = cfg.stats_window;
    std::memcpy(rings, lf.stats.stepSizes.buf.data(), sizeof(double) * w);
    std::memcpy(rings + w, lf.stats.acceptanceRates.buf.data(), sizeof(double) * w);
    std::memcpy(rings + 2 * w, lf.stats.gradsPerIteration.buf.data(), sizeof(double) * w);
  }
}

} /* namespace rno */

/* =====================================================================================================
 * C ABI (rno_ prefix; same shapes as include/rainier_cuda.h)
 * ===================================================================================================== */
using namespace rno;

struct rno_model {
  Model m;
};

extern "C" {

const char* rno_last_error(void) { return g_last_error.c_str(); }
void rno_set_threads(int n) { g_threads = n; }
int rno_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

void rno_config_default(rn_config* c) { /* S/Sampler.scala:17-27 */
  std::memset(c, 0, sizeof(*c));
  c->struct_size = (int32_t)sizeof(*c);
  c->iterations = 1000;
  c->warmup_iterations = 1000;
  c->stats_window = 100;
  c->sampler = RN_SAMPLER_EHMC;
  c->n_steps = 1;
  c->max_steps = 1024;
  c->min_steps = 1;
  c->buf_size = 100;
  c->p_count = 0.1;
  c->step_size_tuner = RN_STEP_DUAL_AVG;
  c->delta = 0.8;
  c->static_step_size = 1.0;
  c->mass_tuner = RN_MASS_DIAGONAL;
  c->initial_window_size = 50;
  c->window_expansion = 1.5;
  c->skip_first = 50;
  c->skip_last = 50;
}

int rno_model_create(const void* rir, size_t len, const double* const* cols, const int64_t* col_rows, int n_cols,
                     int /*device*/, rno_model** out) {
  std::unique_ptr<rno_model> mm(new rno_model());
  int rc = parse_rir(rir, len, mm->m);
  if (rc) return rc;
  Model& m = mm->m;
  if ((int)(m.h.n_inputs - m.h.n_params) != n_cols) return fail(RN_E_INVALID, "column count does not match RIR inputs");
  m.cols.resize(n_cols);
  for (int i = 0; i < n_cols; i++) m.cols[i].assign(cols[i], cols[i] + col_rows[i]);
  for (auto& T : m.targets)
    for (uint32_t j = 0; j < T.n_cols; j++)
      if (m.cols[T.first_input - m.h.n_params + j].size() != T.n_rows)
        return fail(RN_E_INVALID, "column length does not match target row count");
  rc = classify(m);
  if (rc) return rc;
  *out = mm.release();
  return RN_OK;
}
int rno_model_nvars(const rno_model* m) { return m->m.n(); }
void rno_model_destroy(rno_model* m) { delete m; }

int rno_density_batch(rno_model* mm, const double* q, int chains, double* out) {
  const Model& m = mm->m;
  if (!(m.h.flags & RIR_FLAG_GRADIENT)) return fail(RN_E_UNSUPPORTED, "oracle needs a gradient-carrying RIR");
  RirDensity d(m);
  const int n = m.n();
  for (int c = 0; c < chains; c++) {
    d.update(q + (size_t)c * n);
    out[(size_t)c * (n + 1)] = d.density();
    for (int i = 0; i < n; i++) out[(size_t)c * (n + 1) + 1 + i] = d.gradient(i);
  }
  return d.lookup_error ? fail(RN_E_LOOKUP, "lookup index out of range") : RN_OK;
}

/* trace: optional [chains][warmup+iterations][4] test instrumentation, see driver_sample */
int rno_sample_traced(rno_model* mm, const rn_config* cfg, const int64_t* seeds, int chains, double* samples,
                      double* mass, rn_chain_stats* stats, double* trace) {
  const Model& m = mm->m;
  if (!(m.h.flags & RIR_FLAG_GRADIENT)) return fail(RN_E_UNSUPPORTED, "oracle needs a gradient-carrying RIR");
  const int n = m.n();
  const size_t mass_stride = (cfg->mass_tuner == RN_MASS_DENSE || (cfg->mass_tuner == RN_MASS_STATIC && cfg->static_matrix == RN_MATRIX_DENSE))
                                 ? (size_t)n * n
                                 : (size_t)n;
  std::atomic<int> next(0);
  std::atomic<int> err(0);
  auto work = [&]() {
    for (;;) {
      int c = next.fetch_add(1);
      if (c >= chains) break;
      RNG rng;
      if (cfg->rng_states) {
        rng.rand.seed = cfg->rng_states[c].seed48;
        rng.rand.next_next_gaussian = cfg->rng_states[c].next_gaussian;
        rng.rand.have_next_next_gaussian = cfg->rng_states[c].have_next != 0;
      } else {
        rng.rand = JRandom(seeds[c]); /* ScalaRNG(seed), S/RNG.scala:20-26 */
      }
      ChainResult res;
      const size_t its = (size_t)cfg->iterations;
      driver_sample(m, *cfg, rng, samples + (size_t)c * its * n, res, stats ? stats + c : nullptr,
                    cfg->stats_rings ? cfg->stats_rings + (size_t)c * 3 * cfg->stats_window : nullptr,
                    trace ? trace + (size_t)c * (cfg->warmup_iterations + its) * 4 : nullptr);
      if (mass) {
        double* mo = mass + (size_t)c * mass_stride;
        if (res.mass.kind == RN_MATRIX_IDENTITY) {
          if (mass_stride == (size_t)n)
            for (int i = 0; i < n; i++) mo[i] = 1.0;
          else
            for (int i = 0; i < n; i++)
              for (int j = 0; j < n; j++) mo[i * n + j] = (i == j) ? 1.0 : 0.0;
        } else {
          std::memcpy(mo, res.mass.elements.data(), sizeof(double) * res.mass.elements.size());
        }
      }
      if (res.error_flags) err.fetch_or(res.error_flags);
    }
  };
  int nt = g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, chains));
  if (nt == 1) {
    work();
  } else {
    std::vector<std::thread> ts;
    for (int t = 0; t < nt; t++) ts.emplace_back(work);
    for (auto& t : ts) t.join();
  }
  if (err.load() & 1) return fail(RN_E_LOOKUP, "lookup index out of range");
  if (err.load() & 2) return fail(RN_E_INVALID, "requirement failed: adapted mass matrix contains 0.0 (MassMatrix.scala:8,16)");
  return RN_OK;
}

int rno_sample(rno_model* mm, const rn_config* cfg, const int64_t* seeds, int chains, double* samples, double* mass,
               rn_chain_stats* stats) {
  return rno_sample_traced(mm, cfg, seeds, chains, samples, mass, stats, nullptr);
}

/* ---- compiled functions: Compiler.compile(inputs, outputs) (C/Compiler.scala:22-30) evaluated the way
 * Generator.prepare does per posterior draw (K/Generator.scala:76-93): reqValues(i) = CompiledFunction.output(cf, array,
 * globalBuf, i) for i = 0 until numOutputs, in that order (IR/CompiledFunction.scala:122-140).  The flat SSA array is
 * evaluated in definition order, which is the order the generated output methods define their VarDefs in. ---- */
struct rno_function {
  Model m;
  std::vector<int32_t> needed; /* nodes some output reaches, ascending */
};

int rno_function_create(const void* rir, size_t len, int /*device*/, int /*math_mode*/, rno_function** out) {
  std::unique_ptr<rno_function> f(new rno_function());
  int rc = parse_rir(rir, len, f->m);
  if (rc) return rc;
  const Model& m = f->m;
  if (!(m.h.flags & RIR_FLAG_FUNCTION)) return fail(RN_E_INVALID, "RIR: not a function container");
  if (m.h.n_inputs != m.h.n_params || m.targets.size() != 1 || m.targets[0].n_rows != 0 || m.targets[0].n_cols != 0 ||
      m.targets[0].outputs.empty())
    return fail(RN_E_INVALID, "RIR: malformed function container");
  std::vector<char> need(m.h.n_nodes, 0);
  for (uint32_t o : m.targets[0].outputs) need[o] = 1;
  for (int i = (int)m.h.n_nodes - 1; i >= 0; i--) {
    if (!need[i]) continue;
    const rir_node& nd = m.nodes[i];
    switch (nd.kind) {
      case RIR_UNARY: need[nd.a] = 1; break;
      case RIR_BINARY: need[nd.a] = need[nd.b] = 1; break;
      case RIR_LOOKUP:
        need[nd.a] = 1;
        for (int k = 0; k < nd.c; k++) need[m.lookup_refs[nd.b + k]] = 1;
        break;
      default: break;
    }
  }
  for (uint32_t i = 0; i < m.h.n_nodes; i++)
    if (need[i]) f->needed.push_back((int32_t)i);
  *out = f.release();
  return RN_OK;
}
int rno_function_ninputs(const rno_function* f) { return (int)f->m.h.n_params; }
int rno_function_noutputs(const rno_function* f) { return (int)f->m.targets[0].outputs.size(); }
void rno_function_destroy(rno_function* f) { delete f; }

/* x: [count][n] -> out: [count][m] */
int rno_function_eval(rno_function* f, const double* x, int64_t count, double* out) {
  const Model& m = f->m;
  RirDensity d(m); /* reused for its node interpreter (op semantics of IR/MethodGenerator.scala) */
  const size_t n = m.h.n_params, mo = m.targets[0].outputs.size();
  for (int64_t p = 0; p < count; p++) {
    if (n) std::memcpy(d.inputs.data(), x + (size_t)p * n, n * sizeof(double));
    for (int32_t i : f->needed) d.eval_node(i);
    for (size_t j = 0; j < mo; j++) out[(size_t)p * mo + j] = d.vals[m.targets[0].outputs[j]];
  }
  return d.lookup_error ? fail(RN_E_LOOKUP, "lookup index out of range") : RN_OK;
}

/* ---- java.util.Random exposed for test/bench data synthesis (so a Scala harness can reproduce inputs) ---- */
void rno_jr_seed(rn_rng_state* st, int64_t seed) {
  JRandom r(seed);
  st->seed48 = r.seed;
  st->next_gaussian = 0.0;
  st->have_next = 0;
  st->reserved = 0;
}
static JRandom load(const rn_rng_state* st) {
  JRandom r;
  r.seed = st->seed48;
  r.next_next_gaussian = st->next_gaussian;
  r.have_next_next_gaussian = st->have_next != 0;
  return r;
}
static void store(rn_rng_state* st, const JRandom& r) {
  st->seed48 = r.seed;
  st->next_gaussian = r.next_next_gaussian;
  st->have_next = r.have_next_next_gaussian ? 1 : 0;
}
void rno_jr_doubles(rn_rng_state* st, double* out, int64_t n) {
  JRandom r = load(st);
  for (int64_t i = 0; i < n; i++) out[i] = r.next_double();
  store(st, r);
}
void rno_jr_gaussians(rn_rng_state* st, double* out, int64_t n) {
  JRandom r = load(st);
  for (int64_t i = 0; i < n; i++) out[i] = r.next_gaussian();
  store(st, r);
}
double rno_strict_log(double x) { return strict_log(x); }
double rno_jpow(double x, double y) { return jpow(x, y); }
double rno_strict_exp(double x) { return strict_exp(x); }
void rno_vec_math(int which, const double* x, const double* y, double* out, int64_t n) {
  for (int64_t i = 0; i < n; i++) out[i] = which == 0 ? strict_exp(x[i]) : (which == 1 ? strict_log(x[i]) : strict_pow(x[i], y[i]));
}

} /* extern "C" */
