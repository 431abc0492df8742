// rn_inline.hpp -- device-side inlining of separable likelihoods.
//
// The reference folds an "inlinable" likelihood into data-free constants on the JVM before compiling: TargetGroup.inlinable
// decides (rainier-compute/.../compute/Target.scala:136-207), PartialEvaluator.inline folds row by row
// (compute/PartialEvaluator.scala:86-97) -- O(rows x DAG size) of immutable-map churn at model-build time, and only for targets
// its simplifier happens to have expanded (it stops expanding squares at 5 additive terms, compute/LogLineOps.scala:43-66, so a
// Gaussian regression with >= 4 covariates streams its data on every gradient).  Here the STREAMED container is accepted as
// it is and rewritten at rn_model_create:
//   1. plan_inline: every row value of a streamed target is carried as  sum_m  (product of <= 2 column-only values) * coef_m
//      with parameter-only coefficients (ADD / SUB merge, MUL and x^2 multiply out, division by a one-sided value scales);
//      any other operation on a value that mixes parameters and columns makes the target non-separable (left streamed).
//   2. the sums  S_m = sum over rows of the column-only products  are computed ON THE DEVICE: a function-flavour program
//      over the target's columns (one thread per row, the tile-major data in place, rn_k_eval) and a fixed-order row
//      reduction (rn_k_reduce_rows) -- HBM-bound, once per model.
//   3. apply_inline: the target's output becomes  n_rows * coef_() + sum_m S_m * coef_m , a data-free expression; the
//      emitter's reverse mode differentiates it like any other.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rainier_rir.h"

namespace rn {

struct InlineTarget {
  int target = -1;
  std::vector<std::vector<int32_t>> monos;  // per monomial: ids of column-only nodes (sorted, 1..max_degree of them)
  std::vector<int32_t> coef;                // per monomial: id of its parameter-only coefficient node (extended node array)
  int32_t const_coef = -1;                  // coefficient of the empty monomial (a per-row constant of the chain), or -1
};

struct InlinePlan {
  rir_header h;
  std::vector<rir_node> nodes;  // the container's nodes, then appended ones (coefficients, reciprocals of column values)
  std::vector<int32_t> lookup_refs;
  struct RawTarget {
    rir_target t;
    std::vector<uint32_t> outputs;
  };
  std::vector<RawTarget> targets;
  std::vector<InlineTarget> inl;  // separable streamed targets (possibly none)
};

// returns "" on success (plan.inl may be empty: nothing to inline); only primal containers (no RIR_FLAG_GRADIENT) are planned
std::string plan_inline(const void* rir, size_t len, InlinePlan& plan, int max_degree = 2, int max_monomials = 4096);
// RIR_FLAG_FUNCTION container: inputs = the columns of plan.inl[k]'s target (in column order), outputs = its monomials
std::vector<uint8_t> inline_function_rir(const InlinePlan& plan, size_t k);
// the rewritten container; sums[k][m] = sum over the rows of target plan.inl[k] of monomial m
std::vector<uint8_t> apply_inline(const InlinePlan& plan, const std::vector<std::vector<double>>& sums);

}  // namespace rn
