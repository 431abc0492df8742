/*
 * rainier_rir.h -- "RIR": the frozen-DAG wire format handed across the drop-in boundary.
 *
 * The reference never serializes its compute DAG: `Compiler.compile` hands an in-memory
 * `Seq[ir.Param]` + `Seq[(String, ir.Expr)]` to the JVM-bytecode emitter
 * (rainier-compute/.../compute/Compiler.scala:22-30, ir/CompiledFunction.scala:42-120) and wraps it
 * in `ir.DataFunction(cf, numParamInputs, numOutputs, data)` (compute/Compiler.scala:14-20,
 * ir/DataFunction.scala:13-30).  RIR is a 1:1 flat encoding of exactly that hand-off:
 *
 *   - the closed sum type of ir/IR.scala:3-23 (Param / Const / VarDef+VarRef / BinaryIR / UnaryIR /
 *     LookupIR / SeqIR) flattened into an SSA node array in definition order.  A `VarDef(sym, rhs)`
 *     becomes the node that computes `rhs`; every `VarRef(sym)` becomes that node's index; `SeqIR`
 *     (evaluate-first-then-second, ir/ExprMethodGenerator.scala:57-63) disappears because node order
 *     already is evaluation order.  The reference guarantees defs precede refs
 *     ("VarRef was used before its VarDef" is a hard error, compute/Translator.scala:178-179), so the
 *     array is topologically sorted by construction: operands always have smaller indices.
 *   - ops of ir/Ops.scala:3-37.
 *   - `DataFunction`'s layout: inputs = nVars parameters then every target's column placeholders
 *     (compute/Target.scala:38-41); per target a row count and a list of output node ids
 *     (compute/Target.scala:50-56).
 *
 * Two flavours travel in the same container:
 *   RIR_FLAG_GRADIENT set   : every target carries nVars+1 outputs [density, d/dq_0 .. d/dq_{n-1}],
 *                             i.e. what `Compiler.compileTargets` produces today (symbolic gradient,
 *                             compute/Gradient.scala:8-69).  The CPU oracle consumes this flavour and the
 *                             CUDA emitter accepts it too ("symbolic" gradient mode).
 *   RIR_FLAG_GRADIENT clear : every target carries 1 output (the primal log-density); the CUDA emitter
 *                             derives adjoints itself (reverse mode over the SSA array).  This is the
 *                             flavour the Scala wrapper sends.
 *
 * All integers little-endian; all structs packed as declared (natural alignment, no padding surprises:
 * every struct size is a multiple of 8).
 */
#ifndef RAINIER_RIR_H
#define RAINIER_RIR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIR_MAGIC 0x31524952u /* "RIR1" */
#define RIR_VERSION 1u

#define RIR_FLAG_GRADIENT 1u
/* RIR_FLAG_FUNCTION: the container is the hand-off of the OTHER compile seam, `Compiler.compile(inputs: Seq[ir.Param],
 * outputs: Seq[(String, Real)]): ir.CompiledFunction` (compute/Compiler.scala:22-30) as Generator.prepare uses it for the
 * "requirements" of a posterior-predictive generator (core/Generator.scala:59-94, called from Trace.predict,
 * core/Trace.scala:34-41): n_inputs == n_params (parameters only, no columns), n_targets == 1 with n_rows == 0,
 * n_cols == 0 and n_outputs == m >= 1 arbitrary output nodes ("req0".."req{m-1}").  No gradient, no accumulation:
 * output j of a point is the value of node outputs[j].  Consumed by rn_function_create (rainier_cuda.h). */
#define RIR_FLAG_FUNCTION 2u

/* node kinds */
enum {
  RIR_INPUT = 0,  /* a = input index: [0,n_params) parameters, then column placeholders  (ir.Param) */
  RIR_CONST = 1,  /* value                                                                (ir.Const) */
  RIR_UNARY = 2,  /* op = RIR_U_*, a = operand node                                       (UnaryIR)  */
  RIR_BINARY = 3, /* op = RIR_B_*, a = left node, b = right node                          (BinaryIR) */
  RIR_LOOKUP = 4  /* a = index node, b = offset into lookup_refs, c = table length, d = low (LookupIR) */
};

/* binary ops, ir/Ops.scala:3-23.  Subtract/Divide exist in the reference's op set but its Translator never
 * emits them (compute/Translator.scala:154-156); they are encoded for completeness. */
enum { RIR_B_ADD = 0, RIR_B_MUL = 1, RIR_B_SUB = 2, RIR_B_DIV = 3, RIR_B_POW = 4, RIR_B_COMPARE = 5 };

/* unary ops, ir/Ops.scala:25-37 */
enum {
  RIR_U_EXP = 0,
  RIR_U_LOG = 1,
  RIR_U_ABS = 2,
  RIR_U_NOOP = 3,
  RIR_U_SIN = 4,
  RIR_U_COS = 5,
  RIR_U_TAN = 6,
  RIR_U_ASIN = 7,
  RIR_U_ACOS = 8,
  RIR_U_ATAN = 9
};

typedef struct rir_header {
  uint32_t magic;         /* RIR_MAGIC */
  uint32_t version;       /* RIR_VERSION */
  uint32_t n_params;      /* nVars: number of sampled parameters (DataFunction.numParamInputs) */
  uint32_t n_inputs;      /* n_params + total number of column placeholders (cf.numInputs) */
  uint32_t n_nodes;
  uint32_t n_targets;     /* prior + one per likelihood (compute/Target.scala:73-79) */
  uint32_t n_lookup_refs; /* total entries in the lookup-ref side table */
  uint32_t flags;         /* RIR_FLAG_* */
} rir_header; /* 32 bytes */

typedef struct rir_node {
  uint8_t kind; /* RIR_INPUT .. RIR_LOOKUP */
  uint8_t op;   /* RIR_B_* or RIR_U_* */
  uint16_t reserved0;
  int32_t a;
  int32_t b;
  int32_t c;
  int32_t d;
  int32_t reserved1;
  double value; /* RIR_CONST only; IEEE-754 bits incl. +-inf */
} rir_node; /* 32 bytes */

typedef struct rir_target {
  uint64_t n_rows;      /* rows streamed per evaluation; 0 = data-free target (evaluated once) */
  uint32_t first_input; /* index of this target's first column placeholder in the input vector */
  uint32_t n_cols;      /* number of column placeholders (columns ++ gradientColumns) */
  uint32_t n_outputs;   /* n_params+1 if RIR_FLAG_GRADIENT else 1 */
  uint32_t reserved;
  /* followed by n_outputs x uint32_t output node ids, padded to a multiple of 8 bytes */
} rir_target; /* 24 bytes + outputs */

/*
 * File layout:
 *   rir_header
 *   rir_node      nodes[n_nodes]
 *   int32_t       lookup_refs[n_lookup_refs]   (padded to a multiple of 8 bytes)
 *   n_targets x { rir_target, uint32_t outputs[n_outputs] (padded to 8) }
 */

#ifdef __cplusplus
}
#endif
#endif
