// rn_emit.hpp -- CUDA source emitter: Program -> one self-contained translation unit (prelude + emitted
// rn_density() + hand-written sampler kernels).  Replaces the reference's JVM bytecode emitter
// (rainier-compute/.../ir/CompiledFunction.scala:42-120 and the *Generator classes).
#pragma once
#include <string>

#include "rn_graph.hpp"

namespace rn {

struct EmitOptions {
  int backend = 0;        // 0 = thread per chain, 1 = warp per chain (rows across lanes)
  bool fast_math = false; // strength-reduce constant powers beyond what Math.pow itself special-cases
  int mass_max = 0;       // 0 identity only, 1 + diagonal, 2 + dense
  bool enable_ehmc = false;
  std::vector<uint64_t> col_offsets;  // element offset of every column placeholder inside the data buffer
};

// the generated rn_density() only
std::string emit_density(const Program& P, const EmitOptions& opt);
// full translation unit; *wpc_smem_doubles receives the per-warp shared-memory footprint (backend 1)
std::string emit_source(const Program& P, const EmitOptions& opt, int* wpc_smem_doubles = nullptr);

}  // namespace rn
