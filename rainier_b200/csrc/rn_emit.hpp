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
  int tma_stages = 0;     // warp per chain: shared-memory stages of the CTA-shared data-tile pipeline (0 = off)
  int wpc_k = 1;          // warp per chain: warps owning one chain (1, 2, 4 or 8; > 1 for chains with a large state)
  std::vector<uint64_t> target_base;  // per target: element offset of its tile-major [tile][column][32] block in the data buffer
};

// the generated rn_density() only
std::string emit_density(const Program& P, const EmitOptions& opt);
// full translation unit
std::string emit_source(const Program& P, const EmitOptions& opt);
// function flavour (Program from build_function): prelude + emitted rn_function() + rn_function.cuh (rn_k_eval); only
// opt.fast_math is read
std::string emit_function_source(const Program& P, const EmitOptions& opt);
// optimizer flavour: prelude + emitted thread-per-chain rn_density() + rn_optimizer.cuh (rn_k_lbfgs, `history` = the m of
// new LBFGS(x, m, eps)); reads opt.fast_math and opt.target_base
std::string emit_optimizer_source(const Program& P, const EmitOptions& opt, int history);
// shared-memory needs of the warp-per-chain kernels: doubles per warp (chain vectors + density scratch) and doubles of
// the largest data tile (n_cols * 32 over the streamed targets; 0 when nothing is streamed)
struct WpcSizes {
  int per_warp_doubles = 0;  // per CHAIN (its wpc_k warps share the slice)
  int tile_doubles = 0;
};
WpcSizes wpc_sizes(const Program& P, const EmitOptions& opt);

}  // namespace rn
