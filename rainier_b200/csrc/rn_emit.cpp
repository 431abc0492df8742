// rn_emit.cpp -- see rn_emit.hpp.
#include "rn_emit.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <sstream>

namespace rn {

extern const char* kPreludeSource;  // rn_prelude.cuh, embedded at build time
extern const char* kSamplerSource;     // rn_args.h + rn_sampler.cuh
extern const char* kSamplerWpcSource;  // rn_args.h + rn_sampler_wpc.cuh

namespace {

std::string lit(double v) {
  if (std::isnan(v)) return "RN_NAN";
  if (std::isinf(v)) return v > 0 ? "RN_INF" : "(-RN_INF)";
  char buf[64];
  std::snprintf(buf, sizeof(buf), "%a", v);  // exact hex-float literal
  std::string s(buf);
  if (v < 0 || (v == 0 && std::signbit(v))) return "(" + s + ")";
  return s;
}

struct Emitter {
  const Program& P;
  const EmitOptions& opt;
  std::ostringstream os;
  bool wpc = false;
  std::vector<int> smem_slot;             // slot -> index in the shared accumulator block, or -1 (register)
  std::map<int, int> tab_off;             // large-lookup node id -> offset of its table in scratch
  int n_smem_acc = 0, tab_doubles = 0;
  Emitter(const Program& p, const EmitOptions& o) : P(p), opt(o) {}

  std::string acc_ref(int slot) const {
    if (!wpc) return "acc[" + std::to_string(slot) + "]";
    if (smem_slot[slot] >= 0) return "scr[" + std::to_string(tab_doubles + smem_slot[slot]) + "]";
    return "a" + std::to_string(slot);
  }
  bool is_big_table(const Node& n) const {
    if (n.kind != K_LOOKUP || n.c <= 8) return false;
    for (int k = 0; k < n.c; k++)
      if (P.nodes[P.lookup_refs[n.b + k]].region != R_INV_FWD && P.nodes[P.lookup_refs[n.b + k]].region != R_INV_BWD) return false;
    return true;
  }

  std::string val(int id) const {
    const Node& n = P.nodes[id];
    if (n.kind == K_CONST) return lit(n.value);
    if (n.kind == K_INPUT) {
      if ((uint32_t)n.a < P.n_params) return "q[" + std::to_string(n.a) + "]";
      return "c" + std::to_string(n.a - (int)P.n_params);
    }
    return "v" + std::to_string(id);
  }

  std::string pow_expr(int a, int b) const {
    const Node& e = P.nodes[b];
    const std::string x = val(a);
    if (e.kind == K_CONST) {
      const double c = e.value;
      if (c == 1.0) return x;
      if (c == 2.0) return "(" + x + " * " + x + ")";
      if (c == -1.0) return "(1.0 / " + x + ")";
      if (opt.fast_math) {
        if (c == 3.0) return "(" + x + " * " + x + " * " + x + ")";
        if (c == 4.0) return "((" + x + " * " + x + ") * (" + x + " * " + x + "))";
        if (c == -2.0) return "(1.0 / (" + x + " * " + x + "))";
        if (c == 0.5) return "sqrt(" + x + ")";
        if (c == -0.5) return "rsqrt(" + x + ")";
        if (c == 1.5) return "(" + x + " * sqrt(" + x + "))";
      }
    }
    return "rn_pow(" + x + ", " + val(b) + ")";
  }

  void stmt(int id, const char* indent) {
    const Node& n = P.nodes[id];
    if (n.kind == K_CONST || n.kind == K_INPUT) return;
    os << indent << "const double v" << id << " = ";
    switch (n.kind) {
      case K_UNARY: {
        const std::string x = val(n.a);
        switch (n.op) {
          case RIR_U_EXP: os << "rn_exp(" << x << ")"; break;
          case RIR_U_LOG: os << "rn_log(" << x << ")"; break;
          case RIR_U_ABS: os << "fabs(" << x << ")"; break;
          case RIR_U_NOOP: os << x; break;
          case RIR_U_SIN: os << "sin(" << x << ")"; break;
          case RIR_U_COS: os << "cos(" << x << ")"; break;
          case RIR_U_TAN: os << "tan(" << x << ")"; break;
          case RIR_U_ASIN: os << "asin(" << x << ")"; break;
          case RIR_U_ACOS: os << "acos(" << x << ")"; break;
          case RIR_U_ATAN: os << "atan(" << x << ")"; break;
          case U_NEG: os << "(-" << x << ")"; break;
          case U_RECIP: os << "(1.0 / " << x << ")"; break;
          case U_SQRT: os << "sqrt(" << x << ")"; break;
        }
        break;
      }
      case K_BINARY: {
        const std::string x = val(n.a), y = val(n.b);
        switch (n.op) {
          case RIR_B_ADD: os << "(" << x << " + " << y << ")"; break;
          case RIR_B_MUL: os << "(" << x << " * " << y << ")"; break;
          case RIR_B_SUB: os << "(" << x << " - " << y << ")"; break;
          case RIR_B_DIV: os << "(" << x << " / " << y << ")"; break;
          case RIR_B_POW: os << pow_expr(n.a, n.b); break;
          case RIR_B_COMPARE: os << "rn_compare(" << x << ", " << y << ")"; break;
        }
        break;
      }
      case K_LOOKUP: {
        // D2I ; tableswitch ; default -> throw (ir/ExprMethodGenerator.scala:50-56): flag + NaN instead of a fault
        if (wpc && tab_off.count(id)) {
          os << "rn_tab_lookup(scr + " << tab_off.at(id) << ", " << n.c << ", " << n.d << ", " << val(n.a) << ", err)";
          break;
        }
        os << "rn_lookup" << id << "(" << val(n.a);
        for (int k = 0; k < n.c; k++) os << ", " << val(P.lookup_refs[n.b + k]);
        os << ", err)";
        break;
      }
      case K_SELEQ: os << "((rn_d2i(" << val(n.a) << ") == " << n.d << ") ? " << val(n.b) << " : " << val(n.c) << ")"; break;
      case K_ACC: os << acc_ref(n.a); break;
    }
    os << ";\n";
  }

  // small lookups become a helper with a switch (keeps operands in registers)
  void lookup_helpers() {
    for (size_t id = 0; id < P.nodes.size(); id++) {
      const Node& n = P.nodes[id];
      if (n.kind != K_LOOKUP) continue;
      if (wpc && tab_off.count((int)id)) continue;
      os << "RN_DEVICE double rn_lookup" << id << "(double idx";
      for (int k = 0; k < n.c; k++) os << ", double e" << k;
      os << ", int& err) {\n  switch (rn_d2i(idx) - (" << n.d << ")) {\n";
      for (int k = 0; k < n.c; k++) os << "    case " << k << ": return e" << k << ";\n";
      os << "    default: err |= 1; return RN_NAN;\n  }\n}\n";
    }
  }

  void density_tpc() {
    os << "// ---- emitted: log-density and gradient of the frozen DAG (" << (P.symbolic ? "symbolic" : "adjoint")
       << " gradient) ----\n";
    lookup_helpers();
    os << "RN_DEVICE void rn_density(const double (&q)[RN_N], double& dens, double (&grad)[RN_N], "
          "const double* RN_RESTRICT data, int& err) {\n";
    os << "  (void)data; (void)err;\n";
    os << "  double acc[RN_NSLOTS];\n  for (int s = 0; s < RN_NSLOTS; s++) acc[s] = 0.0;\n";
    for (int id : P.inv_fwd) stmt(id, "  ");
    for (size_t t = 0; t < P.targets.size(); t++) {
      const TargetInfo& T = P.targets[t];
      os << "  // target " << t << (T.streamed() ? " (streamed)" : " (data-free)") << "\n";
      const char* ind = "  ";
      if (T.streamed()) {
        os << "  for (long long row = 0; row < " << (long long)T.n_rows << "LL; row++) {\n";
        ind = "    ";
        std::set<int> used;
        for (int id : T.row_fwd)
          if (P.nodes[id].kind == K_INPUT) used.insert(P.nodes[id].a - (int)P.n_params);
        for (int k : used)
          os << ind << "const double c" << k << " = RN_LDG(data + " << (unsigned long long)opt.col_offsets[k]
             << "ULL + row);\n";
        for (int id : T.row_fwd) stmt(id, ind);
        for (int id : T.row_bwd) stmt(id, ind);
      }
      for (const AccStmt& a : T.row_acc) os << ind << "acc[" << a.slot << "] += " << val(a.node) << ";\n";
      for (const ScatterStmt& sc : T.row_scatter) {
        os << ind << "{ const int k = rn_d2i(" << val(sc.index_node) << ") - (" << sc.low << "); if (k < 0 || k >= "
           << sc.len << ") err |= 1; else acc[" << sc.slot_base << " + k] += " << val(sc.node) << "; }\n";
      }
      if (T.streamed()) os << "  }\n";
    }
    os << "  dens = acc[0];\n";
    if (P.symbolic) {
      os << "  RN_UNROLL\n  for (int i = 0; i < RN_N; i++) grad[i] = acc[1 + i];\n";
    } else {
      for (int id : P.inv_bwd) stmt(id, "  ");
      for (uint32_t i = 0; i < P.n_params; i++) os << "  grad[" << i << "] = " << val(P.grad_nodes[i]) << ";\n";
    }
    os << "}\n";
  }

  void density_wpc() {
    wpc = true;
    // which accumulator slots live in shared memory (targets of a scatter) and which in registers
    smem_slot.assign(P.n_slots, -1);
    for (const TargetInfo& T : P.targets)
      for (const ScatterStmt& sc : T.row_scatter)
        for (int k = 0; k < sc.len; k++)
          if (smem_slot[sc.slot_base + k] < 0) smem_slot[sc.slot_base + k] = n_smem_acc++;
    for (const TargetInfo& T : P.targets)
      for (int id : T.row_fwd)
        if (is_big_table(P.nodes[id])) {
          tab_off[id] = tab_doubles;
          tab_doubles += P.nodes[id].c;
        }
    os << "// ---- emitted: log-density and gradient of the frozen DAG (" << (P.symbolic ? "symbolic" : "adjoint")
       << " gradient), warp-per-chain: rows across lanes ----\n";
    os << "#define RN_WPC_SCRATCH " << (tab_doubles + n_smem_acc) << "\n";
    os << "RN_DEVICE double rn_tab_lookup(const double* tab, int len, int low, double idx, int& err) {\n"
          "  const int k = rn_d2i(idx) - low;\n  if (k < 0 || k >= len) { err |= 1; return RN_NAN; }\n  return tab[k];\n}\n";
    os << "RN_DEVICE double rn_warp_sum(double x) {\n  RN_UNROLL\n  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);\n  return x;\n}\n";
    lookup_helpers();
    os << "RN_DEVICE void rn_density(const double* q, double& dens, double* grad, double* scr, "
          "const double* RN_RESTRICT data, int& err) {\n";
    os << "  (void)data; (void)scr;\n  const int lane = (int)(threadIdx.x & 31);\n  (void)lane;\n";
    if (n_smem_acc) os << "  for (int k = lane; k < " << n_smem_acc << "; k += 32) scr[" << tab_doubles << " + k] = 0.0;\n";
    for (int id : P.inv_fwd) stmt(id, "  ");
    for (auto& kv : tab_off) {
      const Node& n = P.nodes[kv.first];
      for (int k = 0; k < n.c; k++) os << "  scr[" << (kv.second + k) << "] = " << val(P.lookup_refs[n.b + k]) << ";\n";
    }
    os << "  __syncwarp();\n";
    for (int sl = 0; sl < P.n_slots; sl++)
      if (smem_slot[sl] < 0) os << "  double a" << sl << " = 0.0;\n";
    for (size_t t = 0; t < P.targets.size(); t++) {
      const TargetInfo& T = P.targets[t];
      os << "  // target " << t << (T.streamed() ? " (streamed, rows across lanes)" : " (data-free)") << "\n";
      if (T.streamed()) {
        os << "  for (long long row = lane; row < " << (long long)T.n_rows << "LL; row += 32) {\n";
        std::set<int> used;
        for (int id : T.row_fwd)
          if (P.nodes[id].kind == K_INPUT) used.insert(P.nodes[id].a - (int)P.n_params);
        for (int k : used)
          os << "    const double c" << k << " = RN_LDG(data + " << (unsigned long long)opt.col_offsets[k] << "ULL + row);\n";
        for (int id : T.row_fwd) stmt(id, "    ");
        for (int id : T.row_bwd) stmt(id, "    ");
        for (const AccStmt& a : T.row_acc) os << "    " << acc_ref(a.slot) << " += " << val(a.node) << ";\n";
        for (const ScatterStmt& sc : T.row_scatter) {
          os << "    { const int k = rn_d2i(" << val(sc.index_node) << ") - (" << sc.low << "); if (k < 0 || k >= " << sc.len
             << ") err |= 1; else atomicAdd(&scr[" << (tab_doubles + smem_slot[sc.slot_base]) << " + k], " << val(sc.node) << "); }\n";
        }
        os << "  }\n";
      } else {
        os << "  if (lane == 0) {\n";  // counted once by the butterfly below
        for (const AccStmt& a : T.row_acc) os << "    " << acc_ref(a.slot) << " += " << val(a.node) << ";\n";
        os << "  }\n";
      }
    }
    for (int sl = 0; sl < P.n_slots; sl++)
      if (smem_slot[sl] < 0) os << "  a" << sl << " = rn_warp_sum(a" << sl << ");\n";
    os << "  err = (int)__reduce_or_sync(0xffffffffu, (unsigned)err);\n  __syncwarp();\n";
    os << "  dens = a0;\n";
    if (P.symbolic) {
      for (uint32_t i = 0; i < P.n_params; i++) os << "  if (lane == 0) grad[" << i << "] = " << acc_ref(1 + (int)i) << ";\n";
    } else {
      for (int id : P.inv_bwd) stmt(id, "  ");
      for (uint32_t i = 0; i < P.n_params; i++) os << "  if (lane == 0) grad[" << i << "] = " << val(P.grad_nodes[i]) << ";\n";
    }
    os << "  __syncwarp();\n}\n";
  }
};

}  // namespace

std::string emit_density(const Program& P, const EmitOptions& opt) {
  Emitter E(P, opt);
  if (opt.backend == 1)
    E.density_wpc();
  else
    E.density_tpc();
  return E.os.str();
}

std::string emit_source(const Program& P, const EmitOptions& opt, int* wpc_smem_doubles) {
  std::ostringstream os;
  os << "// generated by rainier_b200 (CUDA source emitter) -- do not edit\n";
  os << "#define RN_N " << P.n_params << "\n";
  os << "#define RN_NSLOTS " << P.n_slots << "\n";
  os << "#define RN_BACKEND " << opt.backend << "\n";
  os << "#define RN_MASS_MAX " << opt.mass_max << "\n";
  os << "#define RN_ENABLE_EHMC " << (opt.enable_ehmc ? 1 : 0) << "\n";
  if (opt.fast_math) os << "#define RN_FAST_MATH 1\n";
  os << kPreludeSource << "\n";
  os << emit_density(P, opt) << "\n";
  if (opt.backend == 1) {
    os << "#define RN_WPC_SMEM_DOUBLES (" << (opt.enable_ehmc ? 7 : 4) << " * RN_N + RN_WPC_SCRATCH)\n";
    os << kSamplerWpcSource << "\n";
    if (wpc_smem_doubles) {
      Emitter E(P, opt);
      E.density_wpc();
      *wpc_smem_doubles = (opt.enable_ehmc ? 7 : 4) * (int)P.n_params + E.tab_doubles + E.n_smem_acc;
    }
  } else {
    os << kSamplerSource << "\n";
  }
  return os.str();
}

}  // namespace rn
