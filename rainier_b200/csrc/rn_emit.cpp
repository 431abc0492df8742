// rn_emit.cpp -- see rn_emit.hpp.
#include "rn_emit.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <set>
#include <sstream>

namespace rn {

extern const char* kPreludeSource;  // rn_prelude.cuh, embedded at build time
extern const char* kSamplerSource;     // rn_args.h + rn_sampler.cuh
extern const char* kSamplerWpcSource;  // rn_args.h + rn_sampler_wpc.cuh
extern const char* kFunctionSource;    // rn_function.cuh
extern const char* kOptimizerSource;   // rn_args.h + rn_optimizer.cuh

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
  std::vector<int> tab_fill;              // one representative lookup node per distinct table
  int red_off = 0, red_doubles = 0;       // cross-warp reduction scratch (K warps per chain)
  int n_smem_acc = 0, tab_doubles = 0;
  std::string col_suffix;                 // names of column values loaded in the current region of a row body
  Emitter(const Program& p, const EmitOptions& o) : P(p), opt(o) {}

  // The body of one row of a streamed target.  Keeps live ranges short, because a "row" of the reference's
  // Model.observe is 8 unrolled observations over hundreds of columns (core/Model.scala:98-132):
  //   * a column value is loaded right before its first use (not hoisted to the top of the row),
  //   * the reverse sweep re-loads the columns it needs behind a compiler fence instead of keeping the forward
  //     sweep's copies alive (2 loads per element per row; the data sits in L1/L2/shared memory),
  //   * every accumulation / scatter is issued as soon as its operand exists (a_j += w * x_j contracts to one FMA in
  //     fast mode).  The order of the additions into any one slot is unchanged.
  int local_col(const TargetInfo& T, int k) const { return k - ((int)T.first_input - (int)P.n_params); }

  void operands(int id, std::vector<int>& out) const {
    const Node& n = P.nodes[id];
    out.clear();
    switch (n.kind) {
      case K_UNARY: out.push_back(n.a); break;
      case K_BINARY: out.push_back(n.a); out.push_back(n.b); break;
      case K_LOOKUP:
        out.push_back(n.a);
        for (int k = 0; k < n.c; k++) out.push_back(P.lookup_refs[n.b + k]);
        break;
      case K_SELEQ: out.push_back(n.a); out.push_back(n.b); out.push_back(n.c); break;
      default: break;
    }
  }

  // see row_body(): splits the row statements into independent dataflow components and merges them round-robin
  void interleave_components(const TargetInfo& T, const std::set<int>& body, std::vector<int>& order_fwd,
                             std::vector<int>& order_bwd) const {
    const int BIG = 4;  // two components of at least this many statements meeting in one statement = a joiner
    std::map<int, int> parent, size;  // union-find over statement ids
    std::set<int> tail;
    std::function<int(int)> find = [&](int x) {
      while (parent[x] != x) {
        parent[x] = parent[parent[x]];
        x = parent[x];
      }
      return x;
    };
    std::vector<int> ops;
    auto classify = [&](int s) {
      operands(s, ops);
      bool is_tail = false;
      std::set<int> comps;
      for (int o : ops) {
        if (!body.count(o)) continue;
        if (tail.count(o)) {
          is_tail = true;
          break;
        }
        comps.insert(find(o));
      }
      if (!is_tail && comps.size() >= 2) {
        int big = 0;
        for (int c : comps)
          if (size[c] >= BIG) big++;
        if (big >= 2) is_tail = true;
      }
      if (is_tail) {
        tail.insert(s);
        return;
      }
      parent[s] = s;
      size[s] = 1;
      for (int c : comps) {
        const int r = find(s), q = find(c);
        if (r == q) continue;
        parent[q] = r;
        size[r] += size[q];
      }
    };
    std::vector<int> fwd, bwd;
    for (int id : T.row_fwd)
      if (body.count(id)) {
        classify(id);
        fwd.push_back(id);
      }
    for (int id : T.row_bwd)
      if (body.count(id)) {
        classify(id);
        bwd.push_back(id);
      }
    auto schedule = [&](const std::vector<int>& region, std::vector<int>& out) {
      std::vector<int> comp_order;
      std::map<int, std::vector<int>> lists;
      std::vector<int> tails;
      for (int id : region) {
        if (tail.count(id)) {
          tails.push_back(id);
          continue;
        }
        const int c = find(id);
        if (!lists.count(c)) comp_order.push_back(c);
        lists[c].push_back(id);
      }
      std::vector<size_t> pos(comp_order.size(), 0);
      for (bool any = true; any;) {
        any = false;
        for (size_t k = 0; k < comp_order.size(); k++) {
          const std::vector<int>& l = lists[comp_order[k]];
          if (pos[k] < l.size()) {
            out.push_back(l[pos[k]++]);
            any = true;
          }
        }
      }
      out.insert(out.end(), tails.begin(), tails.end());
    };
    schedule(fwd, order_fwd);
    schedule(bwd, order_bwd);
  }

  template <class Load, class AccRef>
  void row_body(const TargetInfo& T, const char* ind, Load load, AccRef accref, bool atomic_scatter, int scatter_base_off) {
    std::set<int> body;
    for (int id : T.row_fwd)
      if (P.nodes[id].kind != K_CONST && P.nodes[id].kind != K_INPUT) body.insert(id);
    for (int id : T.row_bwd)
      if (P.nodes[id].kind != K_CONST && P.nodes[id].kind != K_INPUT) body.insert(id);
    std::map<int, std::vector<const AccStmt*>> acc_at;
    std::map<int, std::vector<const ScatterStmt*>> sc_at;
    std::vector<const AccStmt*> acc_tail;
    std::vector<const ScatterStmt*> sc_tail;
    for (const AccStmt& a : T.row_acc) (body.count(a.node) ? acc_at[a.node] : acc_tail).push_back(&a);
    for (const ScatterStmt& sc : T.row_scatter) (body.count(sc.node) ? sc_at[sc.node] : sc_tail).push_back(&sc);
    std::set<int> declared;
    auto need_col = [&](int o) {
      const Node& n = P.nodes[o];
      if (n.kind != K_INPUT || (uint32_t)n.a < P.n_params) return;
      const int k = n.a - (int)P.n_params;
      if (declared.insert(k).second) os << ind << "const double c" << k << col_suffix << " = " << load(k) << ";\n";
    };
    auto need_operands = [&](int id) {
      const Node& n = P.nodes[id];
      switch (n.kind) {
        case K_UNARY: need_col(n.a); break;
        case K_BINARY: need_col(n.a); need_col(n.b); break;
        case K_LOOKUP:
          need_col(n.a);
          for (int k = 0; k < n.c; k++) need_col(P.lookup_refs[n.b + k]);
          break;
        case K_SELEQ: need_col(n.a); need_col(n.b); need_col(n.c); break;
        default: break;
      }
    };
    auto emit_acc = [&](const AccStmt& a) { os << ind << accref(a.slot) << " += " << val(a.node) << ";\n"; };
    auto emit_scatter = [&](const ScatterStmt& sc) {
      os << ind << "{ const int k = rn_d2i(" << val(sc.index_node) << ") - (" << sc.low << "); if (k < 0 || k >= " << sc.len
         << ") err |= 1; else ";
      if (atomic_scatter)
        os << "atomicAdd(&scr[" << (scatter_base_off + smem_slot[sc.slot_base]) << " + k], " << val(sc.node) << "); }\n";
      else
        os << "acc[" << sc.slot_base << " + k] += " << val(sc.node) << "; }\n";
    };
    auto one = [&](int id) {
      if (!body.count(id)) return;
      need_operands(id);
      stmt(id, ind);
      auto ia = acc_at.find(id);
      if (ia != acc_at.end())
        for (const AccStmt* a : ia->second) emit_acc(*a);
      auto is = sc_at.find(id);
      if (is != sc_at.end())
        for (const ScatterStmt* sc : is->second) {
          need_col(sc->index_node);
          emit_scatter(*sc);
        }
    };
    // Instruction-level parallelism: the unrolled observations of a row are independent dataflow components whose
    // results only meet in a final sum.  Emitting them one after the other leaves each warp with a single serial
    // dependency chain (a 50-term dot product is 50 dependent FMAs; ncu: stall_wait dominates at 8 warps/SM), so the
    // statements of the components are interleaved round-robin; "joiner" statements (and everything downstream of
    // them) follow in their original order.  Values are unchanged (SSA); only the issue order moves.
    std::vector<int> order_fwd, order_bwd;
    interleave_components(T, body, order_fwd, order_bwd);
    col_suffix.clear();
    for (int id : order_fwd) one(id);
    if (!order_bwd.empty()) {
      os << ind << "RN_FENCE();\n";
      declared.clear();
      col_suffix = "b";
      for (int id : order_bwd) one(id);
    }
    for (const AccStmt* a : acc_tail) {
      need_col(a->node);
      emit_acc(*a);
    }
    for (const ScatterStmt* sc : sc_tail) {
      need_col(sc->index_node);
      need_col(sc->node);
      emit_scatter(*sc);
    }
    col_suffix.clear();
  }

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
      return "c" + std::to_string(n.a - (int)P.n_params) + col_suffix;
    }
    return "v" + std::to_string(id);
  }

  // `derived`: the node was created by the emitter's own reverse sweep (there is no reference operation to mirror), so
  // constant powers are strength-reduced in parity mode as well -- d/dx x^-1 = -x^-2 would otherwise cost one fdlibm
  // pow() per observation of a logistic regression
  std::string pow_expr(int a, int b, bool derived, bool row_variant = false) const {
    const Node& e = P.nodes[b];
    const std::string x = val(a);
    if (e.kind == K_CONST) {
      const double c = e.value;
      if (c == 1.0) return x;
      if (c == 2.0) return "(" + x + " * " + x + ")";
      if (c == -1.0) return "(1.0 / " + x + ")";
      if (derived && !opt.fast_math) {
        if (c == 3.0) return "(" + x + " * " + x + " * " + x + ")";
        if (c == 4.0) return "((" + x + " * " + x + ") * (" + x + " * " + x + "))";
        if (c == -2.0) return "(1.0 / (" + x + " * " + x + "))";
        if (c == -3.0) return "(1.0 / (" + x + " * " + x + " * " + x + "))";
        if (c == 0.5) return "sqrt(" + x + ")";
        if (c == -0.5) return "(1.0 / sqrt(" + x + "))";
        if (c == 1.5) return "(" + x + " * sqrt(" + x + "))";
        if (c == -1.5) return "(1.0 / (" + x + " * sqrt(" + x + ")))";
      }
      if (opt.fast_math) {
        if (c == 3.0) return "(" + x + " * " + x + " * " + x + ")";
        if (c == 4.0) return "((" + x + " * " + x + ") * (" + x + " * " + x + "))";
        if (c == -2.0) return "(1.0 / (" + x + " * " + x + "))";
        if (c == 0.5) return "sqrt(" + x + ")";
        if (c == -0.5) return "rsqrt(" + x + ")";
        if (c == 1.5) return "(" + x + " * sqrt(" + x + "))";
      }
    }
    return std::string(row_variant ? "rn_pow_libm(" : "rn_pow(") + x + ", " + val(b) + ")";
  }
  bool row_libm(const Node& n) const { return wpc && (n.region == R_ROW_FWD || n.region == R_ROW_BWD); }

  void stmt(int id, const char* indent) {
    const Node& n = P.nodes[id];
    if (n.kind == K_CONST || n.kind == K_INPUT) return;
    os << indent << "const double v" << id << " = ";
    switch (n.kind) {
      case K_UNARY: {
        const std::string x = val(n.a);
        switch (n.op) {
          // Row-variant transcendentals of the warp-per-chain shape use CUDA's libm (<= 1 ulp, like the JVM's own
          // Math.exp/log intrinsics): rows are summed in tree order there, so those results are not bit-comparable with
          // the oracle anyway (1e-13 agreement), and fdlibm costs twice the instructions.  Everything that stays
          // bit-exact -- invariant parts, data-free targets, the thread-per-chain kernels -- keeps fdlibm.
          case RIR_U_EXP: os << (row_libm(n) && !getenv("RN_ROW_EXP_FDLIBM") ? "exp(" : "rn_exp(") << x << ")"; break;
          case RIR_U_LOG: os << (row_libm(n) ? "log(" : "rn_log(") << x << ")"; break;
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
          case RIR_B_POW: os << pow_expr(n.a, n.b, n.region == R_ROW_BWD || n.region == R_INV_BWD, row_libm(n)); break;
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
      if (T.streamed()) {
        os << "  for (long long row = 0; row < " << (long long)T.n_rows << "LL; row++) {\n";
        os << "    const double* RN_RESTRICT rp = data + " << (unsigned long long)opt.target_base[t] << "ULL + (row >> 5) * "
           << (unsigned long long)T.n_cols * 32 << "LL + (row & 31);\n";
        row_body(
            T, "    ", [&](int k) { return "RN_LDG(rp + " + std::to_string(local_col(T, k) * 32) + ")"; },
            [&](int slot) { return "acc[" + std::to_string(slot) + "]"; }, false, 0);
        os << "  }\n";
      } else {
        for (const AccStmt& a : T.row_acc) os << "  acc[" << a.slot << "] += " << val(a.node) << ";\n";
      }
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

  // Function flavour: forward evaluation of the m outputs of Compiler.compile(inputs, outputs).  An output is stored as
  // soon as its node is defined (500 requirements -- Generator.MaxRequirements -- must not stay live to the end).
  void function_tpc() {
    os << "// ---- emitted: the " << P.fn_outputs.size() << " outputs of the compiled function (forward only) ----\n";
    lookup_helpers();
    os << "RN_DEVICE void rn_function(const double (&q)[RN_NQ], double* RN_RESTRICT out, const long long os, int& err) {\n";
    os << "  (void)q; (void)err;\n";
    std::map<int, std::vector<int>> out_at;
    std::vector<int> tail;
    for (size_t j = 0; j < P.fn_outputs.size(); j++) {
      const Node& n = P.nodes[P.fn_outputs[j]];
      if (n.kind == K_CONST || n.kind == K_INPUT)
        tail.push_back((int)j);
      else
        out_at[P.fn_outputs[j]].push_back((int)j);
    }
    for (int id : P.inv_fwd) {
      stmt(id, "  ");
      auto it = out_at.find(id);
      if (it != out_at.end())
        for (int j : it->second) os << "  out[" << j << "LL * os] = " << val(id) << ";\n";
    }
    for (int j : tail) os << "  out[" << j << "LL * os] = " << val(P.fn_outputs[j]) << ";\n";
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
    std::map<std::vector<int>, int> tab_by_refs;  // Lookups over the same entries (the 8 observe splits) share a table
    for (const TargetInfo& T : P.targets)
      for (int id : T.row_fwd)
        if (is_big_table(P.nodes[id])) {
          const Node& n = P.nodes[id];
          std::vector<int> refs(P.lookup_refs.begin() + n.b, P.lookup_refs.begin() + n.b + n.c);
          auto it = tab_by_refs.find(refs);
          if (it == tab_by_refs.end()) {
            it = tab_by_refs.emplace(refs, tab_doubles).first;
            tab_fill.push_back(id);
            tab_doubles += n.c;
          }
          tab_off[id] = it->second;
        }
    os << "// ---- emitted: log-density and gradient of the frozen DAG (" << (P.symbolic ? "symbolic" : "adjoint")
       << " gradient), warp-per-chain: rows across lanes ----\n";
    const int K = std::max(1, opt.wpc_k);
    int n_reg_acc = 0;
    for (int sl = 0; sl < P.n_slots; sl++)
      if (smem_slot[sl] < 0) n_reg_acc++;
    // cross-warp reduction scratch of the chain's group (K warps): [warp][register accumulators..., err]
    red_off = tab_doubles + n_smem_acc;
    red_doubles = K > 1 ? K * (n_reg_acc + 1) : 0;
    os << "#define RN_WPC_SCRATCH " << (tab_doubles + n_smem_acc + red_doubles) << "\n";
    os << "#define RN_WPC_RED_OFF " << red_off << "\n";
    os << "RN_DEVICE double rn_tab_lookup(const double* tab, int len, int low, double idx, int& err) {\n"
          "  const int k = rn_d2i(idx) - low;\n  if (k < 0 || k >= len) { err |= 1; return RN_NAN; }\n  return tab[k];\n}\n";
    os << "RN_DEVICE double rn_warp_sum(double x) {\n  RN_UNROLL\n  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);\n  return x;\n}\n";
    lookup_helpers();
    os << "RN_DEVICE void rn_density(const double* q, double& dens, double* grad, double* scr, "
          "const double* RN_RESTRICT data, int& err, RnTma& tma) {\n";
    os << "  (void)data; (void)scr; (void)tma;\n  const int lane = (int)(threadIdx.x % RN_G);  // thread of the chain's group\n  (void)lane;\n";
    if (n_smem_acc) os << "  for (int k = lane; k < " << n_smem_acc << "; k += RN_G) scr[" << tab_doubles << " + k] = 0.0;\n";
    for (int id : P.inv_fwd) stmt(id, "  ");
    for (int id : tab_fill) {
      const Node& n = P.nodes[id];
      for (int k = 0; k < n.c; k++) os << "  scr[" << (tab_off.at(id) + k) << "] = " << val(P.lookup_refs[n.b + k]) << ";\n";
    }
    os << "  RN_SYNC();\n";
    for (int sl = 0; sl < P.n_slots; sl++)
      if (smem_slot[sl] < 0) os << "  double a" << sl << " = 0.0;\n";
    for (size_t t = 0; t < P.targets.size(); t++) {
      const TargetInfo& T = P.targets[t];
      os << "  // target " << t << (T.streamed() ? " (streamed, rows across the group's threads)" : " (data-free)") << "\n";
      if (T.streamed()) {
        const unsigned long long base = (unsigned long long)opt.target_base[t], td = (unsigned long long)T.n_cols * 32;
        const unsigned long long rows_per_tile = 32ull * K;  // a "super-tile": K consecutive 32-row tiles, one per warp
        const unsigned long long n_full = T.n_rows / rows_per_tile;
        os << "  {\n    long long row0 = lane;\n";
        if (opt.tma_stages > 0 && n_full > 0) {
          // CTA lockstep over full tiles: tile t+S-1 in flight (one bulk copy) while all warps consume tile t from smem
          os << "    if (tma.on) {\n"
             << "      const unsigned n_full = " << n_full << "u, seq0 = tma.seq;\n"
             << "      const double* RN_RESTRICT src = data + " << base << "ULL;\n"
             << "      if (threadIdx.x == 0)\n"
             << "        for (unsigned p = 0; p + 1 < RN_TMA_STAGES && p < n_full; p++) rn_tma_load(tma, seq0 + p, src + (size_t)p * "
             << td * K << "ULL, " << td * K * 8 << "u);\n"
             << "      for (unsigned tile = 0; tile < n_full; tile++) {\n"
             << "        const unsigned seq = seq0 + tile;\n"
             << "        if (threadIdx.x == 0 && tile + (RN_TMA_STAGES - 1) < n_full)\n"
             << "          rn_tma_load(tma, seq + (RN_TMA_STAGES - 1), src + (size_t)(tile + (RN_TMA_STAGES - 1)) * " << td * K << "ULL, "
             << td * K * 8 << "u);\n"
             << "        rn_mbar_wait(tma.full + (seq % RN_TMA_STAGES), (seq / RN_TMA_STAGES) & 1u);\n"
             << "        const double* rp = tma.stage + (size_t)(seq % RN_TMA_STAGES) * RN_TMA_TILE_DOUBLES + (size_t)(lane >> 5) * " << td
             << " + (lane & 31);\n";
          row_body(
              T, "        ", [&](int k) { return "rp[" + std::to_string(local_col(T, k) * 32) + "]"; },
              [&](int slot) { return acc_ref(slot); }, true, tab_doubles);
          os << "        rn_cta_bar(tma.nthreads);\n"
             << "      }\n"
             << "      tma.seq = seq0 + n_full;\n"
             << "      row0 += " << n_full * rows_per_tile << "LL;\n"
             << "    }\n";
        }
        os << "    for (long long row = row0; row < " << (long long)T.n_rows << "LL; row += RN_G) {\n";
        os << "      const double* RN_RESTRICT rp = data + " << base << "ULL + (row >> 5) * " << td << "LL + (row & 31);\n";
        row_body(
            T, "      ", [&](int k) { return "RN_LDG(rp + " + std::to_string(local_col(T, k) * 32) + ")"; },
            [&](int slot) { return acc_ref(slot); }, true, tab_doubles);
        os << "    }\n  }\n";
      } else {
        os << "  if (lane == 0) {\n";  // counted once by the reduction below
        for (const AccStmt& a : T.row_acc) os << "    " << acc_ref(a.slot) << " += " << val(a.node) << ";\n";
        os << "  }\n";
      }
    }
    for (int sl = 0; sl < P.n_slots; sl++)
      if (smem_slot[sl] < 0) os << "  a" << sl << " = rn_warp_sum(a" << sl << ");\n";
    os << "  err = (int)__reduce_or_sync(0xffffffffu, (unsigned)err);\n";
    if (K > 1) {
      // the K warps of the chain exchange their partial sums through shared memory; every thread adds them in the same
      // order, so all of them hold identical totals afterwards
      const int stride = n_reg_acc + 1;
      os << "  {\n    double* red = scr + " << red_off << ";\n    const int wg = lane >> 5;\n    if ((lane & 31) == 0) {\n";
      int idx = 0;
      for (int sl = 0; sl < P.n_slots; sl++)
        if (smem_slot[sl] < 0) os << "      red[wg * " << stride << " + " << idx++ << "] = a" << sl << ";\n";
      os << "      red[wg * " << stride << " + " << idx << "] = (double)err;\n    }\n    RN_SYNC();\n";
      idx = 0;
      for (int sl = 0; sl < P.n_slots; sl++)
        if (smem_slot[sl] < 0) {
          os << "    a" << sl << " = red[" << idx << "]";
          for (int k = 1; k < K; k++) os << " + red[" << k * stride + idx << "]";
          os << ";\n";
          idx++;
        }
      os << "    for (int k = 0; k < " << K << "; k++) err |= (int)red[k * " << stride << " + " << idx << "];\n  }\n";
    }
    os << "  RN_SYNC();\n";
    os << "  dens = a0;\n";
    if (P.symbolic) {
      for (uint32_t i = 0; i < P.n_params; i++) os << "  if (lane == 0) grad[" << i << "] = " << acc_ref(1 + (int)i) << ";\n";
    } else {
      for (int id : P.inv_bwd) stmt(id, "  ");
      for (uint32_t i = 0; i < P.n_params; i++) os << "  if (lane == 0) grad[" << i << "] = " << val(P.grad_nodes[i]) << ";\n";
    }
    os << "  RN_SYNC();\n}\n";
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

WpcSizes wpc_sizes(const Program& P, const EmitOptions& opt) {
  WpcSizes z;
  Emitter E(P, opt);
  E.density_wpc();
  // chain vectors (q, p, gradient, mass [+ EHMC snapshot]) [+ 2 scratch vectors of the dense mass matrix code] + density scratch
  z.per_warp_doubles = ((opt.enable_ehmc ? 7 : 4) + (opt.mass_max == 2 ? 2 : 0)) * (int)P.n_params + E.tab_doubles + E.n_smem_acc + E.red_doubles;
  for (const TargetInfo& T : P.targets)
    if (T.streamed() && T.n_rows >= 32ull * (uint64_t)std::max(1, opt.wpc_k))
      z.tile_doubles = std::max(z.tile_doubles, (int)T.n_cols * 32 * std::max(1, opt.wpc_k));
  return z;
}

std::string emit_optimizer_source(const Program& P, const EmitOptions& opt, int history) {
  std::ostringstream os;
  os << "// generated by rainier_b200 (CUDA source emitter, optimizer flavour) -- do not edit\n";
  os << "#define RN_N " << P.n_params << "\n";
  os << "#define RN_NSLOTS " << P.n_slots << "\n";
  os << "#define RN_BACKEND " << (opt.backend == 1 ? 1 : 0) << "\n";
  os << "#define RN_LBFGS_M " << history << "\n";
  if (opt.fast_math) os << "#define RN_FAST_MATH 1\n";
  EmitOptions eo = opt;
  if (opt.backend == 1) {  // K warps per start, independent per-warp row loads (no CTA-shared tiles: starts diverge)
    eo.wpc_k = std::max(1, opt.wpc_k);
    eo.tma_stages = 0;
    eo.enable_ehmc = false;
    os << "#define RN_WPC_K " << eo.wpc_k << "\n#define RN_TMA_STAGES 0\n#define RN_TMA_TILE_DOUBLES 0\n";
  } else {
    eo.backend = 0;
  }
  os << kPreludeSource << "\n";
  os << emit_density(P, eo) << "\n" << kOptimizerSource << "\n";
  return os.str();
}

std::string emit_function_source(const Program& P, const EmitOptions& opt) {
  std::ostringstream os;
  os << "// generated by rainier_b200 (CUDA source emitter, function flavour) -- do not edit\n";
  os << "#define RN_N " << P.n_params << "\n";
  os << "#define RN_NQ " << std::max<uint32_t>(1, P.n_params) << "\n";
  os << "#define RN_M " << P.fn_outputs.size() << "\n";
  os << "#define RN_BACKEND 0\n";
  if (opt.fast_math) os << "#define RN_FAST_MATH 1\n";
  os << kPreludeSource << "\n";
  Emitter E(P, opt);
  E.function_tpc();
  os << E.os.str() << "\n" << kFunctionSource << "\n";
  return os.str();
}

std::string emit_source(const Program& P, const EmitOptions& opt) {
  std::ostringstream os;
  os << "// generated by rainier_b200 (CUDA source emitter) -- do not edit\n";
  os << "#define RN_N " << P.n_params << "\n";
  os << "#define RN_NSLOTS " << P.n_slots << "\n";
  os << "#define RN_BACKEND " << opt.backend << "\n";
  os << "#define RN_MASS_MAX " << opt.mass_max << "\n";
  os << "#define RN_ENABLE_EHMC " << (opt.enable_ehmc ? 1 : 0) << "\n";
  if (opt.fast_math) os << "#define RN_FAST_MATH 1\n";
  if (opt.backend == 1) {
    os << "#define RN_WPC_K " << std::max(1, opt.wpc_k) << "\n";
    os << "#define RN_TMA_STAGES " << opt.tma_stages << "\n";
    os << "#define RN_TMA_TILE_DOUBLES " << wpc_sizes(P, opt).tile_doubles << "\n";
  }
  os << kPreludeSource << "\n";
  os << emit_density(P, opt) << "\n";
  if (opt.backend == 1) {
    os << "#define RN_WPC_SMEM_DOUBLES (" << ((opt.enable_ehmc ? 7 : 4) + (opt.mass_max == 2 ? 2 : 0)) << " * RN_N + RN_WPC_SCRATCH)\n";
    os << kSamplerWpcSource << "\n";
  } else {
    os << kSamplerSource << "\n";
  }
  return os.str();
}

}  // namespace rn
