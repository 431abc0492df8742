// rn_inline.cpp -- see rn_inline.hpp.
#include "rn_inline.hpp"

#include <algorithm>
#include <cstring>
#include <functional>
#include <map>

namespace rn {

namespace {
typedef std::vector<int32_t> Mono;           // sorted ids of column-only nodes; empty = parameter-only term
typedef std::map<Mono, int32_t> Form;        // monomial -> parameter-only coefficient node
}  // namespace

std::string plan_inline(const void* rir, size_t len, InlinePlan& P, int max_degree, int max_monomials) {
  const uint8_t* p = (const uint8_t*)rir;
  const uint8_t* end = p + len;
  if (len < sizeof(rir_header)) return "RIR: truncated header";
  std::memcpy(&P.h, p, sizeof(rir_header));
  p += sizeof(rir_header);
  if (P.h.magic != RIR_MAGIC || P.h.version != RIR_VERSION) return "RIR: bad magic/version";
  if ((size_t)(end - p) < (size_t)P.h.n_nodes * sizeof(rir_node)) return "RIR: truncated node array";
  P.nodes.resize(P.h.n_nodes);
  std::memcpy(P.nodes.data(), p, (size_t)P.h.n_nodes * sizeof(rir_node));
  p += (size_t)P.h.n_nodes * sizeof(rir_node);
  const size_t lrb = ((size_t)P.h.n_lookup_refs * 4 + 7) & ~(size_t)7;
  if ((size_t)(end - p) < lrb) return "RIR: truncated lookup refs";
  P.lookup_refs.resize(P.h.n_lookup_refs);
  if (P.h.n_lookup_refs) std::memcpy(P.lookup_refs.data(), p, (size_t)P.h.n_lookup_refs * 4);
  p += lrb;
  P.targets.resize(P.h.n_targets);
  for (auto& T : P.targets) {
    if ((size_t)(end - p) < sizeof(rir_target)) return "RIR: truncated target";
    std::memcpy(&T.t, p, sizeof(rir_target));
    p += sizeof(rir_target);
    const size_t ob = ((size_t)T.t.n_outputs * 4 + 7) & ~(size_t)7;
    if ((size_t)(end - p) < ob) return "RIR: truncated outputs";
    T.outputs.resize(T.t.n_outputs);
    std::memcpy(T.outputs.data(), p, (size_t)T.t.n_outputs * 4);
    p += ob;
  }
  P.inl.clear();
  if (P.h.flags & (RIR_FLAG_GRADIENT | RIR_FLAG_FUNCTION)) return "";  // symbolic-gradient containers stay as they are

  const int N0 = (int)P.nodes.size();
  const int np = (int)P.h.n_params;
  // parameter / column dependence and owning target of every node (operands precede their users)
  std::vector<char> pdep(N0, 0), cdep(N0, 0);
  for (int i = 0; i < N0; i++) {
    const rir_node& n = P.nodes[i];
    auto dep = [&](int x) {
      if (x < 0 || x >= i) return;
      pdep[i] |= pdep[x];
      cdep[i] |= cdep[x];
    };
    switch (n.kind) {
      case RIR_INPUT: (n.a < np ? pdep[i] : cdep[i]) = 1; break;
      case RIR_UNARY: dep(n.a); break;
      case RIR_BINARY: dep(n.a); dep(n.b); break;
      case RIR_LOOKUP:
        dep(n.a);
        for (int k = 0; k < n.c; k++) dep(P.lookup_refs[n.b + k]);
        break;
      default: break;
    }
  }
  auto is_pdep = [&](int id) { return id < N0 ? pdep[id] != 0 : true; };   // appended nodes: tracked below
  std::vector<char> app_cdep;  // column dependence of appended nodes (index id - N0)
  auto is_cdep = [&](int id) { return id < N0 ? cdep[id] != 0 : app_cdep[id - N0] != 0; };
  (void)is_pdep;
  auto add = [&](uint8_t kind, uint8_t op, int32_t a, int32_t b, double value, bool coldep) {
    rir_node n;
    std::memset(&n, 0, sizeof(n));
    n.kind = kind;
    n.op = op;
    n.a = a;
    n.b = b;
    n.value = value;
    P.nodes.push_back(n);
    app_cdep.push_back(coldep ? 1 : 0);
    return (int32_t)P.nodes.size() - 1;
  };
  int32_t one = -1, minus_one = -1;
  auto cst1 = [&]() { return one >= 0 ? one : (one = add(RIR_CONST, 0, 0, 0, 1.0, false)); };
  auto cstm1 = [&]() { return minus_one >= 0 ? minus_one : (minus_one = add(RIR_CONST, 0, 0, 0, -1.0, false)); };
  auto mul = [&](int32_t a, int32_t b) {
    if (a == one && one >= 0) return b;
    if (b == one && one >= 0) return a;
    return add(RIR_BINARY, RIR_B_MUL, a, b, 0.0, false);
  };
  auto addn = [&](int32_t a, int32_t b) { return add(RIR_BINARY, RIR_B_ADD, a, b, 0.0, false); };

  for (size_t t = 0; t < P.targets.size(); t++) {
    const auto& T = P.targets[t];
    if (T.t.n_cols == 0 || T.t.n_rows == 0 || T.outputs.size() != 1) continue;
    const size_t nodes_before = P.nodes.size();
    std::map<int, Form> form;
    bool ok = true;
    std::function<const Form*(int)> get = [&](int id) -> const Form* {
      auto it = form.find(id);
      if (it != form.end()) return &it->second;
      Form f;
      if (!is_cdep(id)) {
        f[Mono()] = id;  // parameter-only or constant: a coefficient of the empty monomial
      } else if (id < N0 && !pdep[id]) {
        f[Mono(1, id)] = cst1();  // column-only: an atom, however complicated
      } else {
        const rir_node n = P.nodes[id];
        auto multiply = [&](const Form& a, const Form& b, Form& out) {
          for (const auto& x : a)
            for (const auto& y : b) {
              Mono m(x.first);
              m.insert(m.end(), y.first.begin(), y.first.end());
              std::sort(m.begin(), m.end());
              if ((int)m.size() > max_degree) return false;
              const int32_t c = mul(x.second, y.second);
              auto e = out.find(m);
              if (e == out.end())
                out[m] = c;
              else
                e->second = addn(e->second, c);
              if ((int)out.size() > max_monomials) return false;
            }
          return true;
        };
        bool good = false;
        if (n.kind == RIR_BINARY && (n.op == RIR_B_ADD || n.op == RIR_B_SUB)) {
          const Form* a = get(n.a);
          const Form* b = ok && a ? get(n.b) : nullptr;
          if (a && b) {
            f = *a;
            for (const auto& y : *b) {
              const int32_t c = n.op == RIR_B_SUB ? mul(cstm1(), y.second) : y.second;
              auto e = f.find(y.first);
              if (e == f.end())
                f[y.first] = c;
              else
                e->second = addn(e->second, c);
            }
            good = (int)f.size() <= max_monomials;
          }
        } else if (n.kind == RIR_BINARY && n.op == RIR_B_MUL) {
          const Form* a = get(n.a);
          const Form* b = ok && a ? get(n.b) : nullptr;
          good = a && b && multiply(*a, *b, f);
        } else if (n.kind == RIR_BINARY && n.op == RIR_B_DIV && !is_cdep(n.b)) {  // division by a parameter-only value scales
          const Form* a = get(n.a);
          if (a) {
            for (const auto& x : *a) f[x.first] = add(RIR_BINARY, RIR_B_DIV, x.second, n.b, 0.0, false);
            good = true;
          }
        } else if (n.kind == RIR_BINARY && n.op == RIR_B_DIV && n.b < N0 && !pdep[n.b]) {  // by a column-only value: times 1/value
          const Form* a = get(n.a);
          if (a) {
            const int32_t recip = add(RIR_BINARY, RIR_B_DIV, cst1(), n.b, 0.0, true);
            Form r;
            r[Mono(1, recip)] = cst1();
            good = multiply(*a, r, f);
          }
        } else if (n.kind == RIR_BINARY && n.op == RIR_B_POW && P.nodes[n.b].kind == RIR_CONST && P.nodes[n.b].value == 2.0) {
          const Form* a = get(n.a);
          good = a && multiply(*a, *a, f);
        } else if (n.kind == RIR_UNARY && n.op == RIR_U_NOOP) {
          const Form* a = get(n.a);
          if (a) {
            f = *a;
            good = true;
          }
        }
        if (!good) {  // a nonlinear operation on a value that mixes parameters and columns: the target stays streamed
          ok = false;
          return nullptr;
        }
      }
      return &form.emplace(id, std::move(f)).first->second;
    };
    const Form* out = get((int)T.outputs[0]);
    if (!ok || !out) {
      P.nodes.resize(nodes_before);  // drop what the attempt appended
      app_cdep.resize(nodes_before - N0);
      if (one >= (int32_t)nodes_before) one = -1;
      if (minus_one >= (int32_t)nodes_before) minus_one = -1;
      continue;
    }
    InlineTarget I;
    I.target = (int)t;
    for (const auto& x : *out) {
      if (x.first.empty()) {
        I.const_coef = x.second;
      } else {
        I.monos.push_back(x.first);
        I.coef.push_back(x.second);
      }
    }
    P.inl.push_back(std::move(I));
  }
  return "";
}

std::vector<uint8_t> inline_function_rir(const InlinePlan& P, size_t k) {
  const InlineTarget& I = P.inl[k];
  const auto& T = P.targets[I.target];
  // the column-only sub-DAG the monomials need, renumbered; inputs: column j of the target -> function input j
  std::map<int32_t, int32_t> id;  // old node -> new node
  std::vector<rir_node> nodes;
  std::vector<int32_t> refs;
  std::function<int32_t(int32_t)> visit = [&](int32_t old) -> int32_t {
    auto it = id.find(old);
    if (it != id.end()) return it->second;
    rir_node n = P.nodes[old];
    switch (n.kind) {
      case RIR_INPUT: n.a = n.a - (int32_t)T.t.first_input; break;
      case RIR_UNARY: n.a = visit(n.a); break;
      case RIR_BINARY: {
        const int32_t a = visit(n.a), b = visit(n.b);
        n.a = a;
        n.b = b;
        break;
      }
      case RIR_LOOKUP: {
        const int32_t a = visit(n.a);
        std::vector<int32_t> e;
        for (int j = 0; j < n.c; j++) e.push_back(visit(P.lookup_refs[n.b + j]));
        n.a = a;
        n.b = (int32_t)refs.size();
        refs.insert(refs.end(), e.begin(), e.end());
        break;
      }
      default: break;
    }
    nodes.push_back(n);
    return id[old] = (int32_t)nodes.size() - 1;
  };
  std::vector<uint32_t> outs;
  for (const Mono& m : I.monos) {
    int32_t acc = visit(m[0]);
    for (size_t j = 1; j < m.size(); j++) {
      const int32_t b = visit(m[j]);
      rir_node n;
      std::memset(&n, 0, sizeof(n));
      n.kind = RIR_BINARY;
      n.op = RIR_B_MUL;
      n.a = acc;
      n.b = b;
      nodes.push_back(n);
      acc = (int32_t)nodes.size() - 1;
    }
    outs.push_back((uint32_t)acc);
  }
  rir_header h;
  std::memset(&h, 0, sizeof(h));
  h.magic = RIR_MAGIC;
  h.version = RIR_VERSION;
  h.n_params = T.t.n_cols;
  h.n_inputs = T.t.n_cols;
  h.n_nodes = (uint32_t)nodes.size();
  h.n_targets = 1;
  h.n_lookup_refs = (uint32_t)refs.size();
  h.flags = RIR_FLAG_FUNCTION;
  std::vector<uint8_t> out;
  auto put = [&](const void* src, size_t n) { out.insert(out.end(), (const uint8_t*)src, (const uint8_t*)src + n); };
  put(&h, sizeof(h));
  put(nodes.data(), nodes.size() * sizeof(rir_node));
  put(refs.data(), refs.size() * 4);
  while (out.size() % 8) out.push_back(0);
  rir_target rt;
  std::memset(&rt, 0, sizeof(rt));
  rt.n_outputs = (uint32_t)outs.size();
  put(&rt, sizeof(rt));
  put(outs.data(), outs.size() * 4);
  while (out.size() % 8) out.push_back(0);
  return out;
}

std::vector<uint8_t> apply_inline(const InlinePlan& P, const std::vector<std::vector<double>>& sums) {
  std::vector<rir_node> nodes = P.nodes;
  std::vector<InlinePlan::RawTarget> targets = P.targets;
  auto add = [&](uint8_t kind, uint8_t op, int32_t a, int32_t b, double value) {
    rir_node n;
    std::memset(&n, 0, sizeof(n));
    n.kind = kind;
    n.op = op;
    n.a = a;
    n.b = b;
    n.value = value;
    nodes.push_back(n);
    return (int32_t)nodes.size() - 1;
  };
  for (size_t k = 0; k < P.inl.size(); k++) {
    const InlineTarget& I = P.inl[k];
    auto& T = targets[I.target];
    int32_t acc = -1;
    auto term = [&](double s, int32_t coef) {
      const int32_t v = add(RIR_BINARY, RIR_B_MUL, add(RIR_CONST, 0, 0, 0, s), coef, 0.0);
      acc = acc < 0 ? v : add(RIR_BINARY, RIR_B_ADD, acc, v, 0.0);
    };
    if (I.const_coef >= 0) term((double)T.t.n_rows, I.const_coef);
    for (size_t m = 0; m < I.monos.size(); m++) term(sums[k][m], I.coef[m]);
    if (acc < 0) acc = add(RIR_CONST, 0, 0, 0, 0.0);
    T.outputs[0] = (uint32_t)acc;
    T.t.n_rows = 0;  // data-free from here on: its column placeholders stay in the input vector, unread
  }
  rir_header h = P.h;
  h.n_nodes = (uint32_t)nodes.size();
  std::vector<uint8_t> out;
  auto put = [&](const void* src, size_t n) { out.insert(out.end(), (const uint8_t*)src, (const uint8_t*)src + n); };
  put(&h, sizeof(h));
  put(nodes.data(), nodes.size() * sizeof(rir_node));
  put(P.lookup_refs.data(), P.lookup_refs.size() * 4);
  while (out.size() % 8) out.push_back(0);
  for (const auto& T : targets) {
    put(&T.t, sizeof(rir_target));
    put(T.outputs.data(), T.outputs.size() * 4);
    while (out.size() % 8) out.push_back(0);
  }
  return out;
}

}  // namespace rn
