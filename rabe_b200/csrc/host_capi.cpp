// C ABI of the host-side policy layer (include/rabe_b200.h, "host-side policy layer").
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/rabe_b200.h"
#include "host_policy.hpp"
#include "internal.h"

struct rb_policy { rbh::Node root; };

namespace {
std::vector<std::string> to_vec(const char* const* a, uint32_t n) {
  std::vector<std::string> v; v.reserve(n);
  for (uint32_t i = 0; i < n; ++i) v.emplace_back(a[i] ? a[i] : "");
  return v;
}
}  // namespace

extern "C" {

int rb_policy_parse(const char* text, int language, rb_policy** out) {
  if (!text || !out) return RB_EINVAL;
  *out = nullptr;
  rb_policy* p = new (std::nothrow) rb_policy();
  if (!p) return RB_ENOMEM;
  std::string err;
  if (!rbh::parse(text, language, p->root, err)) { delete p; return RB_EPOLICY; }
  *out = p;
  return RB_OK;
}
void rb_policy_free(rb_policy* p) { delete p; }

int rb_policy_serialize(const rb_policy* p, int language, char* out, size_t cap, size_t* needed) {
  if (!p || (language != RB_LANG_JSON && language != RB_LANG_HUMAN)) return RB_EINVAL;
  std::string s = rbh::serialize(p->root, language);
  if (needed) *needed = s.size() + 1;
  if (!out) return RB_OK;
  if (cap < s.size() + 1) return RB_EINVAL;
  memcpy(out, s.c_str(), s.size() + 1);
  return RB_OK;
}

int rb_policy_msp(const rb_policy* p, uint32_t* n1, uint32_t* n2, int8_t* m, size_t m_cap, char* names, size_t names_cap,
                  size_t* names_needed) {
  if (!p) return RB_EINVAL;
  rbh::Msp msp; std::string err;
  if (!rbh::calculate_msp(p->root, msp, err)) return RB_EPOLICY;
  size_t need = 0;
  for (auto& s : msp.pi) need += s.size() + 1;
  if (n1) *n1 = (uint32_t)msp.m.size();
  if (n2) *n2 = (uint32_t)msp.c;
  if (names_needed) *names_needed = need;
  if (m) {
    if (m_cap < msp.m.size() * msp.c) return RB_EINVAL;
    for (size_t i = 0; i < msp.m.size(); ++i) memcpy(m + i * msp.c, msp.m[i].data(), msp.c);
  }
  if (names) {
    if (names_cap < need) return RB_EINVAL;
    char* o = names;
    for (auto& s : msp.pi) { memcpy(o, s.c_str(), s.size() + 1); o += s.size() + 1; }
  }
  return RB_OK;
}

int rb_policy_satisfied(const rb_policy* p, const char* const* attrs, uint32_t n, int* out) {
  if (!p || (!attrs && n) || !out) return RB_EINVAL;
  *out = rbh::traverse_policy(to_vec(attrs, n), p->root) ? 1 : 0;
  return RB_OK;
}

int rb_policy_prune(const rb_policy* p, const char* const* attrs, uint32_t n, int* matched, char* out, size_t cap, size_t* needed,
                    uint32_t* n_items) {
  if (!p || (!attrs && n) || !matched) return RB_EINVAL;
  bool match; std::vector<std::pair<std::string, std::string>> list; std::string err;
  if (!rbh::calc_pruned(to_vec(attrs, n), p->root, match, list, err)) return RB_EPOLICY;
  *matched = match ? 1 : 0;
  size_t need = 0;
  for (auto& it : list) need += it.first.size() + it.second.size() + 2;
  if (needed) *needed = need;
  if (n_items) *n_items = (uint32_t)list.size();
  if (out) {
    if (cap < need) return RB_EINVAL;
    char* o = out;
    for (auto& it : list) {
      memcpy(o, it.first.c_str(), it.first.size() + 1); o += it.first.size() + 1;
      memcpy(o, it.second.c_str(), it.second.size() + 1); o += it.second.size() + 1;
    }
  }
  return RB_OK;
}

int rb_hash_to_fr(const char* s, size_t len, uint8_t out[32]) {
  if ((!s && len) || !out) return RB_EINVAL;
  rbh::hash_to_fr(std::string(s ? s : "", len), out);
  return RB_OK;
}

int rb_ac17_msp_from_policy(rb_ctx* c, const rb_policy* p, rb_msp** out) {
  if (!c || !p || !out) return RB_EINVAL;
  *out = nullptr;
  rbh::Msp msp; std::string err;
  if (!rbh::calculate_msp(p->root, msp, err)) return RB_EPOLICY;
  const size_t n1 = msp.m.size(), n2 = msp.c;
  std::vector<int8_t> m(n1 * n2);
  for (size_t i = 0; i < n1; ++i) memcpy(m.data() + i * n2, msp.m[i].data(), n2);
  std::vector<uint8_t> h_row(n1 * 6 * 32), h_col(n2 * 6 * 32);
  for (size_t i = 0; i < n1; ++i)
    for (int l = 0; l < 3; ++l)
      for (int t = 0; t < 2; ++t)                                   // ac17/mod.rs:333-339
        rbh::hash_to_fr(msp.pi[i] + std::to_string(l) + std::to_string(t), h_row.data() + 32 * (i * 6 + l * 2 + t));
  for (size_t j = 0; j < n2; ++j)
    for (int l = 0; l < 3; ++l)
      for (int t = 0; t < 2; ++t)                                   // ac17/mod.rs:305-328
        rbh::hash_to_fr("0" + std::to_string(j + 1) + std::to_string(l) + std::to_string(t), h_col.data() + 32 * (j * 6 + l * 2 + t));
  return rb_msp_load(c, (uint32_t)n1, (uint32_t)n2, m.data(), h_row.data(), h_col.data(), out);
}

int rb_ac17_attr_hashes(const char* const* attrs, uint32_t n, uint8_t* h_attr, uint8_t h_01[192]) {
  if ((!attrs && n) || (!h_attr && n) || !h_01) return RB_EINVAL;
  for (uint32_t x = 0; x < n; ++x) {
    if (!attrs[x]) return RB_EINVAL;
    for (int l = 0; l < 3; ++l)
      for (int t = 0; t < 2; ++t)                                   // ac17/mod.rs:231-235
        rbh::hash_to_fr(std::string(attrs[x]) + std::to_string(l) + std::to_string(t), h_attr + 32 * ((size_t)x * 6 + l * 2 + t));
  }
  for (int l = 0; l < 3; ++l)
    for (int t = 0; t < 2; ++t)                                     // ac17/mod.rs:250-254
      rbh::hash_to_fr("01" + std::to_string(l) + std::to_string(t), h_01 + 32 * (l * 2 + t));
  return RB_OK;
}

int rb_ac17_decrypt_lists(const rb_policy* p, const char* const* sk_attrs, uint32_t n_sk, const char* const* ct_names, uint32_t n_ct,
                          int* matched, uint32_t* ct_idx, size_t ct_cap, uint32_t* n_ct_idx, uint32_t* sk_idx, size_t sk_cap,
                          uint32_t* n_sk_idx) {
  if (!p || (!sk_attrs && n_sk) || (!ct_names && n_ct) || !matched || !n_ct_idx || !n_sk_idx) return RB_EINVAL;
  std::vector<std::string> attrs = to_vec(sk_attrs, n_sk), rows = to_vec(ct_names, n_ct);
  *n_ct_idx = 0; *n_sk_idx = 0;
  if (!rbh::traverse_policy(attrs, p->root)) { *matched = 0; return RB_OK; }           // ac17/mod.rs:389
  bool match; std::vector<std::pair<std::string, std::string>> list; std::string err;
  if (!rbh::calc_pruned(attrs, p->root, match, list, err)) return RB_EPOLICY;
  *matched = match ? 1 : 0;
  if (!match) return RB_OK;
  size_t nc = 0, ns = 0;
  for (auto& cur : list) {                                                                // ac17/mod.rs:404-413
    for (uint32_t i = 0; i < n_ct; ++i) if (rows[i] == cur.first) { if (ct_idx) { if (nc >= ct_cap) return RB_EINVAL; ct_idx[nc] = i; } ++nc; }
    for (uint32_t i = 0; i < n_sk; ++i) if (attrs[i] == cur.first) { if (sk_idx) { if (ns >= sk_cap) return RB_EINVAL; sk_idx[ns] = i; } ++ns; }
  }
  *n_ct_idx = (uint32_t)nc; *n_sk_idx = (uint32_t)ns;
  return RB_OK;
}

}  // extern "C"

namespace {
// DFS flattening shared by the share plan (mode 0: one term per polynomial coefficient) and the
// reconstruction coefficients (mode 1: one term per AND gate on the path).
struct Flat { std::vector<uint32_t> terms, leaf_offs; uint32_t n_coefs = 0; bool ok = true; };
struct PathGate { uint32_t coef_base, k, child; };
void flatten(const rbh::Node& n, int mode, std::vector<PathGate>& path, Flat& f) {
  if (n.kind == rbh::LEAF) {
    for (const PathGate& g : path) {
      if (mode == 0) for (uint32_t i = 1; i < g.k; ++i) { f.terms.insert(f.terms.end(), {g.coef_base + i - 1, g.child + 1, i, 0u}); }
      else f.terms.insert(f.terms.end(), {0u, g.child + 1, g.k, 0u});
    }
    f.leaf_offs.push_back((uint32_t)(f.terms.size() / 4));
    return;
  }
  if (n.kids.empty()) { f.ok = false; return; }
  if (n.kind == rbh::AND) {
    PathGate g{f.n_coefs, (uint32_t)n.kids.size(), 0};
    f.n_coefs += (uint32_t)n.kids.size() - 1;            // gen_shares draws k-1 coefficients before recursing
    for (uint32_t c = 0; c < n.kids.size(); ++c) { g.child = c; path.push_back(g); flatten(n.kids[c], mode, path, f); path.pop_back(); }
  } else {
    for (const rbh::Node& k : n.kids) flatten(k, mode, path, f);   // OR: (1,n) sharing, every child gets the secret
  }
}
}  // namespace

extern "C" {

int rb_policy_leaf_labels(const rb_policy* p, char* out, size_t cap, size_t* needed, uint32_t* n_leaves) {
  if (!p) return RB_EINVAL;
  std::vector<const rbh::Node*> leaves; rbh::leaves_dfs(p->root, leaves);
  size_t need = 0;
  for (auto* l : leaves) need += rbh::node_index(*l).size() + 1;
  if (needed) *needed = need;
  if (n_leaves) *n_leaves = (uint32_t)leaves.size();
  if (out) {
    if (cap < need) return RB_EINVAL;
    char* o = out;
    for (auto* l : leaves) { std::string s = rbh::node_index(*l); memcpy(o, s.c_str(), s.size() + 1); o += s.size() + 1; }
  }
  return RB_OK;
}

int rb_share_plan_create(rb_ctx* c, const rb_policy* p, rb_share_plan** out) {
  if (!c || !p || !out) return RB_EINVAL;
  Flat f; f.leaf_offs.push_back(0);
  std::vector<PathGate> path;
  flatten(p->root, 0, path, f);
  if (!f.ok) return RB_EPOLICY;
  return rb_share_plan_create_raw(c, f.terms.data(), (uint32_t)(f.terms.size() / 4), f.leaf_offs.data(), (uint32_t)(f.leaf_offs.size() - 1), f.n_coefs, out);
}

int rb_policy_coefficients(rb_ctx* c, const rb_policy* p, uint8_t* out) {
  if (!c || !p || !out) return RB_EINVAL;
  Flat f; f.leaf_offs.push_back(0);
  std::vector<PathGate> path;
  flatten(p->root, 1, path, f);
  if (!f.ok) return RB_EPOLICY;
  return rb_lagrange_raw(c, f.terms.data(), (uint32_t)(f.terms.size() / 4), f.leaf_offs.data(), (uint32_t)(f.leaf_offs.size() - 1), out);
}

}  // extern "C"
