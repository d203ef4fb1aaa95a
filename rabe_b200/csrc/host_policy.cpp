// Host-side policy layer (see host_policy.hpp for the reference map).  Recursive-descent parsers
// with PEG semantics (ordered choice, no backtracking into a committed choice, implicit
// WHITESPACE/COMMENT skipping between tokens of non-atomic rules).
#include "host_policy.hpp"

#include <algorithm>
#include <cstring>
#include <numeric>

namespace rbh {

namespace {

struct Parser {
  const std::string& s;
  bool number_leaf = false;        // the reference panics on numeric leaves; reported as an error
  explicit Parser(const std::string& text) : s(text) {}

  size_t col(size_t pos) const {   // 1-based column in characters since the last '\n'
    size_t start = 0;
    for (size_t i = pos; i > 0; --i) if (s[i - 1] == '\n') { start = i; break; }
    size_t n = 0;
    for (size_t i = start; i < pos; ++i) if ((static_cast<unsigned char>(s[i]) & 0xC0) != 0x80) ++n;
    return n + 1;
  }
  size_t skip(size_t pos) const {
    for (;;) {
      if (pos < s.size() && (s[pos] == ' ' || s[pos] == '\t' || s[pos] == '\r' || s[pos] == '\n')) { ++pos; continue; }
      if (s.compare(pos, 2, "/*") == 0) {
        size_t e = s.find("*/", pos + 2);
        if (e == std::string::npos) return pos;
        pos = e + 2; continue;
      }
      return pos;
    }
  }
  bool lit(size_t pos, const char* const* words, size_t& end) const {
    for (; *words; ++words) {
      size_t n = strlen(*words);
      if (s.compare(pos, n, *words) == 0) { end = pos + n; return true; }
    }
    return false;
  }
  static bool is_hex(char c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F'); }
  bool string_leaf(size_t pos, Node& out, size_t& end) const {
    if (pos >= s.size() || s[pos] != '"') return false;
    size_t i = pos + 1, start = i;
    while (i < s.size()) {
      char ch = s[i];
      if (ch == '"') { out = Node(); out.kind = LEAF; out.name = s.substr(start, i - start); out.col = col(start); end = i + 1; return true; }
      if (ch == '\\') {
        if (i + 1 < s.size() && strchr("\"\\/bfnrt", s[i + 1])) i += 2;
        else if (i + 5 < s.size() && s[i + 1] == 'u' && is_hex(s[i + 2]) && is_hex(s[i + 3]) && is_hex(s[i + 4]) && is_hex(s[i + 5])) i += 6;
        else return false;
      } else ++i;
    }
    return false;
  }
  bool looks_like_number(size_t pos) const { return pos < s.size() && (s[pos] == '-' || (s[pos] >= '0' && s[pos] <= '9')); }
  // andinner / orinner / NAME / CHILDREN: bare keyword, or QUOTE ~ keyword ~ QUOTE
  bool keyword(size_t pos, const char* const* words, size_t& end) const {
    if (lit(pos, words, end)) return true;
    if (pos < s.size() && s[pos] == '"') {
      size_t p = skip(pos + 1), e;
      if (lit(p, words, e)) { p = skip(e); if (p < s.size() && s[p] == '"') { end = p + 1; return true; } }
    }
    return false;
  }
};

const char* const W_AND[] = {"and", "AND", "&&", nullptr};
const char* const W_OR[] = {"or", "OR", "||", nullptr};
const char* const W_NAME[] = {"name", "NAME", nullptr};
const char* const W_CHILDREN[] = {"children", "CHILDREN", nullptr};

struct Human : Parser {
  using Parser::Parser;
  bool node(size_t pos, Node& out, size_t& end);
  bool value(size_t pos, Node& out, size_t& end) {
    if (string_leaf(pos, out, end)) return true;
    if (looks_like_number(pos)) { number_leaf = true; return false; }
    if (pos < s.size() && (s[pos] == '(' || s[pos] == '[' || s[pos] == '{')) {
      size_t e;
      if (node(skip(pos + 1), out, e)) {
        size_t p = skip(e);
        if (p < s.size() && (s[p] == ')' || s[p] == ']' || s[p] == '}')) { end = p + 1; return true; }
      }
    }
    return false;
  }
  bool term(size_t pos, Node& out, size_t& end) {
    if (value(pos, out, end)) return true;
    if (pos < s.size() && s[pos] == '(') {
      size_t e;
      if (node(skip(pos + 1), out, e)) {
        size_t p = skip(e);
        if (p < s.size() && s[p] == ')') { end = p + 1; return true; }
      }
    }
    return false;
  }
  bool gate(size_t pos, const char* const* words, Kind kind, Node& out, size_t& end) {
    Node first; size_t p;
    if (!term(pos, first, p)) return false;
    Node g; g.kind = kind; g.kids.push_back(std::move(first));
    for (;;) {
      size_t q = skip(p), e;
      if (!keyword(q, words, e)) break;
      Node nxt; size_t e2;
      if (!term(skip(e), nxt, e2)) break;
      g.kids.push_back(std::move(nxt)); p = e2;
    }
    if (g.kids.size() < 2) return false;
    out = std::move(g); end = p; return true;
  }
};
bool Human::node(size_t pos, Node& out, size_t& end) {
  return gate(pos, W_AND, AND, out, end) || gate(pos, W_OR, OR, out, end) || term(pos, out, end);
}

struct Json : Parser {
  using Parser::Parser;
  bool tok(size_t pos, char ch, size_t& end) const { if (pos < s.size() && s[pos] == ch) { end = pos + 1; return true; } return false; }
  bool node(size_t pos, Node& out, size_t& end);
  bool gate(size_t pos, const char* const* words, Kind kind, Node& out, size_t& end) {
    size_t p;
    if (!keyword(pos, words, p)) return false;
    if (!tok(skip(p), ',', p)) return false;
    if (!keyword(skip(p), W_CHILDREN, p)) return false;
    if (!tok(skip(p), ':', p)) return false;
    if (!tok(skip(p), '[', p)) return false;
    Node g; g.kind = kind;
    size_t q;
    if (tok(skip(p), ']', q)) { out = std::move(g); end = q; return true; }
    Node first;
    if (!node(skip(p), first, p)) return false;
    g.kids.push_back(std::move(first));
    for (;;) {
      size_t c;
      if (!tok(skip(p), ',', c)) break;
      Node nxt; size_t e;
      if (!node(skip(c), nxt, e)) break;
      g.kids.push_back(std::move(nxt)); p = e;
    }
    if (!tok(skip(p), ']', p)) return false;
    out = std::move(g); end = p; return true;
  }
};
bool Json::node(size_t pos, Node& out, size_t& end) {
  size_t p;
  if (!tok(pos, '{', p)) return false;
  if (!keyword(skip(p), W_NAME, p)) return false;
  if (!tok(skip(p), ':', p)) return false;
  size_t body = skip(p), e, close;
  Node n;
  if (string_leaf(body, n, e) && tok(skip(e), '}', close)) { out = std::move(n); end = close; return true; }
  if (gate(body, W_AND, AND, n, e) && tok(skip(e), '}', close)) { out = std::move(n); end = close; return true; }
  if (gate(body, W_OR, OR, n, e) && tok(skip(e), '}', close)) { out = std::move(n); end = close; return true; }
  if (looks_like_number(body)) number_leaf = true;
  return false;
}

}  // namespace

bool parse(const std::string& text, int lang, Node& out, std::string& err) {
  size_t end = 0; bool ok; bool number;
  if (lang == LANG_HUMAN) { Human p(text); ok = p.node(p.skip(0), out, end) && p.skip(end) == text.size(); number = p.number_leaf; }
  else if (lang == LANG_JSON) { Json p(text); ok = p.node(p.skip(0), out, end) && p.skip(end) == text.size(); number = p.number_leaf; }
  else { err = "unknown policy language"; return false; }
  if (!ok) { err = number ? "number leaves are not supported (the reference panics on them)" : "policy parse error"; return false; }
  return true;
}

std::string serialize(const Node& n, int lang) {       // pest/mod.rs:68-111
  if (lang == LANG_JSON) {
    if (n.kind == LEAF) return "{\"name\": \"" + n.name + "\"}";
    std::string in;
    for (size_t i = 0; i < n.kids.size(); ++i) { if (i) in += ", "; in += serialize(n.kids[i], lang); }
    return std::string("{\"name\": \"") + (n.kind == AND ? "and" : "or") + "\", \"children\": [" + in + "]}";
  }
  if (n.kind == LEAF) return n.name;
  std::string in;
  for (size_t i = 0; i < n.kids.size(); ++i) { if (i) in += (n.kind == AND ? " and " : " or "); in += serialize(n.kids[i], lang); }
  return "(" + in + ")";
}

namespace {
bool lw(Msp& msp, const Node& p, const std::vector<int8_t>& v, std::string& err) {      // msp.rs:102-147
  if (p.kind == LEAF) {
    msp.m.insert(msp.m.begin(), v);
    msp.pi.insert(msp.pi.begin(), p.name);
    return true;
  }
  if (p.kids.size() < 2) { err = "lw: policy with just a single attribute is not allowed"; return false; }
  if (p.kind == OR) {
    bool ret = true;
    for (const Node& k : p.kids) { if (!lw(msp, k, v, err)) { if (!err.empty()) return false; ret = false; } }
    return ret;
  }
  if (p.kids.size() != 2) { err = "lw: Invalid policy. Number of arguments under AND != 2"; return false; }
  std::vector<int8_t> right = v, left;
  right.resize(msp.c, 0); right.push_back(1);
  left.resize(msp.c, 0); left.push_back(-1);
  msp.c += 1;
  return lw(msp, p.kids[0], right, err) && lw(msp, p.kids[1], left, err);
}
}  // namespace

bool calculate_msp(const Node& root, Msp& out, std::string& err) {                       // msp.rs:78-99
  out = Msp();
  std::vector<int8_t> v{1};
  if (!lw(out, root, v, err)) { if (err.empty()) err = "lewko waters algorithm failed =("; return false; }
  for (auto& row : out.m) row.resize(out.c, 0);
  std::vector<size_t> order(out.pi.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return out.pi[a] < out.pi[b]; });
  Msp sorted; sorted.c = out.c;
  for (size_t i : order) { sorted.m.push_back(out.m[i]); sorted.pi.push_back(out.pi[i]); }
  out = std::move(sorted);
  return true;
}

std::string node_index(const Node& leaf) { return leaf.name + "_" + std::to_string(leaf.col); }
std::string remove_index(const std::string& label) { return label.substr(0, label.find('_')); }

void leaves_dfs(const Node& n, std::vector<const Node*>& out) {
  if (n.kind == LEAF) { out.push_back(&n); return; }
  for (const Node& k : n.kids) leaves_dfs(k, out);
}

bool traverse_policy(const std::vector<std::string>& attrs, const Node& n) {             // tools/mod.rs:31-61
  if (attrs.empty()) return false;
  if (n.kind == LEAF) return std::find(attrs.begin(), attrs.end(), n.name) != attrs.end();
  if (n.kind == AND) { bool r = true; for (const Node& k : n.kids) r &= traverse_policy(attrs, k); return r; }
  bool r = false; for (const Node& k : n.kids) r |= traverse_policy(attrs, k); return r;
}

bool calc_pruned(const std::vector<std::string>& attrs, const Node& n, bool& match,
                 std::vector<std::pair<std::string, std::string>>& list, std::string& err) {   // secretsharing:143-201
  list.clear();
  if (n.kind == LEAF) {
    match = std::find(attrs.begin(), attrs.end(), n.name) != attrs.end();
    if (match) list.emplace_back(n.name, node_index(n));
    return true;
  }
  if (n.kids.size() < 2) { err = "Invalid policy (gate with just a single child)"; return false; }
  if (n.kind == AND) {
    bool ok = true;
    for (const Node& k : n.kids) {
      bool found; std::vector<std::pair<std::string, std::string>> sub;
      if (!calc_pruned(attrs, k, found, sub, err)) return false;
      ok = ok && found;
      if (ok) list.insert(list.end(), sub.begin(), sub.end());
    }
    if (!ok) list.clear();
    match = ok; return true;
  }
  for (const Node& k : n.kids) {
    bool found; std::vector<std::pair<std::string, std::string>> sub;
    if (!calc_pruned(attrs, k, found, sub, err)) return false;
    if (found) { list = std::move(sub); match = true; return true; }
  }
  match = false; return true;
}

// --------------------------------------------------------------------------------------------- SHA3-256 (FIPS 202)
namespace {
inline uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
void keccak_f1600(uint64_t a[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull, 0x0000000080000001ull,
      0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
      0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull,
      0x000000000000800aull, 0x800000008000000aull, 0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  static const int RHO[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
  static const int PI[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
  for (int r = 0; r < 24; ++r) {
    uint64_t c[5];
    for (int x = 0; x < 5; ++x) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; ++x) { uint64_t d = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1); for (int y = 0; y < 25; y += 5) a[y + x] ^= d; }
    uint64_t cur = a[1];
    for (int i = 0; i < 24; ++i) { int j = PI[i]; uint64_t t = a[j]; a[j] = rotl64(cur, RHO[i]); cur = t; }
    for (int y = 0; y < 25; y += 5) {
      uint64_t row[5]; for (int x = 0; x < 5; ++x) row[x] = a[y + x];
      for (int x = 0; x < 5; ++x) a[y + x] = row[x] ^ (~row[(x + 1) % 5] & row[(x + 2) % 5]);
    }
    a[0] ^= RC[r];
  }
}
}  // namespace

void sha3_256(const uint8_t* data, size_t len, uint8_t out[32]) {
  const size_t rate = 136;
  uint64_t st[25]; memset(st, 0, sizeof st);
  while (len >= rate) {
    for (size_t i = 0; i < rate / 8; ++i) { uint64_t w; memcpy(&w, data + 8 * i, 8); st[i] ^= w; }
    keccak_f1600(st); data += rate; len -= rate;
  }
  uint8_t block[136]; memset(block, 0, sizeof block);
  if (len) memcpy(block, data, len);
  block[len] ^= 0x06; block[rate - 1] ^= 0x80;
  for (size_t i = 0; i < rate / 8; ++i) { uint64_t w; memcpy(&w, block + 8 * i, 8); st[i] ^= w; }
  keccak_f1600(st);
  memcpy(out, st, 32);
}

void hash_to_fr(const std::string& s, uint8_t out[32]) {
  // Fr::from_slice(SHA3-256(s)): 256-bit big-endian integer reduced mod r (at most 5 subtractions)
  static const uint8_t R_BE[32] = {0x30, 0x64, 0x4e, 0x72, 0xe1, 0x31, 0xa0, 0x29, 0xb8, 0x50, 0x45, 0xb6, 0x81, 0x81, 0x58, 0x5d,
                                   0x28, 0x33, 0xe8, 0x48, 0x79, 0xb9, 0x70, 0x91, 0x43, 0xe1, 0xf5, 0x93, 0xf0, 0x00, 0x00, 0x01};
  sha3_256(reinterpret_cast<const uint8_t*>(s.data()), s.size(), out);
  while (memcmp(out, R_BE, 32) >= 0) {
    int borrow = 0;
    for (int i = 31; i >= 0; --i) { int d = (int)out[i] - (int)R_BE[i] - borrow; borrow = d < 0; out[i] = (uint8_t)(d + (borrow ? 256 : 0)); }
  }
}

}  // namespace rbh
