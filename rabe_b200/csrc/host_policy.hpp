// Host-side policy layer of rabe_b200 (C++17, strings only -- no field or group arithmetic).
// Mirrors the reference's L2 helpers so that the numeric C ABI can be driven from policy text:
//   parse               /root/reference/src/utils/policy/pest/mod.rs:40-66 with the PEG grammars
//                       src/human.policy.pest, src/json.policy.pest and the tree builders
//                       pest/human.rs:8-49, pest/json.rs:8-48 (leaf = text + pest column)
//   calculate_msp / lw  src/utils/policy/msp.rs:78-147
//   calc_pruned         src/utils/secretsharing/mod.rs:143-201
//   traverse_policy     src/utils/tools/mod.rs:31-61
//   node_index / remove_index   src/utils/secretsharing/mod.rs:74-80
//   SHA3-256 -> Fr      src/utils/hash/mod.rs:23-32 (big-endian digest reduced mod r)
// rabe panics on malformed trees (msp.rs:121,133; secretsharing:167,187); here every such case is
// an error return.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace rbh {

enum Kind { LEAF = 0, AND = 1, OR = 2 };
enum Lang { LANG_JSON = 0, LANG_HUMAN = 1 };   // order of rabe's PolicyLanguage enum (pest/mod.rs:17-22)

struct Node {
  Kind kind = LEAF;
  std::string name;      // leaf text (inner of the quoted string)
  size_t col = 0;        // pest line_col().1 of the inner text: 1-based, in characters, per line
  std::vector<Node> kids;
};

struct Msp {
  std::vector<std::vector<int8_t>> m;   // n1 rows x c columns, entries in {-1,0,1}
  std::vector<std::string> pi;          // row labels, sorted (stable) by byte order
  size_t c = 1;
};

bool parse(const std::string& text, int lang, Node& out, std::string& err);
std::string serialize(const Node& n, int lang);
bool calculate_msp(const Node& root, Msp& out, std::string& err);
bool traverse_policy(const std::vector<std::string>& attrs, const Node& n);
// returns false on a malformed tree; `match` and `list` as rabe's (bool, Vec<(name, node_index)>)
bool calc_pruned(const std::vector<std::string>& attrs, const Node& n, bool& match,
                 std::vector<std::pair<std::string, std::string>>& list, std::string& err);
std::string node_index(const Node& leaf);
std::string remove_index(const std::string& label);
void leaves_dfs(const Node& n, std::vector<const Node*>& out);

void sha3_256(const uint8_t* data, size_t len, uint8_t out[32]);
void hash_to_fr(const std::string& s, uint8_t out[32]);   // canonical big-endian, < r

}  // namespace rbh
