// TEST INFRASTRUCTURE ONLY -- reference-sequence restatement of rabe's AC17 (FAME) CP-ABE.
//
// Follows /root/reference/src/schemes/ac17/mod.rs statement by statement (line numbers cited at
// each function), with the naive per-call algorithms of the zcash-bn lineage (bn254.hpp): every
// `G * Fr` is a 254-bit double-and-add, every `sha3_hash(g, s)` is SHA3 + such a scalar mul, every
// pairing() runs its own final exponentiation.  All randomness that the reference draws from
// `rand::thread_rng()` is an explicit input, in the order the reference draws it.
//
// Byte layouts (canonical encodings of bn254.hpp, fields in the order of the reference structs):
//   Ac17PublicKey  (ac17/mod.rs:62)  g[64] | h_a[3][128] | e_gh_ka[2][384]            = 1216 B
//   Ac17MasterKey  (ac17/mod.rs:72)  g[64] | h[128] | g_k[3][64] | a[2][32] | b[2][32] =  512 B
#include "bn254.hpp"

using namespace orc;

namespace {
Fr fr_in(const uint8_t* p) { return Fr::from_be_reduce(p); }
}

extern "C" {

// ac17/mod.rs:141-188.  rnd = [rho_g, rho_h, a0, b0, a1, b1, k0, k1, k2] (9 x 32 B); the reference
// draws g and h as random group elements, here g = G1gen*rho_g, h = G2gen*rho_h.
int orc_ac17_setup(const uint8_t* rnd, uint8_t* pk, uint8_t* msk) {
  G1 g = g1_generator().mul(fr_in(rnd));
  G2 h = g2_generator().mul(fr_in(rnd + 32));
  Fq12 e_gh = pairing(g, h);
  Fr a[2], b[2], k[3];
  for (int i = 0; i < 2; ++i) { a[i] = fr_in(rnd + 64 + 64 * i); b[i] = fr_in(rnd + 96 + 64 * i); }
  for (int i = 0; i < 3; ++i) k[i] = fr_in(rnd + 192 + 32 * i);
  G2 h_a[3];
  for (int i = 0; i < 2; ++i) h_a[i] = h.mul(a[i]);
  h_a[2] = h;
  G1 g_k[3];
  for (int i = 0; i < 3; ++i) g_k[i] = g.mul(k[i]);
  Fq12 e_gh_ka[2];
  for (int i = 0; i < 2; ++i) e_gh_ka[i] = e_gh.pow((k[i] * a[i] + k[2]).to_u256());
  g1_to_bytes(g, pk);
  for (int i = 0; i < 3; ++i) g2_to_bytes(h_a[i], pk + 64 + 128 * i);
  for (int i = 0; i < 2; ++i) e_gh_ka[i].to_be(pk + 448 + 384 * i);
  g1_to_bytes(g, msk);
  g2_to_bytes(h, msk + 64);
  for (int i = 0; i < 3; ++i) g1_to_bytes(g_k[i], msk + 192 + 64 * i);
  for (int i = 0; i < 2; ++i) { a[i].to_be(msk + 384 + 32 * i); b[i].to_be(msk + 448 + 32 * i); }
  return 0;
}

// ac17/mod.rs:191-264.  rnd = [r0, r1, sigma_attr[0..n), sigma] ((n+3) x 32 B).
// out: k_0[3][128], k[n][3][64], k_p[3][64]
int orc_ac17_cp_keygen(const uint8_t* msk, int n, const char* const* attrs, const uint8_t* rnd,
                       uint8_t* k_0, uint8_t* k, uint8_t* k_p) {
  if (n <= 0) return -1;                                              // "empty attributes!" :197
  G1 g; G2 h; G1 g_k[3]; Fr a[2], b[2];
  if (!g1_from_bytes(msk, g) || !g2_from_bytes(msk + 64, h)) return -2;
  for (int i = 0; i < 3; ++i) if (!g1_from_bytes(msk + 192 + 64 * i, g_k[i])) return -2;
  for (int i = 0; i < 2; ++i) { a[i] = fr_in(msk + 384 + 32 * i); b[i] = fr_in(msk + 448 + 32 * i); }
  Fr r[2], sum = Fr::zero();
  for (int i = 0; i < 2; ++i) { r[i] = fr_in(rnd + 32 * i); sum = sum + r[i]; }
  Fr br[3];
  for (int i = 0; i < 2; ++i) br[i] = b[i] * r[i];
  br[2] = sum;
  for (int i = 0; i < 3; ++i) g2_to_bytes(h.mul(br[i]), k_0 + 128 * i);
  for (int x = 0; x < n; ++x) {
    std::string attr(attrs[x]);
    Fr sigma_attr = fr_in(rnd + 64 + 32 * x);
    for (int t = 0; t < 2; ++t) {
      G1 prod = G1::zero();
      Fr a_t = a[t].inverse();
      for (int l = 0; l < 3; ++l) {
        std::string hs = attr + std::to_string(l) + std::to_string(t);
        prod = prod + sha3_hash(g, hs).mul(br[l] * a_t);
      }
      prod = prod + g.mul(sigma_attr * a_t);
      g1_to_bytes(prod, k + 192 * x + 64 * t);
    }
    g1_to_bytes(g.mul(sigma_attr.neg()), k + 192 * x + 128);
  }
  Fr sigma = fr_in(rnd + 64 + 32 * n);
  for (int t = 0; t < 2; ++t) {
    G1 prod = g_k[t];
    Fr a_t = a[t].inverse();
    for (int l = 0; l < 3; ++l) {
      std::string hs = std::string("01") + std::to_string(l) + std::to_string(t);
      prod = prod + sha3_hash(g, hs).mul(br[l] * a_t);
    }
    prod = prod + g.mul(sigma * a_t);
    g1_to_bytes(prod, k_p + 64 * t);
  }
  g1_to_bytes(g_k[2] + g.mul(sigma.neg()), k_p + 128);
  return 0;
}

// ac17/mod.rs:274-376 (the part after parse()/AbePolicy::from_policy, which the Python oracle
// layer oracle/policy.py restates).  m = n1 x n2 row-major i8, pi = n1 row labels.
// rnd = [s0, s1] ; msg = the reference's random Gt.  out: c_0[3][128], c[n1][3][64], c_p[384]
int orc_ac17_cp_encrypt(const uint8_t* pk, int n1, int n2, const int8_t* m, const char* const* pi,
                        const uint8_t* rnd, const uint8_t* msg_in, uint8_t* c_0, uint8_t* c, uint8_t* c_p) {
  G1 g; G2 h_a[3]; Fq12 e_gh_ka[2], msg;
  if (!g1_from_bytes(pk, g)) return -2;
  for (int i = 0; i < 3; ++i) if (!g2_from_bytes(pk + 64 + 128 * i, h_a[i])) return -2;
  for (int i = 0; i < 2; ++i) if (!Fq12::from_be(pk + 448 + 384 * i, e_gh_ka[i])) return -2;
  if (!Fq12::from_be(msg_in, msg)) return -2;
  Fr s[2], sum = Fr::zero();
  for (int i = 0; i < 2; ++i) { s[i] = fr_in(rnd + 32 * i); sum = sum + s[i]; }
  for (int i = 0; i < 2; ++i) g2_to_bytes(h_a[i].mul(s[i]), c_0 + 128 * i);
  g2_to_bytes(h_a[2].mul(sum), c_0 + 256);
  // pre-computed hashes :305-328
  std::vector<G1> table((size_t)n2 * 6);
  for (int j = 0; j < n2; ++j)
    for (int l = 0; l < 3; ++l)
      for (int t = 0; t < 2; ++t)
        table[(size_t)j * 6 + l * 2 + t] = sha3_hash(g, std::string("0") + std::to_string(j + 1) + std::to_string(l) + std::to_string(t));
  // :330-356
  for (int i = 0; i < n1; ++i) {
    for (int l = 0; l < 3; ++l) {
      G1 prod = G1::zero();
      for (int t = 0; t < 2; ++t) {
        G1 hash = sha3_hash(g, std::string(pi[i]) + std::to_string(l) + std::to_string(t));
        for (int j = 0; j < n2; ++j) {
          int8_t v = m[(size_t)i * n2 + j];
          if (v == 1) hash = hash + table[(size_t)j * 6 + l * 2 + t];
          else if (v == -1) hash = hash - table[(size_t)j * 6 + l * 2 + t];
        }
        prod = prod + hash.mul(s[t]);
      }
      g1_to_bytes(prod, c + 192 * (size_t)i + 64 * l);
    }
  }
  Fq12 cp = Fq12::one();
  for (int i = 0; i < 2; ++i) cp = cp * e_gh_ka[i].pow(s[i].to_u256());
  (cp * msg).to_be(c_p);
  return 0;
}

// ac17/mod.rs:385-430 (after parse/traverse_policy/calc_pruned).  `list` = the attribute names of
// the pruned set in calc_pruned order; matching against ct rows / key rows is by name, every
// match is added (duplicates included), exactly as the nested loops at :404-413 do.
// out: msg[384] = c_p * (prod2 * prod1^{-1})
int orc_ac17_cp_decrypt(int n_list, const char* const* list,
                        int n_ct, const char* const* ct_names, const uint8_t* c_0, const uint8_t* c, const uint8_t* c_p,
                        int n_sk, const char* const* sk_names, const uint8_t* k_0, const uint8_t* k, const uint8_t* k_p,
                        uint8_t* msg_out) {
  Fq12 prod1 = Fq12::one(), prod2 = Fq12::one(), cp;
  if (!Fq12::from_be(c_p, cp)) return -2;
  for (int i = 0; i < 3; ++i) {
    G1 prod_h = G1::zero(), prod_g = G1::zero();
    for (int cur = 0; cur < n_list; ++cur) {
      for (int x = 0; x < n_ct; ++x)
        if (strcmp(ct_names[x], list[cur]) == 0) { G1 p; if (!g1_from_bytes(c + 192 * (size_t)x + 64 * i, p)) return -2; prod_g = prod_g + p; }
      for (int x = 0; x < n_sk; ++x)
        if (strcmp(sk_names[x], list[cur]) == 0) { G1 p; if (!g1_from_bytes(k + 192 * (size_t)x + 64 * i, p)) return -2; prod_h = prod_h + p; }
    }
    G1 kp; G2 c0, k0;
    if (!g1_from_bytes(k_p + 64 * i, kp) || !g2_from_bytes(c_0 + 128 * i, c0) || !g2_from_bytes(k_0 + 128 * i, k0)) return -2;
    prod1 = prod1 * pairing(kp + prod_h, c0);
    prod2 = prod2 * pairing(prod_g, k0);
  }
  (cp * (prod2 * prod1.inverse())).to_be(msg_out);
  return 0;
}

}  // extern "C"
