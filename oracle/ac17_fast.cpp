// TEST INFRASTRUCTURE ONLY -- the RESTRUCTURED AC17 algorithm on the CPU: the same algebra the CUDA path uses
// (DESIGN.md section 3), restated over the oracle's own arithmetic (bn254.hpp), so that bench.py can split the
// GPU-vs-reference ratio into "algorithm" (this file vs ac17.cpp on the same cores) and "hardware" (GPU vs this file):
//
//   cp_encrypt   every sha3_hash(g, s) is g * H(s) (hash/mod.rs:10-20), so the scalars are folded in Fr first,
//                A[i][l][t] = H(pi_i l t) + sum_j M_ij H("0" (j+1) l t),  and  c[i][l] = g * (s0 A[i][l][0] + s1 A[i][l][1])
//                is ONE fixed-base multiplication per output (8-bit windows: 32 additions) instead of four 254-bit
//                double-and-adds plus n2 point additions;  c_0 / c_p likewise from window tables (ac17/mod.rs:297-368).
//   cp_decrypt   six Miller loops, ONE final exponentiation (the reference runs one per pairing, ac17/mod.rs:415-416).
//
// Outputs are byte-identical to ac17.cpp's (tests/test_oracle_fast.py).  Never linked into the product.
#include "bn254.hpp"
#include <vector>

using namespace orc;

namespace {
Fr fr_in(const uint8_t* p) { return Fr::from_be_reduce(p); }

template <class G> struct WindowTable {          // tab[w][d] = d * 2^(8w) * base, d < 256 (Jacobian, not normalised)
  std::vector<G> t;
  void build(const G& base) {
    t.assign(32 * 256, G::zero());
    G wb = base;
    for (int w = 0; w < 32; ++w) {
      G acc = G::zero();
      for (int d = 1; d < 256; ++d) { acc = acc + wb; t[w * 256 + d] = acc; }
      for (int i = 0; i < 8; ++i) wb = wb.dbl();
    }
  }
  G mul(const Fr& k) const {
    U256 e = k.to_u256();
    G acc = G::zero();
    for (int w = 0; w < 32; ++w) {
      unsigned d = (unsigned)(e.v[w >> 3] >> (8 * (w & 7))) & 0xffu;
      if (d) acc = acc + t[w * 256 + d];
    }
    return acc;
  }
};
struct GtTable {
  std::vector<Fq12> t;
  void build(const Fq12& base) {
    t.assign(32 * 256, Fq12::one());
    Fq12 wb = base;
    for (int w = 0; w < 32; ++w) {
      Fq12 acc = Fq12::one();
      for (int d = 1; d < 256; ++d) { acc = acc * wb; t[w * 256 + d] = acc; }
      for (int i = 0; i < 8; ++i) wb = wb * wb;
    }
  }
  Fq12 pow(const Fr& k) const {
    U256 e = k.to_u256();
    Fq12 acc = Fq12::one();
    for (int w = 0; w < 32; ++w) {
      unsigned d = (unsigned)(e.v[w >> 3] >> (8 * (w & 7))) & 0xffu;
      if (d) acc = acc * t[w * 256 + d];
    }
    return acc;
  }
};
struct FastPk { WindowTable<G1> g; WindowTable<G2> h_a[3]; GtTable e[2]; };
struct FastMsp { int n1; std::vector<Fr> A; };     // A[i][l][t]
}  // namespace

extern "C" {

// one-off per public key (the GPU path builds its tables in rb_ac17_pk_load)
void* orc_ac17_fast_pk_new(const uint8_t* pk) {
  G1 g; G2 h; Fq12 e;
  if (!g1_from_bytes(pk, g)) return nullptr;
  FastPk* p = new FastPk();
  p->g.build(g);
  for (int i = 0; i < 3; ++i) { if (!g2_from_bytes(pk + 64 + 128 * i, h)) { delete p; return nullptr; } p->h_a[i].build(h); }
  for (int i = 0; i < 2; ++i) { if (!Fq12::from_be(pk + 448 + 384 * i, e)) { delete p; return nullptr; } p->e[i].build(e); }
  return p;
}
void orc_ac17_fast_pk_free(void* p) { delete static_cast<FastPk*>(p); }

// one-off per policy (rb_msp_load): the folded scalar table
void* orc_ac17_fast_msp_new(int n1, int n2, const int8_t* m, const char* const* pi) {
  FastMsp* f = new FastMsp();
  f->n1 = n1; f->A.resize((size_t)n1 * 6);
  std::vector<Fr> col((size_t)n2 * 6);
  for (int j = 0; j < n2; ++j)
    for (int l = 0; l < 3; ++l)
      for (int t = 0; t < 2; ++t)
        col[(size_t)j * 6 + l * 2 + t] = sha3_hash_fr(std::string("0") + std::to_string(j + 1) + std::to_string(l) + std::to_string(t));
  for (int i = 0; i < n1; ++i)
    for (int l = 0; l < 3; ++l)
      for (int t = 0; t < 2; ++t) {
        Fr a = sha3_hash_fr(std::string(pi[i]) + std::to_string(l) + std::to_string(t));
        for (int j = 0; j < n2; ++j) {
          int8_t v = m[(size_t)i * n2 + j];
          if (v == 1) a = a + col[(size_t)j * 6 + l * 2 + t];
          else if (v == -1) a = a - col[(size_t)j * 6 + l * 2 + t];
        }
        f->A[(size_t)i * 6 + l * 2 + t] = a;
      }
  return f;
}
void orc_ac17_fast_msp_free(void* p) { delete static_cast<FastMsp*>(p); }

int orc_ac17_cp_encrypt_fast(const void* pkh, const void* msph, const uint8_t* rnd, const uint8_t* msg_in, uint8_t* c_0, uint8_t* c, uint8_t* c_p) {
  const FastPk* pk = static_cast<const FastPk*>(pkh); const FastMsp* msp = static_cast<const FastMsp*>(msph);
  Fq12 msg; if (!pk || !msp || !Fq12::from_be(msg_in, msg)) return -2;
  Fr s[2] = {fr_in(rnd), fr_in(rnd + 32)};
  g2_to_bytes(pk->h_a[0].mul(s[0]), c_0); g2_to_bytes(pk->h_a[1].mul(s[1]), c_0 + 128); g2_to_bytes(pk->h_a[2].mul(s[0] + s[1]), c_0 + 256);
  for (int i = 0; i < msp->n1; ++i)
    for (int l = 0; l < 3; ++l)
      g1_to_bytes(pk->g.mul(s[0] * msp->A[(size_t)i * 6 + l * 2] + s[1] * msp->A[(size_t)i * 6 + l * 2 + 1]), c + 192 * (size_t)i + 64 * l);
  (pk->e[0].pow(s[0]) * pk->e[1].pow(s[1]) * msg).to_be(c_p);
  return 0;
}

// gather lists as indices (what the name matching of ac17/mod.rs:404-413 selects); one final exponentiation
int orc_ac17_cp_decrypt_fast(int n_ct_idx, const uint32_t* ct_idx, int n_sk_idx, const uint32_t* sk_idx, const uint8_t* c_0, const uint8_t* c,
                             const uint8_t* c_p, const uint8_t* k_0, const uint8_t* k, const uint8_t* k_p, uint8_t* msg_out) {
  Fq12 f = Fq12::one(), cp;
  if (!Fq12::from_be(c_p, cp)) return -2;
  for (int i = 0; i < 3; ++i) {
    G1 prod_h = G1::zero(), prod_g = G1::zero(), p;
    for (int x = 0; x < n_ct_idx; ++x) { if (!g1_from_bytes(c + 192 * (size_t)ct_idx[x] + 64 * i, p)) return -2; prod_g = prod_g + p; }
    for (int x = 0; x < n_sk_idx; ++x) { if (!g1_from_bytes(k + 192 * (size_t)sk_idx[x] + 64 * i, p)) return -2; prod_h = prod_h + p; }
    G1 kp; G2 c0, k0;
    if (!g1_from_bytes(k_p + 64 * i, kp) || !g2_from_bytes(c_0 + 128 * i, c0) || !g2_from_bytes(k_0 + 128 * i, k0)) return -2;
    G1 a = (kp + prod_h).neg();
    if (!a.is_zero() && !c0.is_zero()) f = f * miller_loop(a, c0);
    if (!prod_g.is_zero() && !k0.is_zero()) f = f * miller_loop(prod_g, k0);
  }
  (cp * final_exponentiation(f)).to_be(msg_out);
  return 0;
}

}  // extern "C"
