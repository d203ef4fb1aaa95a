// TEST INFRASTRUCTURE ONLY -- C entry points of the oracle's primitives (ctypes-loaded by
// oracle/__init__.py).  One call = one rabe_bn operator call at the sites listed in SURVEY.md 2
// ("L1 operator call-site inventory"); encodings are the canonical ones of bn254.hpp.
#include "bn254.hpp"

using namespace orc;

extern "C" {

void orc_ops_reset() { g_ops = OpCount(); }
void orc_ops_get(uint64_t out[2]) { out[0] = g_ops.fp_mul; out[1] = g_ops.fr_mul; }

// Montgomery constants, little-endian 64-bit words: [p, R mod p, R^2 mod p, -p^-1, r, R mod r, R^2 mod r, -r^-1]
void orc_constants(uint64_t out[26]) {
  memcpy(out, Fq::N.v, 32); memcpy(out + 4, Fq::R1.v, 32); memcpy(out + 8, Fq::R2.v, 32); out[12] = Fq::INV;
  memcpy(out + 13, Fr::N.v, 32); memcpy(out + 17, Fr::R1.v, 32); memcpy(out + 21, Fr::R2.v, 32); out[25] = Fr::INV;
}

void orc_g1_generator(uint8_t out[64]) { g1_to_bytes(g1_generator(), out); }
void orc_g2_generator(uint8_t out[128]) { g2_to_bytes(g2_generator(), out); }

int orc_g1_check(const uint8_t in[64]) { G1 p; return g1_from_bytes(in, p) ? 0 : -2; }
// Decoding a G2 value in the zcash-bn lineage (`AffineG2::new`): the curve equation and then the order check
// `(p * (-Fr::one())) + p == zero`, i.e. [r]p == O.  rabe surfaces a failure as FieldError::NotMember -> RabeError
// (/root/reference/src/error.rs:60-69).  orc_g2_on_curve is the first half alone.
int orc_g2_on_curve(const uint8_t in[128]) { G2 p; return g2_from_bytes(in, p) ? 0 : -2; }
int orc_g2_check(const uint8_t in[128]) {
  G2 p; if (!g2_from_bytes(in, p)) return -2;
  if (p.is_zero()) return 0;
  return (p.mul(Fr::one().neg()) + p).is_zero() ? 0 : -2;
}

int orc_g1_mul(const uint8_t base[64], const uint8_t k[32], uint8_t out[64]) {
  G1 p; if (!g1_from_bytes(base, p)) return -2;
  g1_to_bytes(p.mul(Fr::from_be_reduce(k)), out); return 0;
}
int orc_g2_mul(const uint8_t base[128], const uint8_t k[32], uint8_t out[128]) {
  G2 p; if (!g2_from_bytes(base, p)) return -2;
  g2_to_bytes(p.mul(Fr::from_be_reduce(k)), out); return 0;
}
int orc_g1_add(const uint8_t a[64], const uint8_t b[64], uint8_t out[64]) {
  G1 p, q; if (!g1_from_bytes(a, p) || !g1_from_bytes(b, q)) return -2;
  g1_to_bytes(p + q, out); return 0;
}
int orc_g2_add(const uint8_t a[128], const uint8_t b[128], uint8_t out[128]) {
  G2 p, q; if (!g2_from_bytes(a, p) || !g2_from_bytes(b, q)) return -2;
  g2_to_bytes(p + q, out); return 0;
}
int orc_g1_neg(const uint8_t a[64], uint8_t out[64]) {
  G1 p; if (!g1_from_bytes(a, p)) return -2;
  g1_to_bytes(p.neg(), out); return 0;
}
int orc_pairing(const uint8_t p1[64], const uint8_t q2[128], uint8_t out[384]) {
  G1 p; G2 q; if (!g1_from_bytes(p1, p) || !g2_from_bytes(q2, q)) return -2;
  pairing(p, q).to_be(out); return 0;
}
// miller loop only / final exponentiation only (for kernel-level parity of intermediate stages,
// compared after final exponentiation because Miller values are not canonical)
int orc_final_exp(const uint8_t in[384], uint8_t out[384]) {
  Fq12 f; if (!Fq12::from_be(in, f)) return -2;
  final_exponentiation(f).to_be(out); return 0;
}
int orc_gt_pow(const uint8_t a[384], const uint8_t k[32], uint8_t out[384]) {
  Fq12 f; if (!Fq12::from_be(a, f)) return -2;
  f.pow(Fr::from_be_reduce(k).to_u256()).to_be(out); return 0;
}
int orc_gt_mul(const uint8_t a[384], const uint8_t b[384], uint8_t out[384]) {
  Fq12 f, g; if (!Fq12::from_be(a, f) || !Fq12::from_be(b, g)) return -2;
  (f * g).to_be(out); return 0;
}
int orc_gt_inverse(const uint8_t a[384], uint8_t out[384]) {
  Fq12 f; if (!Fq12::from_be(a, f)) return -2;
  f.inverse().to_be(out); return 0;
}
int orc_gt_cyclotomic_sqr(const uint8_t a[384], uint8_t out[384]) {
  Fq12 f; if (!Fq12::from_be(a, f)) return -2;
  f.cyclotomic_sqr().to_be(out); return 0;
}
int orc_gt_frobenius(const uint8_t a[384], int j, uint8_t out[384]) {
  Fq12 f; if (!Fq12::from_be(a, f)) return -2;
  f.frobenius(j).to_be(out); return 0;
}
// Fr helpers: op 0 add, 1 sub, 2 mul, 3 inverse(a), 4 neg(a), 5 pow(a, b)
void orc_fr_op(int op, const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  Fr x = Fr::from_be_reduce(a), y = b ? Fr::from_be_reduce(b) : Fr::zero(), r;
  switch (op) {
    case 0: r = x + y; break;
    case 1: r = x - y; break;
    case 2: r = x * y; break;
    case 3: r = x.inverse(); break;
    case 4: r = x.neg(); break;
    default: r = x.pow(y.to_u256()); break;
  }
  r.to_be(out);
}
void orc_fq_op(int op, const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  Fq x = Fq::from_be_reduce(a), y = b ? Fq::from_be_reduce(b) : Fq::zero(), r;
  switch (op) {
    case 0: r = x + y; break;
    case 1: r = x - y; break;
    case 2: r = x * y; break;
    case 3: r = x.inverse(); break;
    default: r = x.neg(); break;
  }
  r.to_be(out);
}
void orc_sha3_256(const uint8_t* data, size_t len, uint8_t out[32]) { sha3_256(data, len, out); }
void orc_sha3_fr(const char* s, size_t len, uint8_t out[32]) { sha3_hash_fr(std::string(s, len)).to_be(out); }

}  // extern "C"
