// TEST INFRASTRUCTURE ONLY -- CPU oracle for the rabe hot path (see oracle/README.md).
//
// A plain C++17 restatement of the BN254 arithmetic rabe obtains from the external crate
// `rabe-bn 0.4.23` (/root/reference/Cargo.toml:33; source NOT in /root/reference, a fork of the
// zcash `bn` crate per /root/reference/README.md:8).  Algorithms follow that lineage's published
// structure: 4x64-bit Montgomery Fq/Fr (R = 2^256), tower Fq2 = Fq[i]/(i^2+1),
// Fq6 = Fq2[v]/(v^3-(9+i)), Fq12 = Fq6[w]/(w^2-v), Jacobian G1/G2 with MSB-first double-and-add
// `G * Fr`, optimal-ate Miller loop over 6u+2 with homogeneous-projective twist formulas, and the
// final exponentiation whose hard part raises to LAMBDA = K*(p^4-p^2+1)/r.
//
// PARITY UNPINNED against rabe itself: no reference test asserts a group element, Gt value or
// serialized byte.  This oracle is pinned instead against oracle/pyref.py (an
// independent flat-Fp12/affine/textbook statement) and the algebraic known answers in
// tests/test_oracle_*.py.
//
// Nothing under rabe_b200/ may include this file.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace orc {

typedef uint64_t u64;
typedef unsigned __int128 u128;

// ---------------------------------------------------------------------------------------------
// operation counters (Fp-mul equivalents; the W_ref figures of BASELINE.md come from these)
struct OpCount { u64 fp_mul = 0; u64 fr_mul = 0; };
extern thread_local OpCount g_ops;

// ---------------------------------------------------------------------------------------------
// 256-bit helpers, constexpr so that the Montgomery constants are derived, not typed in.
struct U256 { u64 v[4]; };

constexpr bool u256_geq(const U256& a, const U256& b) {
  for (int i = 3; i >= 0; --i) { if (a.v[i] != b.v[i]) return a.v[i] > b.v[i]; }
  return true;
}
constexpr U256 u256_sub(const U256& a, const U256& b) {
  U256 r{}; u64 borrow = 0;
  for (int i = 0; i < 4; ++i) {
    u128 d = (u128)a.v[i] - b.v[i] - borrow;
    r.v[i] = (u64)d; borrow = (u64)(d >> 64) & 1;
  }
  return r;
}
// (2*a) mod n for a < n < 2^255
constexpr U256 u256_dbl_mod(const U256& a, const U256& n) {
  U256 r{};
  u64 carry = 0;
  for (int i = 0; i < 4; ++i) { r.v[i] = (a.v[i] << 1) | carry; carry = a.v[i] >> 63; }
  if (carry || u256_geq(r, n)) r = u256_sub(r, n);
  return r;
}
constexpr U256 pow2_mod(int e, const U256& n) {   // 2^e mod n
  U256 r{{1, 0, 0, 0}};
  for (int i = 0; i < e; ++i) r = u256_dbl_mod(r, n);
  return r;
}
constexpr u64 neg_inv64(u64 n0) {                 // -n^{-1} mod 2^64
  u64 x = 1;
  for (int i = 0; i < 6; ++i) x *= 2 - n0 * x;
  return ~x + 1;
}

// p = 36u^4+36u^3+24u^2+6u+1, r = 36u^4+36u^3+18u^2+6u+1, u = 4965661367192848881
constexpr U256 MOD_P{{0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull}};
constexpr U256 MOD_R{{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull}};
constexpr u64 BN_U = 4965661367192848881ull;

struct PTag { static constexpr U256 N = MOD_P; static constexpr bool is_fp = true; };
struct RTag { static constexpr U256 N = MOD_R; static constexpr bool is_fp = false; };

// ---------------------------------------------------------------------------------------------
// Montgomery field element mod T::N
template <class T>
struct Fm {
  u64 v[4];
  static constexpr U256 N = T::N;
  static constexpr U256 R1 = pow2_mod(256, T::N);
  static constexpr U256 R2 = pow2_mod(512, T::N);
  static constexpr u64 INV = neg_inv64(T::N.v[0]);

  static Fm zero() { Fm r; memset(r.v, 0, 32); return r; }
  static Fm one() { Fm r; memcpy(r.v, R1.v, 32); return r; }
  static Fm from_u64(u64 x) { U256 t{{x, 0, 0, 0}}; return from_u256(t); }
  // value must already be < N
  static Fm from_u256(const U256& a) { Fm t; memcpy(t.v, a.v, 32); Fm r2; memcpy(r2.v, R2.v, 32); return t * r2; }
  // any 256-bit value, reduced mod N (the lineage's `new_mul_factor` behaviour used by from_slice)
  static Fm from_u256_reduce(U256 a) { while (u256_geq(a, N)) a = u256_sub(a, N); return from_u256(a); }
  U256 to_u256() const {
    Fm o; memset(o.v, 0, 32); o.v[0] = 1;
    Fm r = mont_mul(*this, o, false);
    U256 out; memcpy(out.v, r.v, 32); return out;
  }
  // 32 bytes big-endian, canonical (non-Montgomery)
  void to_be(uint8_t* out) const {
    U256 a = to_u256();
    for (int i = 0; i < 4; ++i) for (int b = 0; b < 8; ++b) out[31 - (8 * i + b)] = (uint8_t)(a.v[i] >> (8 * b));
  }
  static U256 be_to_u256(const uint8_t* in) {
    U256 a{};
    for (int i = 0; i < 4; ++i) for (int b = 0; b < 8; ++b) a.v[i] |= (u64)in[31 - (8 * i + b)] << (8 * b);
    return a;
  }
  // returns false if the encoded integer is >= N
  static bool from_be(const uint8_t* in, Fm& out) {
    U256 a = be_to_u256(in);
    if (u256_geq(a, N)) return false;
    out = from_u256(a); return true;
  }
  static Fm from_be_reduce(const uint8_t* in) { return from_u256_reduce(be_to_u256(in)); }

  bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
  bool operator==(const Fm& o) const { return memcmp(v, o.v, 32) == 0; }
  bool operator!=(const Fm& o) const { return !(*this == o); }

  Fm operator+(const Fm& o) const {
    Fm r; u64 c = 0;
    for (int i = 0; i < 4; ++i) { u128 s = (u128)v[i] + o.v[i] + c; r.v[i] = (u64)s; c = (u64)(s >> 64); }
    U256 t; memcpy(t.v, r.v, 32);
    if (c || u256_geq(t, N)) { t = u256_sub(t, N); memcpy(r.v, t.v, 32); }
    return r;
  }
  Fm operator-(const Fm& o) const {
    Fm r; u64 b = 0;
    for (int i = 0; i < 4; ++i) { u128 d = (u128)v[i] - o.v[i] - b; r.v[i] = (u64)d; b = (u64)(d >> 64) & 1; }
    if (b) { u64 c = 0; for (int i = 0; i < 4; ++i) { u128 s = (u128)r.v[i] + N.v[i] + c; r.v[i] = (u64)s; c = (u64)(s >> 64); } }
    return r;
  }
  Fm neg() const { return zero() - *this; }
  Fm dbl() const { return *this + *this; }

  static Fm mont_mul(const Fm& a, const Fm& b, bool count = true) {
    if (count) { if (T::is_fp) g_ops.fp_mul++; else g_ops.fr_mul++; }
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
      u64 c = 0;
      for (int j = 0; j < 4; ++j) {
        u128 s = (u128)a.v[j] * b.v[i] + t[j] + c;
        t[j] = (u64)s; c = (u64)(s >> 64);
      }
      u128 s = (u128)t[4] + c; t[4] = (u64)s; t[5] = (u64)(s >> 64);
      u64 m = t[0] * INV;
      s = (u128)m * N.v[0] + t[0]; c = (u64)(s >> 64);
      for (int j = 1; j < 4; ++j) {
        s = (u128)m * N.v[j] + t[j] + c;
        t[j - 1] = (u64)s; c = (u64)(s >> 64);
      }
      s = (u128)t[4] + c; t[3] = (u64)s; t[4] = t[5] + (u64)(s >> 64);
    }
    U256 r{{t[0], t[1], t[2], t[3]}};
    if (t[4] || u256_geq(r, N)) r = u256_sub(r, N);
    Fm out; memcpy(out.v, r.v, 32); return out;
  }
  Fm operator*(const Fm& o) const { return mont_mul(*this, o); }
  Fm sqr() const { return mont_mul(*this, *this); }

  // square-and-multiply, MSB first, over all 256 bits of e (the lineage's generic pow)
  Fm pow(const U256& e) const {
    Fm acc = one();
    for (int i = 255; i >= 0; --i) {
      acc = acc.sqr();
      if ((e.v[i >> 6] >> (i & 63)) & 1) acc = acc * *this;
    }
    return acc;
  }
  Fm inverse() const {            // Fermat; undefined for zero (returns zero)
    U256 two{{2, 0, 0, 0}};
    return pow(u256_sub(N, two));
  }
};

typedef Fm<PTag> Fq;
typedef Fm<RTag> Fr;

// ---------------------------------------------------------------------------------------------
struct Fq2 {
  Fq a, b;                                    // a + b i
  static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  static Fq2 one() { return {Fq::one(), Fq::zero()}; }
  bool is_zero() const { return a.is_zero() && b.is_zero(); }
  bool operator==(const Fq2& o) const { return a == o.a && b == o.b; }
  bool operator!=(const Fq2& o) const { return !(*this == o); }
  Fq2 operator+(const Fq2& o) const { return {a + o.a, b + o.b}; }
  Fq2 operator-(const Fq2& o) const { return {a - o.a, b - o.b}; }
  Fq2 neg() const { return {a.neg(), b.neg()}; }
  Fq2 dbl() const { return {a.dbl(), b.dbl()}; }
  Fq2 conj() const { return {a, b.neg()}; }
  Fq2 operator*(const Fq2& o) const {         // Karatsuba, 3 Fq mul
    Fq aa = a * o.a, bb = b * o.b;
    Fq s = (a + b) * (o.a + o.b);
    return {aa - bb, s - aa - bb};
  }
  Fq2 scale(const Fq& k) const { return {a * k, b * k}; }
  Fq2 sqr() const {                           // complex squaring, 2 Fq mul
    Fq ab = a * b;
    return {(a + b) * (a - b), ab.dbl()};
  }
  Fq2 mul_xi() const {                        // * (9 + i)
    Fq a2 = a.dbl(), a4 = a2.dbl(), a8 = a4.dbl();
    Fq b2 = b.dbl(), b4 = b2.dbl(), b8 = b4.dbl();
    return {a8 + a - b, b8 + b + a};
  }
  Fq2 inverse() const {
    Fq n = (a.sqr() + b.sqr()).inverse();
    return {a * n, (b * n).neg()};
  }
  Fq2 pow(const U256& e) const {
    Fq2 acc = one();
    for (int i = 255; i >= 0; --i) {
      acc = acc.sqr();
      if ((e.v[i >> 6] >> (i & 63)) & 1) acc = acc * *this;
    }
    return acc;
  }
};

struct Fq6 {
  Fq2 c0, c1, c2;                             // c0 + c1 v + c2 v^2,  v^3 = xi
  static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
  static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
  bool operator==(const Fq6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
  Fq6 operator+(const Fq6& o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
  Fq6 operator-(const Fq6& o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
  Fq6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
  Fq6 mul_v() const { return {c2.mul_xi(), c0, c1}; }
  Fq6 operator*(const Fq6& o) const {         // Karatsuba, 6 Fq2 mul
    Fq2 t0 = c0 * o.c0, t1 = c1 * o.c1, t2 = c2 * o.c2;
    Fq2 r0 = ((c1 + c2) * (o.c1 + o.c2) - t1 - t2).mul_xi() + t0;
    Fq2 r1 = (c0 + c1) * (o.c0 + o.c1) - t0 - t1 + t2.mul_xi();
    Fq2 r2 = (c0 + c2) * (o.c0 + o.c2) - t0 - t2 + t1;
    return {r0, r1, r2};
  }
  Fq6 scale(const Fq2& k) const { return {c0 * k, c1 * k, c2 * k}; }
  Fq6 sqr() const { return *this * *this; }
  Fq6 inverse() const {
    Fq2 t0 = c0.sqr() - (c1 * c2).mul_xi();
    Fq2 t1 = c2.sqr().mul_xi() - c0 * c1;
    Fq2 t2 = c1.sqr() - c0 * c2;
    Fq2 d = (c0 * t0 + (c2 * t1).mul_xi() + (c1 * t2).mul_xi()).inverse();
    return {t0 * d, t1 * d, t2 * d};
  }
};

struct FrobConsts {
  Fq2 g1[6], g2[6], g3[6];                    // xi^{k(p^j-1)/6}, k = 0..5, j = 1,2,3
  Fq2 tw_x1, tw_y1, tw_x2, tw_y2;             // twist Frobenius factors
  Fq2 twist_b;                                // 3/xi
  Fq two_inv;
};
const FrobConsts& frob();

struct Fq12 {
  Fq6 c0, c1;                                 // c0 + c1 w, w^2 = v
  static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
  bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
  bool operator!=(const Fq12& o) const { return !(*this == o); }
  Fq12 operator*(const Fq12& o) const {       // 3 Fq6 mul
    Fq6 aa = c0 * o.c0, bb = c1 * o.c1;
    return {aa + bb.mul_v(), (c0 + c1) * (o.c0 + o.c1) - aa - bb};
  }
  Fq12 sqr() const {                          // complex squaring, 2 Fq6 mul
    Fq6 ab = c0 * c1;
    Fq6 t = (c0 + c1) * (c0 + c1.mul_v()) - ab - ab.mul_v();
    return {t, ab + ab};
  }
  Fq12 conj() const { return {c0, c1.neg()}; }
  Fq12 inverse() const {
    Fq6 d = (c0.sqr() - c1.sqr().mul_v()).inverse();
    return {c0 * d, (c1 * d).neg()};
  }
  // coefficient of w^k, k = 0..5  (w^0:c0.c0 w^1:c1.c0 w^2:c0.c1 w^3:c1.c1 w^4:c0.c2 w^5:c1.c2)
  Fq2& wk(int k) { Fq6& h = (k & 1) ? c1 : c0; int j = k >> 1; return j == 0 ? h.c0 : (j == 1 ? h.c1 : h.c2); }
  const Fq2& wk(int k) const { return const_cast<Fq12*>(this)->wk(k); }
  Fq12 frobenius(int j) const {               // x -> x^(p^j), j = 1,2,3
    const FrobConsts& F = frob();
    Fq12 r;
    for (int k = 0; k < 6; ++k) {
      Fq2 z = wk(k);
      if (j & 1) z = z.conj();
      r.wk(k) = z * (j == 1 ? F.g1[k] : (j == 2 ? F.g2[k] : F.g3[k]));
    }
    return r;
  }
  // multiplication by the sparse line value  l0 + l3 w^3 + l4 w^4
  Fq12 mul_by_line(const Fq2& l0, const Fq2& l3, const Fq2& l4) const {
    // s = s0 + s1 w with s0 = l0 + l4 v^2, s1 = l3 v; Karatsuba over Fq6 with sparse factors
    // (6 + 3 + 6 = 15 Fq2 products).
    const Fq6& a = c0; const Fq6& b = c1;
    Fq6 aa = {a.c0 * l0 + (a.c1 * l4).mul_xi(), a.c1 * l0 + (a.c2 * l4).mul_xi(), a.c2 * l0 + a.c0 * l4};
    Fq6 bb = {(b.c2 * l3).mul_xi(), b.c0 * l3, b.c1 * l3};
    Fq6 sum = {l0, l3, l4};
    Fq6 cross = (a + b) * sum - aa - bb;
    return {aa + bb.mul_v(), cross};
  }
  // Granger-Scott squaring, valid only in the cyclotomic subgroup
  Fq12 cyclotomic_sqr() const {
    Fq2 z0 = c0.c0, z4 = c0.c1, z3 = c0.c2, z2 = c1.c0, z1 = c1.c1, z5 = c1.c2;
    Fq2 tmp = z0 * z1;
    Fq2 t0 = (z0 + z1) * (z0 + z1.mul_xi()) - tmp - tmp.mul_xi();
    Fq2 t1 = tmp.dbl();
    tmp = z2 * z3;
    Fq2 t2 = (z2 + z3) * (z2 + z3.mul_xi()) - tmp - tmp.mul_xi();
    Fq2 t3 = tmp.dbl();
    tmp = z4 * z5;
    Fq2 t4 = (z4 + z5) * (z4 + z5.mul_xi()) - tmp - tmp.mul_xi();
    Fq2 t5 = tmp.dbl();
    z0 = (t0 - z0).dbl() + t0;
    z1 = (t1 + z1).dbl() + t1;
    tmp = t5.mul_xi();
    z2 = (tmp + z2).dbl() + tmp;
    z3 = (t4 - z3).dbl() + t4;
    z4 = (t2 - z4).dbl() + t2;
    z5 = (t3 + z5).dbl() + t3;
    Fq12 r; r.c0 = {z0, z4, z3}; r.c1 = {z2, z1, z5};
    return r;
  }
  Fq12 cyclotomic_exp_u() const {             // x^u, u = BN_U (63 bits)
    Fq12 acc = *this;
    for (int i = 61; i >= 0; --i) {
      acc = acc.cyclotomic_sqr();
      if ((BN_U >> i) & 1) acc = acc * *this;
    }
    return acc;
  }
  // generic square-and-multiply over the 256 bits of a canonical Fr (the lineage's Gt::pow)
  Fq12 pow(const U256& e) const {
    Fq12 acc = one();
    bool started = false;
    for (int i = 255; i >= 0; --i) {
      if (started) acc = acc.sqr();
      if ((e.v[i >> 6] >> (i & 63)) & 1) { acc = started ? acc * *this : *this; started = true; }
    }
    return acc;
  }
  void to_be(uint8_t* out) const {            // 12 x 32 bytes, struct order
    const Fq2* z[6] = {&c0.c0, &c0.c1, &c0.c2, &c1.c0, &c1.c1, &c1.c2};
    for (int k = 0; k < 6; ++k) { z[k]->a.to_be(out + 64 * k); z[k]->b.to_be(out + 64 * k + 32); }
  }
  static bool from_be(const uint8_t* in, Fq12& o) {
    Fq2* z[6] = {&o.c0.c0, &o.c0.c1, &o.c0.c2, &o.c1.c0, &o.c1.c1, &o.c1.c2};
    bool ok = true;
    for (int k = 0; k < 6; ++k) { ok &= Fq::from_be(in + 64 * k, z[k]->a); ok &= Fq::from_be(in + 64 * k + 32, z[k]->b); }
    return ok;
  }
};

// ---------------------------------------------------------------------------------------------
// Jacobian points over F (Fq for G1, Fq2 for G2); z == 0 encodes infinity.
template <class F>
struct Jac {
  F x, y, z;
  static Jac zero() { return {F::zero(), F::one(), F::zero()}; }
  bool is_zero() const { return z.is_zero(); }
  Jac neg() const { return {x, y.neg(), z}; }
  Jac dbl() const {
    if (is_zero()) return *this;
    F a = x.sqr(), b = y.sqr(), c = b.sqr();
    F d = ((x + b).sqr() - a - c).dbl();
    F e = a.dbl() + a;
    F f = e.sqr();
    F x3 = f - d.dbl();
    F eightc = c.dbl().dbl().dbl();
    F y3 = e * (d - x3) - eightc;
    F z3 = (y * z).dbl();
    return {x3, y3, z3};
  }
  Jac operator+(const Jac& o) const {         // full Jacobian addition (11M + 5S)
    if (is_zero()) return o;
    if (o.is_zero()) return *this;
    F z1z1 = z.sqr(), z2z2 = o.z.sqr();
    F u1 = x * z2z2, u2 = o.x * z1z1;
    F s1 = y * o.z * z2z2, s2 = o.y * z * z1z1;
    if (u1 == u2) {
      if (s1 == s2) return dbl();
      return zero();
    }
    F h = u2 - u1;
    F i = h.dbl().sqr();
    F j = h * i;
    F rr = (s2 - s1).dbl();
    F v = u1 * i;
    F x3 = rr.sqr() - j - v.dbl();
    F y3 = rr * (v - x3) - (s1 * j).dbl();
    F z3 = ((z + o.z).sqr() - z1z1 - z2z2) * h;
    return {x3, y3, z3};
  }
  Jac operator-(const Jac& o) const { return *this + o.neg(); }
  // `G * Fr` of the lineage: MSB-first double-and-add over the canonical scalar bits
  Jac mul(const Fr& k) const {
    U256 e = k.to_u256();
    Jac res = zero(); bool found = false;
    for (int i = 255; i >= 0; --i) {
      if (found) res = res.dbl();
      if ((e.v[i >> 6] >> (i & 63)) & 1) { found = true; res = res + *this; }
    }
    return res;
  }
  bool to_affine(F& ax, F& ay) const {        // false for infinity
    if (is_zero()) return false;
    F zi = z.inverse(), zi2 = zi.sqr();
    ax = x * zi2; ay = y * zi2 * zi;
    return true;
  }
  static Jac from_affine(const F& ax, const F& ay) { return {ax, ay, F::one()}; }
};
typedef Jac<Fq> G1;
typedef Jac<Fq2> G2;

G1 g1_generator();
G2 g2_generator();

// canonical encodings: affine, big-endian, non-Montgomery; infinity = all zero bytes.
void g1_to_bytes(const G1& p, uint8_t out[64]);
bool g1_from_bytes(const uint8_t in[64], G1& p);      // checks range + curve equation
void g2_to_bytes(const G2& p, uint8_t out[128]);      // x.re x.im y.re y.im
bool g2_from_bytes(const uint8_t in[128], G2& p);

// pairing() of the lineage: one Miller loop + one final exponentiation per call
Fq12 miller_loop(const G1& p, const G2& q);
Fq12 final_exponentiation(const Fq12& f);
Fq12 pairing(const G1& p, const G2& q);

// SHA3-256 (FIPS 202) and the two hash helpers of src/utils/hash/mod.rs
void sha3_256(const uint8_t* data, size_t len, uint8_t out[32]);
Fr sha3_hash_fr(const std::string& s);                              // hash/mod.rs:23-32
template <class G> G sha3_hash(const G& g, const std::string& s) {  // hash/mod.rs:10-20
  return g.mul(sha3_hash_fr(s));
}

}  // namespace orc
