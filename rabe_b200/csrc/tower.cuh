// Extension tower over Fq for BN254: Fq2 = Fq[i]/(i^2+1), Fq6 = Fq2[v]/(v^3 - xi), xi = 9+i,
// Fq12 = Fq6[w]/(w^2 - v).  Replaces rabe_bn's Fq2/Fq6/Fq12 behind `Gt` (rabe call sites:
// `Gt * Gt`, `Gt.pow`, `Gt.inverse` -- /root/reference/src/schemes/ac17/mod.rs:357-359,415-418,
// bsw/mod.rs:291-294,308, lsw/mod.rs:275-280, aw11/mod.rs:340-350).
//
// Device layout: an Fq12 value is six Fq2 coefficients in struct order
//   c[0]=c0.c0  c[1]=c0.c1  c[2]=c0.c2  c[3]=c1.c0  c[4]=c1.c1  c[5]=c1.c2
// which is also the canonical serialization order (12 x 32-byte big-endian Fq, re then im).
// Fq12 values live in per-thread local memory; the Fq2 product is the unit that runs in
// registers (one non-inlined routine shared by every caller keeps the instruction footprint small).
#pragma once
#include "fp.cuh"

namespace rb {

struct Fp2 { Fp a, b; };          // a + b i

#if defined(RB_HOST_SIM)
#define RB_CONST static const
#else
#define RB_CONST __constant__
#endif
}  // namespace rb
#include "consts_gen.cuh"
namespace rb {

RB_FN Fp2 fp2_zero() { return {fe_zero<ModP>(), fe_zero<ModP>()}; }
RB_FN Fp2 fp2_one() { return {fe_one<ModP>(), fe_zero<ModP>()}; }
RB_FN bool fp2_is_zero(const Fp2& x) { return fe_is_zero(x.a) && fe_is_zero(x.b); }
RB_FN bool fp2_eq(const Fp2& x, const Fp2& y) { return fe_eq(x.a, y.a) && fe_eq(x.b, y.b); }
RB_FN Fp2 fp2_add(const Fp2& x, const Fp2& y) { return {x.a + y.a, x.b + y.b}; }
RB_FN Fp2 fp2_sub(const Fp2& x, const Fp2& y) { return {x.a - y.a, x.b - y.b}; }
RB_FN Fp2 fp2_neg(const Fp2& x) { return {fe_neg(x.a), fe_neg(x.b)}; }
RB_FN Fp2 fp2_dbl(const Fp2& x) { return {fe_dbl(x.a), fe_dbl(x.b)}; }
RB_FN Fp2 fp2_half(const Fp2& x) { return {fe_half(x.a), fe_half(x.b)}; }
RB_FN Fp2 fp2_conj(const Fp2& x) { return {x.a, fe_neg(x.b)}; }
RB_FN Fp2 fp2_mul_fp(const Fp2& x, const Fp& k) { return {x.a * k, x.b * k}; }
RB_FN Fp2 fp2_mul_xi(const Fp2& x) {   // * (9 + i)
  Fp a2 = fe_dbl(x.a), a4 = fe_dbl(a2), a8 = fe_dbl(a4);
  Fp b2 = fe_dbl(x.b), b4 = fe_dbl(b2), b8 = fe_dbl(b4);
  return {a8 + x.a - x.b, b8 + x.b + x.a};
}

// Karatsuba product (3 Fq mul) and complex squaring (2 Fq mul); out-of-line on purpose.
// Operands travel BY VALUE: with pointer parameters nvcc 12.9 treats the result slot of the
// by-reference wrapper as non-aliasing, merges it with a dying operand slot in the caller and
// reorders the callee's loads past its stores -- wrong results in some inlined instances
// (found with the G2 variable-base kernel; tools/dbg/).  By-value parameters have no aliasing.
static RB_NOINLINE Fp2 fp2_mul_nv(Fp2 x, Fp2 y) {
  Fp aa = x.a * y.a, bb = x.b * y.b;
  Fp s = (x.a + x.b) * (y.a + y.b);
  return {aa - bb, s - aa - bb};
}
static RB_NOINLINE Fp2 fp2_sqr_nv(Fp2 x) {
  Fp ab = x.a * x.b;
  return {(x.a + x.b) * (x.a - x.b), fe_dbl(ab)};
}
static RB_NOINLINE Fp2 fp2_inv_nv(Fp2 x) {
  Fp n = fe_inv(fe_sqr(x.a) + fe_sqr(x.b));
  return {x.a * n, fe_neg(x.b * n)};
}
RB_FN Fp2 fp2_mul(const Fp2& x, const Fp2& y) { return fp2_mul_nv(x, y); }
RB_FN Fp2 fp2_sqr(const Fp2& x) { return fp2_sqr_nv(x); }
RB_FN Fp2 fp2_inv(const Fp2& x) { return fp2_inv_nv(x); }

// uniform names so curve code can be written once for Fq and Fq2
RB_FN Fp f_add(const Fp& a, const Fp& b) { return a + b; }
RB_FN Fp f_sub(const Fp& a, const Fp& b) { return a - b; }
RB_FN Fp f_mul(const Fp& a, const Fp& b) { return a * b; }
RB_FN Fp f_sqr(const Fp& a) { return fe_sqr(a); }
RB_FN Fp f_dbl(const Fp& a) { return fe_dbl(a); }
RB_FN Fp f_neg(const Fp& a) { return fe_neg(a); }
RB_FN bool f_is_zero(const Fp& a) { return fe_is_zero(a); }
RB_FN bool f_eq(const Fp& a, const Fp& b) { return fe_eq(a, b); }
RB_FN void f_set_zero(Fp& a) { a = fe_zero<ModP>(); }
RB_FN void f_set_one(Fp& a) { a = fe_one<ModP>(); }
RB_FN Fp2 f_add(const Fp2& a, const Fp2& b) { return fp2_add(a, b); }
RB_FN Fp2 f_sub(const Fp2& a, const Fp2& b) { return fp2_sub(a, b); }
RB_FN Fp2 f_mul(const Fp2& a, const Fp2& b) { return fp2_mul(a, b); }
RB_FN Fp2 f_sqr(const Fp2& a) { return fp2_sqr(a); }
RB_FN Fp2 f_dbl(const Fp2& a) { return fp2_dbl(a); }
RB_FN Fp2 f_neg(const Fp2& a) { return fp2_neg(a); }
RB_FN bool f_is_zero(const Fp2& a) { return fp2_is_zero(a); }
RB_FN bool f_eq(const Fp2& a, const Fp2& b) { return fp2_eq(a, b); }
RB_FN void f_set_zero(Fp2& a) { a = fp2_zero(); }
RB_FN void f_set_one(Fp2& a) { a = fp2_one(); }

typedef Fp2 FullFp2;
#define RB_K2(c) (c)
#define RB_KL(c) (c)
#define RB_FP2_MUL_HOT(x, y) fp2_mul(x, y)

#include "tower_body.inc"

// canonical bytes <-> Montgomery limbs
RB_FN void fp12_load_be(Fp12& r, const uint8_t* p) {
  RB_UNROLL for (int k = 0; k < 6; ++k) {
    f12c(r, k).a = fe_to_mont(fe_load_be<ModP>(p + 64 * k));
    f12c(r, k).b = fe_to_mont(fe_load_be<ModP>(p + 64 * k + 32));
  }
}
RB_FN void fp12_store_be(uint8_t* p, const Fp12& x) {
  RB_UNROLL for (int k = 0; k < 6; ++k) {
    fe_store_be(p + 64 * k, fe_from_mont(f12c(x, k).a));
    fe_store_be(p + 64 * k + 32, fe_from_mont(f12c(x, k).b));
  }
}

}  // namespace rb
