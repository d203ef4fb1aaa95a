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

// ------------------------------------------------------------------------------------------ Fq6
struct Fp6 { Fp2 c[3]; };

static RB_NOINLINE Fp6 fp6_mul_nv(Fp6 x, Fp6 y) {
  Fp2 t0 = fp2_mul(x.c[0], y.c[0]);
  Fp2 t1 = fp2_mul(x.c[1], y.c[1]);
  Fp2 t2 = fp2_mul(x.c[2], y.c[2]);
  Fp2 u0 = fp2_mul(fp2_add(x.c[1], x.c[2]), fp2_add(y.c[1], y.c[2]));
  Fp2 u1 = fp2_mul(fp2_add(x.c[0], x.c[1]), fp2_add(y.c[0], y.c[1]));
  Fp2 u2 = fp2_mul(fp2_add(x.c[0], x.c[2]), fp2_add(y.c[0], y.c[2]));
  Fp6 r;
  r.c[0] = fp2_add(fp2_mul_xi(fp2_sub(fp2_sub(u0, t1), t2)), t0);
  r.c[1] = fp2_add(fp2_sub(fp2_sub(u1, t0), t1), fp2_mul_xi(t2));
  r.c[2] = fp2_add(fp2_sub(fp2_sub(u2, t0), t2), t1);
  return r;
}
RB_FN Fp6 fp6_add(const Fp6& x, const Fp6& y) { return {{fp2_add(x.c[0], y.c[0]), fp2_add(x.c[1], y.c[1]), fp2_add(x.c[2], y.c[2])}}; }
RB_FN Fp6 fp6_sub(const Fp6& x, const Fp6& y) { return {{fp2_sub(x.c[0], y.c[0]), fp2_sub(x.c[1], y.c[1]), fp2_sub(x.c[2], y.c[2])}}; }
RB_FN Fp6 fp6_neg(const Fp6& x) { return {{fp2_neg(x.c[0]), fp2_neg(x.c[1]), fp2_neg(x.c[2])}}; }
RB_FN Fp6 fp6_mul_v(const Fp6& x) { return {{fp2_mul_xi(x.c[2]), x.c[0], x.c[1]}}; }
RB_FN Fp6 fp6_mul(const Fp6& x, const Fp6& y) { return fp6_mul_nv(x, y); }

static RB_NOINLINE Fp6 fp6_inv_nv(Fp6 x) {
  Fp2 t0 = fp2_sub(fp2_sqr(x.c[0]), fp2_mul_xi(fp2_mul(x.c[1], x.c[2])));
  Fp2 t1 = fp2_sub(fp2_mul_xi(fp2_sqr(x.c[2])), fp2_mul(x.c[0], x.c[1]));
  Fp2 t2 = fp2_sub(fp2_sqr(x.c[1]), fp2_mul(x.c[0], x.c[2]));
  Fp2 d = fp2_add(fp2_mul(x.c[0], t0), fp2_mul_xi(fp2_add(fp2_mul(x.c[2], t1), fp2_mul(x.c[1], t2))));
  d = fp2_inv(d);
  Fp6 r;
  r.c[0] = fp2_mul(t0, d); r.c[1] = fp2_mul(t1, d); r.c[2] = fp2_mul(t2, d);
  return r;
}

// ------------------------------------------------------------------------------------------ Fq12
struct Fp12 { Fp6 h[2]; };        // h[0] + h[1] w

// coefficient k of the flat struct order (k = 0..5 -> c0.c0 c0.c1 c0.c2 c1.c0 c1.c1 c1.c2)
RB_FN Fp2& f12c(Fp12& x, int k) { return x.h[k / 3].c[k % 3]; }
RB_FN const Fp2& f12c(const Fp12& x, int k) { return x.h[k / 3].c[k % 3]; }

RB_FN void fp12_set_one(Fp12& r) {
  r.h[0].c[0] = fp2_one(); r.h[0].c[1] = fp2_zero(); r.h[0].c[2] = fp2_zero();
  r.h[1].c[0] = fp2_zero(); r.h[1].c[1] = fp2_zero(); r.h[1].c[2] = fp2_zero();
}

// All out-of-line Fq12 routines take and return values (see the note at fp2_mul_nv); the *_to
// forms below are thin inlined wrappers so that call sites may freely update in place.
static RB_NOINLINE Fp12 fp12_mul_nv(Fp12 x, Fp12 y) {
  Fp6 aa = fp6_mul(x.h[0], y.h[0]);
  Fp6 bb = fp6_mul(x.h[1], y.h[1]);
  Fp6 cr = fp6_mul(fp6_add(x.h[0], x.h[1]), fp6_add(y.h[0], y.h[1]));
  Fp12 r;
  r.h[0] = fp6_add(aa, fp6_mul_v(bb));
  r.h[1] = fp6_sub(fp6_sub(cr, aa), bb);
  return r;
}
static RB_NOINLINE Fp12 fp12_sqr_nv(Fp12 x) {
  Fp6 ab = fp6_mul(x.h[0], x.h[1]);
  Fp6 m = fp6_mul(fp6_add(x.h[0], x.h[1]), fp6_add(x.h[0], fp6_mul_v(x.h[1])));
  Fp12 r;
  r.h[0] = fp6_sub(fp6_sub(m, ab), fp6_mul_v(ab));
  r.h[1] = fp6_add(ab, ab);
  return r;
}
static RB_NOINLINE Fp12 fp12_inv_nv(Fp12 x) {
  Fp6 a2 = fp6_mul(x.h[0], x.h[0]), b2 = fp6_mul(x.h[1], x.h[1]);
  Fp6 di = fp6_inv_nv(fp6_sub(a2, fp6_mul_v(b2)));
  Fp12 r;
  r.h[0] = fp6_mul(x.h[0], di);
  r.h[1] = fp6_neg(fp6_mul(x.h[1], di));
  return r;
}
RB_FN void fp12_mul_to(Fp12* r, const Fp12* x, const Fp12* y) { *r = fp12_mul_nv(*x, *y); }
RB_FN void fp12_sqr_to(Fp12* r, const Fp12* x) { *r = fp12_sqr_nv(*x); }
RB_FN void fp12_inv_to(Fp12* r, const Fp12* x) { *r = fp12_inv_nv(*x); }
RB_FN void fp12_conj_to(Fp12* r, const Fp12* x) {
  Fp12 t; t.h[0] = x->h[0]; t.h[1] = fp6_neg(x->h[1]); *r = t;
}

// index of the Fq2 coefficient that multiplies w^k, k = 0..5
RB_FN constexpr int wk_index(int k) { return (k & 1) ? 3 + (k >> 1) : (k >> 1); }

// x -> x^(p^j), j = 1, 2, 3
static RB_NOINLINE Fp12 fp12_frobenius_nv(Fp12 x, int j) {
  const Fp2* g = (j == 1) ? FROB1 : ((j == 2) ? FROB2 : FROB3);
  Fp12 r;
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
  for (int k = 0; k < 6; ++k) {
    int idx = wk_index(k);
    Fp2 z = f12c(x, idx);
    if (j & 1) z = fp2_conj(z);
    Fp2 gk = g[k];
    f12c(r, idx) = fp2_mul(z, gk);
  }
  return r;
}
RB_FN void fp12_frobenius_to(Fp12* r, const Fp12* x, int j) { *r = fp12_frobenius_nv(*x, j); }

// f * (l0 + l3 w^3 + l4 w^4)  (value of a Miller line, see pairing.cuh); 15 Fq2 products.
static RB_NOINLINE Fp12 fp12_mul_by_line_nv(Fp12 f, Fp2 l0, Fp2 l3, Fp2 l4) {
  const Fp6& a = f.h[0]; const Fp6& b = f.h[1];
  Fp6 aa, bb, sum;
  aa.c[0] = fp2_add(fp2_mul(a.c[0], l0), fp2_mul_xi(fp2_mul(a.c[1], l4)));
  aa.c[1] = fp2_add(fp2_mul(a.c[1], l0), fp2_mul_xi(fp2_mul(a.c[2], l4)));
  aa.c[2] = fp2_add(fp2_mul(a.c[2], l0), fp2_mul(a.c[0], l4));
  bb.c[0] = fp2_mul_xi(fp2_mul(b.c[2], l3));
  bb.c[1] = fp2_mul(b.c[0], l3);
  bb.c[2] = fp2_mul(b.c[1], l3);
  sum.c[0] = l0; sum.c[1] = l3; sum.c[2] = l4;
  Fp6 cr = fp6_mul(fp6_add(a, b), sum);
  Fp12 r;
  r.h[0] = fp6_add(aa, fp6_mul_v(bb));
  r.h[1] = fp6_sub(fp6_sub(cr, aa), bb);
  return r;
}
RB_FN void fp12_mul_by_line(Fp12* f, const Fp2* l0, const Fp2* l3, const Fp2* l4) { *f = fp12_mul_by_line_nv(*f, *l0, *l3, *l4); }

// Granger-Scott squaring; only valid for elements of the cyclotomic subgroup.
static RB_NOINLINE Fp12 fp12_cyclotomic_sqr_nv(Fp12 x) {
  Fp2 z0 = f12c(x, 0), z4 = f12c(x, 1), z3 = f12c(x, 2), z2 = f12c(x, 3), z1 = f12c(x, 4), z5 = f12c(x, 5);
  Fp2 tmp, t0, t1, t2, t3, t4, t5;
  tmp = fp2_mul(z0, z1);
  t0 = fp2_sub(fp2_sub(fp2_mul(fp2_add(z0, z1), fp2_add(z0, fp2_mul_xi(z1))), tmp), fp2_mul_xi(tmp));
  t1 = fp2_dbl(tmp);
  tmp = fp2_mul(z2, z3);
  t2 = fp2_sub(fp2_sub(fp2_mul(fp2_add(z2, z3), fp2_add(z2, fp2_mul_xi(z3))), tmp), fp2_mul_xi(tmp));
  t3 = fp2_dbl(tmp);
  tmp = fp2_mul(z4, z5);
  t4 = fp2_sub(fp2_sub(fp2_mul(fp2_add(z4, z5), fp2_add(z4, fp2_mul_xi(z5))), tmp), fp2_mul_xi(tmp));
  t5 = fp2_dbl(tmp);
  Fp12 r;
  f12c(r, 0) = fp2_add(fp2_dbl(fp2_sub(t0, z0)), t0);
  f12c(r, 4) = fp2_add(fp2_dbl(fp2_add(t1, z1)), t1);
  tmp = fp2_mul_xi(t5);
  f12c(r, 3) = fp2_add(fp2_dbl(fp2_add(tmp, z2)), tmp);
  f12c(r, 2) = fp2_add(fp2_dbl(fp2_sub(t4, z3)), t4);
  f12c(r, 1) = fp2_add(fp2_dbl(fp2_sub(t2, z4)), t2);
  f12c(r, 5) = fp2_add(fp2_dbl(fp2_add(t3, z5)), t3);
  return r;
}
RB_FN void fp12_cyclotomic_sqr_to(Fp12* r, const Fp12* x) { *r = fp12_cyclotomic_sqr_nv(*x); }

// x^u for the BN parameter u = 4965661367192848881 (63 bits), x in the cyclotomic subgroup
static RB_NOINLINE Fp12 fp12_cyclotomic_exp_u_nv(Fp12 x) {
  const uint64_t u = 4965661367192848881ull;
  Fp12 acc = x;
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
  for (int i = 61; i >= 0; --i) {
    acc = fp12_cyclotomic_sqr_nv(acc);
    if ((u >> i) & 1) acc = fp12_mul_nv(acc, x);
  }
  return acc;
}
RB_FN void fp12_cyclotomic_exp_u_to(Fp12* r, const Fp12* x) { *r = fp12_cyclotomic_exp_u_nv(*x); }

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
