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

// ------------------------------------------------------------------------------------------ Fq6 / Fq12
// Calling convention for everything wider than Fq2 (a hard-won rule for nvcc 12.9, see the note
// at fp2_mul_nv and tools/dbg/): values live in storage OWNED BY THE CALLER (named locals of the
// kernel or of an out-of-line routine), routines take plain pointers, the result pointer may alias
// an operand, and nothing wider than Fq2 is ever returned by value or passed by value.
struct Fp6 { Fp2 c[3]; };         // c0 + c1 v + c2 v^2
struct Fp12 { Fp6 h[2]; };        // h0 + h1 w

// coefficient k of the flat struct order (k = 0..5 -> c0.c0 c0.c1 c0.c2 c1.c0 c1.c1 c1.c2)
RB_FN Fp2& f12c(Fp12& x, int k) { return x.h[k / 3].c[k % 3]; }
RB_FN const Fp2& f12c(const Fp12& x, int k) { return x.h[k / 3].c[k % 3]; }

RB_FN void fp6_add_p(Fp6* r, const Fp6* x, const Fp6* y) { RB_UNROLL for (int k = 0; k < 3; ++k) r->c[k] = fp2_add(x->c[k], y->c[k]); }
RB_FN void fp6_sub_p(Fp6* r, const Fp6* x, const Fp6* y) { RB_UNROLL for (int k = 0; k < 3; ++k) r->c[k] = fp2_sub(x->c[k], y->c[k]); }
RB_FN void fp6_neg_p(Fp6* r, const Fp6* x) { RB_UNROLL for (int k = 0; k < 3; ++k) r->c[k] = fp2_neg(x->c[k]); }
RB_FN void fp6_mul_v_p(Fp6* r, const Fp6* x) {        // * v ; r may alias x
  Fp2 t = fp2_mul_xi(x->c[2]), c0 = x->c[0], c1 = x->c[1];
  r->c[0] = t; r->c[1] = c0; r->c[2] = c1;
}

// Karatsuba, 6 Fq2 products; r may alias x or y (all reads happen before the first write)
static RB_NOINLINE void fp6_mul_p(Fp6* r, const Fp6* x, const Fp6* y) {
  Fp2 x0 = x->c[0], x1 = x->c[1], x2 = x->c[2], y0 = y->c[0], y1 = y->c[1], y2 = y->c[2];
  Fp2 t0 = fp2_mul(x0, y0);
  Fp2 t1 = fp2_mul(x1, y1);
  Fp2 t2 = fp2_mul(x2, y2);
  Fp2 u0 = fp2_mul(fp2_add(x1, x2), fp2_add(y1, y2));
  Fp2 u1 = fp2_mul(fp2_add(x0, x1), fp2_add(y0, y1));
  Fp2 u2 = fp2_mul(fp2_add(x0, x2), fp2_add(y0, y2));
  r->c[0] = fp2_add(fp2_mul_xi(fp2_sub(fp2_sub(u0, t1), t2)), t0);
  r->c[1] = fp2_add(fp2_sub(fp2_sub(u1, t0), t1), fp2_mul_xi(t2));
  r->c[2] = fp2_add(fp2_sub(fp2_sub(u2, t0), t2), t1);
}
static RB_NOINLINE void fp6_inv_p(Fp6* r, const Fp6* x) {
  Fp2 x0 = x->c[0], x1 = x->c[1], x2 = x->c[2];
  Fp2 t0 = fp2_sub(fp2_sqr(x0), fp2_mul_xi(fp2_mul(x1, x2)));
  Fp2 t1 = fp2_sub(fp2_mul_xi(fp2_sqr(x2)), fp2_mul(x0, x1));
  Fp2 t2 = fp2_sub(fp2_sqr(x1), fp2_mul(x0, x2));
  Fp2 d = fp2_add(fp2_mul(x0, t0), fp2_mul_xi(fp2_add(fp2_mul(x2, t1), fp2_mul(x1, t2))));
  d = fp2_inv(d);
  r->c[0] = fp2_mul(t0, d); r->c[1] = fp2_mul(t1, d); r->c[2] = fp2_mul(t2, d);
}

RB_FN void fp12_set_one(Fp12& r) {
  r.h[0].c[0] = fp2_one(); r.h[0].c[1] = fp2_zero(); r.h[0].c[2] = fp2_zero();
  r.h[1].c[0] = fp2_zero(); r.h[1].c[1] = fp2_zero(); r.h[1].c[2] = fp2_zero();
}
RB_FN void fp12_copy(Fp12* r, const Fp12* x) {
  RB_UNROLL for (int k = 0; k < 6; ++k) f12c(*r, k) = f12c(*x, k);
}

// r = x * y  (3 Fq6 products); r may alias x or y
static RB_NOINLINE void fp12_mul_to(Fp12* r, const Fp12* x, const Fp12* y) {
  Fp6 aa, bb, cr;                        // cr doubles as the y-sum
  fp6_mul_p(&aa, &x->h[0], &y->h[0]);
  fp6_mul_p(&bb, &x->h[1], &y->h[1]);
  Fp6 sx;
  fp6_add_p(&sx, &x->h[0], &x->h[1]);
  fp6_add_p(&cr, &y->h[0], &y->h[1]);
  fp6_mul_p(&cr, &sx, &cr);
  fp6_sub_p(&cr, &cr, &aa);
  fp6_sub_p(&r->h[1], &cr, &bb);
  fp6_mul_v_p(&bb, &bb);
  fp6_add_p(&r->h[0], &aa, &bb);
}
// r = x^2  (complex squaring, 2 Fq6 products); r may alias x
static RB_NOINLINE void fp12_sqr_to(Fp12* r, const Fp12* x) {
  Fp6 ab, s, m;
  fp6_mul_p(&ab, &x->h[0], &x->h[1]);
  fp6_add_p(&s, &x->h[0], &x->h[1]);
  fp6_mul_v_p(&m, &x->h[1]);
  fp6_add_p(&m, &x->h[0], &m);
  fp6_mul_p(&m, &s, &m);
  fp6_sub_p(&m, &m, &ab);
  fp6_add_p(&r->h[1], &ab, &ab);
  fp6_mul_v_p(&ab, &ab);
  fp6_sub_p(&r->h[0], &m, &ab);
}
RB_FN void fp12_conj_to(Fp12* r, const Fp12* x) {
  RB_UNROLL for (int k = 0; k < 3; ++k) { r->h[0].c[k] = x->h[0].c[k]; r->h[1].c[k] = fp2_neg(x->h[1].c[k]); }
}
static RB_NOINLINE void fp12_inv_to(Fp12* r, const Fp12* x) {
  Fp6 a2, b2, d;
  fp6_mul_p(&a2, &x->h[0], &x->h[0]);
  fp6_mul_p(&b2, &x->h[1], &x->h[1]);
  fp6_mul_v_p(&b2, &b2);
  fp6_sub_p(&d, &a2, &b2);
  fp6_inv_p(&d, &d);
  fp6_mul_p(&a2, &x->h[1], &d);
  fp6_mul_p(&r->h[0], &x->h[0], &d);
  fp6_neg_p(&r->h[1], &a2);
}

// index of the Fq2 coefficient that multiplies w^k, k = 0..5
RB_FN constexpr int wk_index(int k) { return (k & 1) ? 3 + (k >> 1) : (k >> 1); }

// x -> x^(p^j), j = 1, 2, 3; r may alias x
static RB_NOINLINE void fp12_frobenius_to(Fp12* r, const Fp12* x, int j) {
  const Fp2* g = (j == 1) ? FROB1 : ((j == 2) ? FROB2 : FROB3);
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
  for (int k = 0; k < 6; ++k) {
    int idx = wk_index(k);
    Fp2 z = f12c(*x, idx);
    if (j & 1) z = fp2_conj(z);
    Fp2 gk = g[k];
    f12c(*r, idx) = fp2_mul(z, gk);
  }
}

// f *= l0 + l3 w^3 + l4 w^4  (value of a Miller line, see pairing.cuh); 15 Fq2 products.
static RB_NOINLINE void fp12_mul_by_line(Fp12* f, const Fp2* pl0, const Fp2* pl3, const Fp2* pl4) {
  Fp2 l0 = *pl0, l3 = *pl3, l4 = *pl4;
  Fp2 a0 = f->h[0].c[0], a1 = f->h[0].c[1], a2 = f->h[0].c[2];
  Fp2 b0 = f->h[1].c[0], b1 = f->h[1].c[1], b2 = f->h[1].c[2];
  Fp6 aa, bb, s, cr;
  aa.c[0] = fp2_add(fp2_mul(a0, l0), fp2_mul_xi(fp2_mul(a1, l4)));
  aa.c[1] = fp2_add(fp2_mul(a1, l0), fp2_mul_xi(fp2_mul(a2, l4)));
  aa.c[2] = fp2_add(fp2_mul(a2, l0), fp2_mul(a0, l4));
  bb.c[0] = fp2_mul_xi(fp2_mul(b2, l3));
  bb.c[1] = fp2_mul(b0, l3);
  bb.c[2] = fp2_mul(b1, l3);
  cr.c[0] = l0; cr.c[1] = l3; cr.c[2] = l4;
  s.c[0] = fp2_add(a0, b0); s.c[1] = fp2_add(a1, b1); s.c[2] = fp2_add(a2, b2);
  fp6_mul_p(&cr, &s, &cr);
  fp6_sub_p(&cr, &cr, &aa);
  fp6_sub_p(&f->h[1], &cr, &bb);
  fp6_mul_v_p(&bb, &bb);
  fp6_add_p(&f->h[0], &aa, &bb);
}

// f *= (a0 + a3 w^3 + a4 w^4) * (b0 + b3 w^3 + b4 w^4): two Miller lines that meet the same f
// (two pairs of one pairing product sharing the accumulator).  The lines are multiplied first --
// 6 Fq2 products, and the w^5 coefficient of the result is zero -- then f takes one 17-product
// multiplication: 23 Fq2 products instead of 2 x 15.
static RB_NOINLINE void fp12_mul_by_line_pair(Fp12* f, const Fp2* pa0, const Fp2* pa3, const Fp2* pa4,
                                              const Fp2* pb0, const Fp2* pb3, const Fp2* pb4) {
  Fp2 a0 = *pa0, a3 = *pa3, a4 = *pa4, b0 = *pb0, b3 = *pb3, b4 = *pb4;
  Fp6 s0, s1, aa, bb, sf;                       // line product = s0 + s1 w  (s1.c[2] == 0); all function scope (pairing.cuh note)
  Fp2 t00 = fp2_mul(a0, b0), t33 = fp2_mul(a3, b3), t44 = fp2_mul(a4, b4);
  Fp2 t04 = fp2_sub(fp2_sub(fp2_mul(fp2_add(a0, a4), fp2_add(b0, b4)), t00), t44);   // a0 b4 + a4 b0
  Fp2 t03 = fp2_sub(fp2_sub(fp2_mul(fp2_add(a0, a3), fp2_add(b0, b3)), t00), t33);   // a0 b3 + a3 b0
  Fp2 t34 = fp2_sub(fp2_sub(fp2_mul(fp2_add(a3, a4), fp2_add(b3, b4)), t33), t44);   // a3 b4 + a4 b3
  s0.c[0] = fp2_add(t00, fp2_mul_xi(t33)); s0.c[1] = fp2_mul_xi(t44); s0.c[2] = t04;
  s1.c[0] = fp2_mul_xi(t34); s1.c[1] = t03; s1.c[2] = fp2_zero();
  fp6_mul_p(&aa, &f->h[0], &s0);
  // bb = f.h1 * (s1.c0 + s1.c1 v): 5 Fq2 products
  Fp2 f0 = f->h[1].c[0], f1 = f->h[1].c[1], f2 = f->h[1].c[2], u0 = s1.c[0], u1 = s1.c[1];
  Fp2 p00 = fp2_mul(f0, u0), p11 = fp2_mul(f1, u1);
  Fp2 mid = fp2_sub(fp2_sub(fp2_mul(fp2_add(f0, f1), fp2_add(u0, u1)), p00), p11);   // f0 u1 + f1 u0
  Fp2 p21 = fp2_mul(f2, u1), p20 = fp2_mul(f2, u0);
  bb.c[0] = fp2_add(p00, fp2_mul_xi(p21)); bb.c[1] = mid; bb.c[2] = fp2_add(p11, p20);
  fp6_add_p(&sf, &f->h[0], &f->h[1]);
  fp6_add_p(&s0, &s0, &s1);
  fp6_mul_p(&s0, &sf, &s0);
  fp6_sub_p(&s0, &s0, &aa);
  fp6_sub_p(&f->h[1], &s0, &bb);
  fp6_mul_v_p(&bb, &bb);
  fp6_add_p(&f->h[0], &aa, &bb);
}

// Granger-Scott squaring; only valid for elements of the cyclotomic subgroup; r may alias x
static RB_NOINLINE void fp12_cyclotomic_sqr_to(Fp12* r, const Fp12* x) {
  Fp2 z0 = f12c(*x, 0), z4 = f12c(*x, 1), z3 = f12c(*x, 2), z2 = f12c(*x, 3), z1 = f12c(*x, 4), z5 = f12c(*x, 5);
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
  f12c(*r, 0) = fp2_add(fp2_dbl(fp2_sub(t0, z0)), t0);
  f12c(*r, 4) = fp2_add(fp2_dbl(fp2_add(t1, z1)), t1);
  tmp = fp2_mul_xi(t5);
  f12c(*r, 3) = fp2_add(fp2_dbl(fp2_add(tmp, z2)), tmp);
  f12c(*r, 2) = fp2_add(fp2_dbl(fp2_sub(t4, z3)), t4);
  f12c(*r, 1) = fp2_add(fp2_dbl(fp2_sub(t2, z4)), t2);
  f12c(*r, 5) = fp2_add(fp2_dbl(fp2_add(t3, z5)), t3);
}

// r = x^u for the BN parameter u = 4965661367192848881 (63 bits), x in the cyclotomic subgroup;
// r must NOT alias x
static RB_NOINLINE void fp12_cyclotomic_exp_u_to(Fp12* r, const Fp12* x) {
  const uint64_t u = 4965661367192848881ull;
  fp12_copy(r, x);
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
  for (int i = 61; i >= 0; --i) {
    fp12_cyclotomic_sqr_to(r, r);
    if ((u >> i) & 1) fp12_mul_to(r, r, x);
  }
}

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
