// G1 (over Fq) and G2 (over Fq2, the sextic D-twist y^2 = x^3 + 3/(9+i)) group arithmetic in
// extended-Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; ZZ = 0 is infinity).
// Replaces rabe_bn's `G1`/`G2` operators used by the schemes (`G * Fr`, `G + G`, `G - G`, `-G`,
// `G::zero()`: /root/reference/src/schemes/ac17/mod.rs:235,343-348,406-415; bsw/mod.rs:103-148,
// 233-241; lsw/mod.rs:99-153,204-212; aw11/mod.rs:145,224,275-276).
//
// XYZZ is chosen because the hot operation here is "accumulator += affine table entry" (fixed-base
// windows, gather sums): that mixed addition costs 8M + 2S, one less than Jacobian.
// All formulas are complete in the sense that the exceptional cases (equal / opposite / infinite
// operands) are detected and handled, so results are correct for arbitrary inputs.
#pragma once
#include "tower.cuh"

namespace rb {

template <class F> struct Affine { F x, y; };                 // (0,0) encodes infinity
template <class F> struct Xyzz { F x, y, zz, zzz; };

typedef Affine<Fp> G1Affine;
typedef Affine<Fp2> G2Affine;
typedef Xyzz<Fp> G1Xyzz;
typedef Xyzz<Fp2> G2Xyzz;

template <class F> RB_FN bool aff_is_inf(const Affine<F>& p) { return f_is_zero(p.x) && f_is_zero(p.y); }
template <class F> RB_FN void xyzz_set_inf(Xyzz<F>& r) { f_set_one(r.x); f_set_one(r.y); f_set_zero(r.zz); f_set_zero(r.zzz); }
template <class F> RB_FN bool xyzz_is_inf(const Xyzz<F>& p) { return f_is_zero(p.zz); }
template <class F> RB_FN void xyzz_from_affine(Xyzz<F>& r, const Affine<F>& p) {
  if (aff_is_inf(p)) { xyzz_set_inf(r); return; }
  r.x = p.x; r.y = p.y; f_set_one(r.zz); f_set_one(r.zzz);
}

// r = 2p  (dbl-2008-s-1, a = 0)
template <class F> RB_FN void xyzz_dbl(Xyzz<F>& r, const Xyzz<F>& p) {
  if (xyzz_is_inf(p)) { r = p; return; }
  F u = f_dbl(p.y);
  F v = f_sqr(u);
  F w = f_mul(u, v);
  F s = f_mul(p.x, v);
  F xx = f_sqr(p.x);
  F m = f_add(f_dbl(xx), xx);
  F x3 = f_sub(f_sqr(m), f_dbl(s));
  F y3 = f_sub(f_mul(m, f_sub(s, x3)), f_mul(w, p.y));
  r.zz = f_mul(v, p.zz);
  r.zzz = f_mul(w, p.zzz);
  r.x = x3; r.y = y3;
}

// r = 2q for affine q
template <class F> RB_FN void xyzz_dbl_affine(Xyzz<F>& r, const Affine<F>& q) {
  if (aff_is_inf(q)) { xyzz_set_inf(r); return; }
  F u = f_dbl(q.y);
  F v = f_sqr(u);
  F w = f_mul(u, v);
  F s = f_mul(q.x, v);
  F xx = f_sqr(q.x);
  F m = f_add(f_dbl(xx), xx);
  r.x = f_sub(f_sqr(m), f_dbl(s));
  r.y = f_sub(f_mul(m, f_sub(s, r.x)), f_mul(w, q.y));
  r.zz = v; r.zzz = w;
}

// acc += q, q affine  (madd-2008-s; 8M + 2S)
template <class F> RB_FN void xyzz_add_affine(Xyzz<F>& acc, const Affine<F>& q) {
  if (aff_is_inf(q)) return;
  if (xyzz_is_inf(acc)) { acc.x = q.x; acc.y = q.y; f_set_one(acc.zz); f_set_one(acc.zzz); return; }
  F u2 = f_mul(q.x, acc.zz);
  F s2 = f_mul(q.y, acc.zzz);
  F p = f_sub(u2, acc.x);
  F r = f_sub(s2, acc.y);
  if (f_is_zero(p)) {
    if (f_is_zero(r)) xyzz_dbl_affine(acc, q); else xyzz_set_inf(acc);
    return;
  }
  F pp = f_sqr(p);
  F ppp = f_mul(p, pp);
  F q1 = f_mul(acc.x, pp);
  F x3 = f_sub(f_sub(f_sqr(r), ppp), f_dbl(q1));
  F y3 = f_sub(f_mul(r, f_sub(q1, x3)), f_mul(acc.y, ppp));
  acc.zz = f_mul(acc.zz, pp);
  acc.zzz = f_mul(acc.zzz, ppp);
  acc.x = x3; acc.y = y3;
}

// acc = a + b for two AFFINE points (mmadd-2008-s: 4M + 2S -- the first addition of a fixed-base walk, where the
// accumulator is still an unscaled table entry); complete like the others
template <class F> RB_FN void xyzz_from_two_affine(Xyzz<F>& acc, const Affine<F>& a, const Affine<F>& b) {
  if (aff_is_inf(a)) { xyzz_from_affine(acc, b); return; }
  if (aff_is_inf(b)) { xyzz_from_affine(acc, a); return; }
  F p = f_sub(b.x, a.x);
  F r = f_sub(b.y, a.y);
  if (f_is_zero(p)) {
    if (f_is_zero(r)) xyzz_dbl_affine(acc, b); else xyzz_set_inf(acc);
    return;
  }
  F pp = f_sqr(p);
  F ppp = f_mul(p, pp);
  F q1 = f_mul(a.x, pp);
  F x3 = f_sub(f_sub(f_sqr(r), ppp), f_dbl(q1));
  acc.y = f_sub(f_mul(r, f_sub(q1, x3)), f_mul(a.y, ppp));
  acc.x = x3; acc.zz = pp; acc.zzz = ppp;
}

// acc += b  (add-2008-s; 12M + 2S)
template <class F> RB_FN void xyzz_add(Xyzz<F>& acc, const Xyzz<F>& b) {
  if (xyzz_is_inf(b)) return;
  if (xyzz_is_inf(acc)) { acc = b; return; }
  F u1 = f_mul(acc.x, b.zz), u2 = f_mul(b.x, acc.zz);
  F s1 = f_mul(acc.y, b.zzz), s2 = f_mul(b.y, acc.zzz);
  F p = f_sub(u2, u1), r = f_sub(s2, s1);
  if (f_is_zero(p)) {
    if (f_is_zero(r)) { Xyzz<F> t = acc; xyzz_dbl(acc, t); } else xyzz_set_inf(acc);
    return;
  }
  F pp = f_sqr(p);
  F ppp = f_mul(p, pp);
  F q1 = f_mul(u1, pp);
  F x3 = f_sub(f_sub(f_sqr(r), ppp), f_dbl(q1));
  F y3 = f_sub(f_mul(r, f_sub(q1, x3)), f_mul(s1, ppp));
  acc.zz = f_mul(f_mul(acc.zz, b.zz), pp);
  acc.zzz = f_mul(f_mul(acc.zzz, b.zzz), ppp);
  acc.x = x3; acc.y = y3;
}

template <class F> RB_FN Affine<F> aff_neg(const Affine<F>& p) { return {p.x, f_neg(p.y)}; }

// canonical scalar (NOT Montgomery) helpers: bit / window extraction from 8 x 32-bit limbs
RB_FN uint32_t scalar_window(const uint32_t* k, int bit, int width) {
  int limb = bit >> 5, sh = bit & 31;
  uint64_t two = (uint64_t)k[limb] | ((limb + 1 < 8) ? ((uint64_t)k[limb + 1] << 32) : 0ull);
  return (uint32_t)(two >> sh) & ((1u << width) - 1u);
}

// r = k * q by MSB-first double-and-add (variable base; used for table construction and the
// variable-base entry points, not on the fixed-base hot path)
template <class F> RB_FN void xyzz_mul_affine(Xyzz<F>& r, const Affine<F>& q, const uint32_t* k, int nbits) {
  xyzz_set_inf(r);
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
  for (int i = nbits - 1; i >= 0; --i) {
    Xyzz<F> t = r;
    xyzz_dbl(r, t);
    if ((k[i >> 5] >> (i & 31)) & 1u) xyzz_add_affine(r, q);
  }
}

// affine from XYZZ given inv = 1/(zz*zzz):  1/zz = inv*zzz, 1/zzz = inv*zz
template <class F> RB_FN Affine<F> xyzz_to_affine_with(const Xyzz<F>& p, const F& inv) {
  Affine<F> a;
  a.x = f_mul(p.x, f_mul(inv, p.zzz));
  a.y = f_mul(p.y, f_mul(inv, p.zz));
  return a;
}

// G1 canonical bytes (x|y big-endian, all-zero = infinity) <-> Montgomery affine
RB_FN G1Affine g1_load_be(const uint8_t* p) {
  G1Affine a;
  a.x = fe_to_mont(fe_load_be<ModP>(p));
  a.y = fe_to_mont(fe_load_be<ModP>(p + 32));
  return a;
}
RB_FN void g1_store_be(uint8_t* p, const G1Affine& a) {
  fe_store_be(p, fe_from_mont(a.x));
  fe_store_be(p + 32, fe_from_mont(a.y));
}
RB_FN G2Affine g2_load_be(const uint8_t* p) {
  G2Affine a;
  a.x.a = fe_to_mont(fe_load_be<ModP>(p));
  a.x.b = fe_to_mont(fe_load_be<ModP>(p + 32));
  a.y.a = fe_to_mont(fe_load_be<ModP>(p + 64));
  a.y.b = fe_to_mont(fe_load_be<ModP>(p + 96));
  return a;
}
RB_FN void g2_store_be(uint8_t* p, const G2Affine& a) {
  fe_store_be(p, fe_from_mont(a.x.a));
  fe_store_be(p + 32, fe_from_mont(a.x.b));
  fe_store_be(p + 64, fe_from_mont(a.y.a));
  fe_store_be(p + 96, fe_from_mont(a.y.b));
}

// membership checks for untrusted inputs (range is checked by the caller on the raw limbs)
RB_FN bool g1_on_curve(const G1Affine& a) {
  if (aff_is_inf(a)) return true;
  Fp three = fe_to_mont(Fp{{3, 0, 0, 0, 0, 0, 0, 0}});
  return fe_eq(fe_sqr(a.y), fe_sqr(a.x) * a.x + three);
}
RB_FN bool g2_on_curve(const G2Affine& a) {
  if (aff_is_inf(a)) return true;
  Fp2 b = TWIST_B;
  return fp2_eq(fp2_sqr(a.y), fp2_add(fp2_mul(fp2_sqr(a.x), a.x), b));
}

}  // namespace rb
