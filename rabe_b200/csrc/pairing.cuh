// Optimal-ate pairing on BN254 for sm_100a: Miller loop over 6u+2 with the G2 argument kept in
// homogeneous projective coordinates on the twist, sparse line multiplication, and the final
// exponentiation f^((p^6-1)(p^2+1)) followed by the hard part f^LAMBDA with
//   LAMBDA = p^3(12u^3+6u^2+4u-1) + p^2(12u^3+6u^2+6u) + p(12u^3+6u^2+4u) + (12u^3+12u^2+6u+1)
// = 2u(6u^2+3u+1) * (p^4-p^2+1)/r  -- the exponent of the zcash-bn lineage that rabe-bn 0.4.23
// forks, so Gt values agree with the oracle bit for bit.
//
// Replaces `rabe_bn::pairing(G1, G2)` at /root/reference/src/schemes/ac17/mod.rs:148,415-416,
// bsw/mod.rs:108,292-293,308, lsw/mod.rs:103,275-276, aw11/mod.rs:144,263,274,340-341.
// A product of pairings shares one final exponentiation (the Miller values are multiplied first).
#pragma once
#include "curve.cuh"

namespace rb {

struct G2Homog { Fp2 x, y, z; };

// Line through T (tangent) evaluated at P, scaled into l0 + (l3*yP) w^3 + (l4*xP) w^4; T <- 2T.
RB_FN void miller_dbl_step(G2Homog* t, Fp2* l0, Fp2* l3, Fp2* l4) {
  Fp two_inv = TWO_INV;
  Fp2 tb = TWIST_B;
  Fp2 a = fp2_mul_fp(fp2_mul(t->x, t->y), two_inv);
  Fp2 b = fp2_sqr(t->y);
  Fp2 c = fp2_sqr(t->z);
  Fp2 e = fp2_mul(tb, fp2_add(fp2_dbl(c), c));
  Fp2 f = fp2_add(fp2_dbl(e), e);
  Fp2 g = fp2_mul_fp(fp2_add(b, f), two_inv);
  Fp2 h = fp2_sub(fp2_sqr(fp2_add(t->y, t->z)), fp2_add(b, c));
  Fp2 j = fp2_sqr(t->x);
  Fp2 e2 = fp2_sqr(e);
  *l0 = fp2_mul_xi(fp2_sub(e, b));
  *l3 = fp2_neg(h);
  *l4 = fp2_add(fp2_dbl(j), j);
  t->x = fp2_mul(a, fp2_sub(b, f));
  t->y = fp2_sub(fp2_sqr(g), fp2_add(fp2_dbl(e2), e2));
  t->z = fp2_mul(b, h);
}

// Line through T and the affine point Q evaluated at P; T <- T + Q.
RB_FN void miller_add_step(G2Homog* t, const Fp2* qx, const Fp2* qy, Fp2* l0, Fp2* l3, Fp2* l4) {
  Fp2 d = fp2_sub(t->x, fp2_mul(*qx, t->z));
  Fp2 e = fp2_sub(t->y, fp2_mul(*qy, t->z));
  Fp2 f = fp2_sqr(d);
  Fp2 g = fp2_sqr(e);
  Fp2 h = fp2_mul(d, f);
  Fp2 i = fp2_mul(t->x, f);
  Fp2 j = fp2_sub(fp2_add(h, fp2_mul(t->z, g)), fp2_dbl(i));
  *l0 = fp2_mul_xi(fp2_sub(fp2_mul(e, *qx), fp2_mul(d, *qy)));
  *l3 = d;
  *l4 = fp2_neg(e);
  t->x = fp2_mul(d, j);
  t->y = fp2_sub(fp2_mul(e, fp2_sub(i, j)), fp2_mul(h, t->y));
  t->z = fp2_mul(t->z, h);
}

// f = miller(P, Q) for affine, finite P and Q: f_{6u+2,Q}(P) times the two Frobenius lines, with
// 6u+2 walked in non-adjacent form (65 doubling steps, 21 addition steps of +-Q).  Miller values
// are defined up to factors that the final exponentiation kills, so only post-exponentiation
// values are comparable with other implementations.
struct MillerLine { Fp2 l0, l3, l4; };        // line value = l0 + (l3*yP) w^3 + (l4*xP) w^4
constexpr int MILLER_LINES = (ATE_NAF_LEN - 1) + 21 + 2;   // 65 tangents + 21 chords + 2 Frobenius chords

// NOTE (nvcc 12.9): every object whose address is handed to an out-of-line routine below is
// declared at FUNCTION scope.  Block-scoped temporaries inside the loops had their stack slots
// merged with live objects (wrong results on the device, correct on the host; tools/dbg/).

static RB_NOINLINE void miller_single(Fp12* f, const G1Affine* p, const G2Affine* q) {
  G2Homog t; t.x = q->x; t.y = q->y; t.z = fp2_one();
  G2Affine qq = *q;
  G1Affine pt = *p;
  Fp2 nqy = fp2_neg(qq.y);
  Fp2 l0, l3, l4, qx2, qy2;
  fp12_set_one(*f);
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
  for (int i = ATE_NAF_LEN - 2; i >= 0; --i) {
    if (i != ATE_NAF_LEN - 2) fp12_sqr_to(f, f);
    miller_dbl_step(&t, &l0, &l3, &l4);
    l3 = fp2_mul_fp(l3, pt.y); l4 = fp2_mul_fp(l4, pt.x);
    fp12_mul_by_line(f, &l0, &l3, &l4);
    int d = ATE_NAF[i];
    if (d != 0) {
      qy2 = d > 0 ? qq.y : nqy;
      miller_add_step(&t, &qq.x, &qy2, &l0, &l3, &l4);
      l3 = fp2_mul_fp(l3, pt.y); l4 = fp2_mul_fp(l4, pt.x);
      fp12_mul_by_line(f, &l0, &l3, &l4);
    }
  }
  // Frobenius endomorphism steps: Q1 = pi(Q), Q2 = -pi^2(Q)
  qx2 = fp2_mul(fp2_conj(qq.x), FROB1[2]);
  qy2 = fp2_mul(fp2_conj(qq.y), FROB1[3]);
  miller_add_step(&t, &qx2, &qy2, &l0, &l3, &l4);
  l3 = fp2_mul_fp(l3, pt.y); l4 = fp2_mul_fp(l4, pt.x);
  fp12_mul_by_line(f, &l0, &l3, &l4);
  qx2 = fp2_mul(qq.x, FROB2[2]);
  qy2 = fp2_neg(fp2_mul(qq.y, FROB2[3]));
  miller_add_step(&t, &qx2, &qy2, &l0, &l3, &l4);
  l3 = fp2_mul_fp(l3, pt.y); l4 = fp2_mul_fp(l4, pt.x);
  fp12_mul_by_line(f, &l0, &l3, &l4);
}

// Fixed-argument pairing: everything that depends only on Q (the walk of T and the line
// coefficients) is computed once; `lines` receives MILLER_LINES entries in evaluation order.
static RB_NOINLINE void miller_lines_for(MillerLine* lines, const G2Affine* q) {
  G2Homog t; t.x = q->x; t.y = q->y; t.z = fp2_one();
  G2Affine qq = *q;
  Fp2 nqy = fp2_neg(qq.y);
  Fp2 l0, l3, l4, qx2, qy2;
  int n = 0;
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
  for (int i = ATE_NAF_LEN - 2; i >= 0; --i) {
    miller_dbl_step(&t, &l0, &l3, &l4);
    lines[n].l0 = l0; lines[n].l3 = l3; lines[n].l4 = l4; ++n;
    int d = ATE_NAF[i];
    if (d != 0) {
      qy2 = d > 0 ? qq.y : nqy;
      miller_add_step(&t, &qq.x, &qy2, &l0, &l3, &l4);
      lines[n].l0 = l0; lines[n].l3 = l3; lines[n].l4 = l4; ++n;
    }
  }
  qx2 = fp2_mul(fp2_conj(qq.x), FROB1[2]);
  qy2 = fp2_mul(fp2_conj(qq.y), FROB1[3]);
  miller_add_step(&t, &qx2, &qy2, &l0, &l3, &l4);
  lines[n].l0 = l0; lines[n].l3 = l3; lines[n].l4 = l4; ++n;
  qx2 = fp2_mul(qq.x, FROB2[2]);
  qy2 = fp2_neg(fp2_mul(qq.y, FROB2[3]));
  miller_add_step(&t, &qx2, &qy2, &l0, &l3, &l4);
  lines[n].l0 = l0; lines[n].l3 = l3; lines[n].l4 = l4;
}

// f = miller(P, Q) from the precomputed lines of Q (same value as miller_single(P, Q))
static RB_NOINLINE void miller_fixed(Fp12* f, const G1Affine* p, const MillerLine* lines) {
  // all operands of the out-of-line Fq12 routines are function-scope objects (see tower.cuh note)
  Fp2 l0, l3, l4;
  G1Affine pt = *p;
  fp12_set_one(*f);
  int li = 0;
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
  for (int i = ATE_NAF_LEN - 2; i >= 0; --i) {
    if (i != ATE_NAF_LEN - 2) fp12_sqr_to(f, f);
    l0 = lines[li].l0; l3 = fp2_mul_fp(lines[li].l3, pt.y); l4 = fp2_mul_fp(lines[li].l4, pt.x); ++li;
    fp12_mul_by_line(f, &l0, &l3, &l4);
    if (ATE_NAF[i] != 0) {
      l0 = lines[li].l0; l3 = fp2_mul_fp(lines[li].l3, pt.y); l4 = fp2_mul_fp(lines[li].l4, pt.x); ++li;
      fp12_mul_by_line(f, &l0, &l3, &l4);
    }
  }
  l0 = lines[li].l0; l3 = fp2_mul_fp(lines[li].l3, pt.y); l4 = fp2_mul_fp(lines[li].l4, pt.x); ++li;
  fp12_mul_by_line(f, &l0, &l3, &l4);
  l0 = lines[li].l0; l3 = fp2_mul_fp(lines[li].l3, pt.y); l4 = fp2_mul_fp(lines[li].l4, pt.x);
  fp12_mul_by_line(f, &l0, &l3, &l4);
}

// f = miller(pv, q) * miller(pf, Q_fixed): one variable-argument pair and one fixed-argument pair
// (precomputed `lines` of Q_fixed) walked together.  Both pairs follow the same NAF of 6u+2, so
// they share the accumulator: one f^2 per step instead of two, and the two line values of a step
// are multiplied into f together (fp12_mul_by_line_pair).  Field arithmetic is exact, so the value
// equals miller_single(pv, q) * miller_fixed(pf, lines) coefficient for coefficient.
// (AC17 decrypt: e(-(k_p[j]+prod_h_j), c_0[j]) * e(prod_g_j, k_0[j]), ac17/mod.rs:415-416.)
static RB_NOINLINE void miller_pair(Fp12* f, const G1Affine* pv, const G2Affine* q, const G1Affine* pf, const MillerLine* lines) {
  G2Homog t; t.x = q->x; t.y = q->y; t.z = fp2_one();
  G2Affine qq = *q;
  G1Affine ptv = *pv, ptf = *pf;
  Fp2 nqy = fp2_neg(qq.y);
  Fp2 l0, l3, l4, m0, m3, m4, qx2, qy2;       // function scope: see the note above miller_single
  fp12_set_one(*f);
  int li = 0;
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
  for (int i = ATE_NAF_LEN - 2; i >= 0; --i) {
    if (i != ATE_NAF_LEN - 2) fp12_sqr_to(f, f);
    miller_dbl_step(&t, &l0, &l3, &l4);
    l3 = fp2_mul_fp(l3, ptv.y); l4 = fp2_mul_fp(l4, ptv.x);
    m0 = lines[li].l0; m3 = fp2_mul_fp(lines[li].l3, ptf.y); m4 = fp2_mul_fp(lines[li].l4, ptf.x); ++li;
    fp12_mul_by_line_pair(f, &l0, &l3, &l4, &m0, &m3, &m4);
    int d = ATE_NAF[i];
    if (d != 0) {
      qy2 = d > 0 ? qq.y : nqy;
      miller_add_step(&t, &qq.x, &qy2, &l0, &l3, &l4);
      l3 = fp2_mul_fp(l3, ptv.y); l4 = fp2_mul_fp(l4, ptv.x);
      m0 = lines[li].l0; m3 = fp2_mul_fp(lines[li].l3, ptf.y); m4 = fp2_mul_fp(lines[li].l4, ptf.x); ++li;
      fp12_mul_by_line_pair(f, &l0, &l3, &l4, &m0, &m3, &m4);
    }
  }
  qx2 = fp2_mul(fp2_conj(qq.x), FROB1[2]);
  qy2 = fp2_mul(fp2_conj(qq.y), FROB1[3]);
  miller_add_step(&t, &qx2, &qy2, &l0, &l3, &l4);
  l3 = fp2_mul_fp(l3, ptv.y); l4 = fp2_mul_fp(l4, ptv.x);
  m0 = lines[li].l0; m3 = fp2_mul_fp(lines[li].l3, ptf.y); m4 = fp2_mul_fp(lines[li].l4, ptf.x); ++li;
  fp12_mul_by_line_pair(f, &l0, &l3, &l4, &m0, &m3, &m4);
  qx2 = fp2_mul(qq.x, FROB2[2]);
  qy2 = fp2_neg(fp2_mul(qq.y, FROB2[3]));
  miller_add_step(&t, &qx2, &qy2, &l0, &l3, &l4);
  l3 = fp2_mul_fp(l3, ptv.y); l4 = fp2_mul_fp(l4, ptv.x);
  m0 = lines[li].l0; m3 = fp2_mul_fp(lines[li].l3, ptf.y); m4 = fp2_mul_fp(lines[li].l4, ptf.x);
  fp12_mul_by_line_pair(f, &l0, &l3, &l4, &m0, &m3, &m4);
}

// r = f^(-u) for f in the cyclotomic subgroup
RB_FN void exp_neg_u(Fp12* r, const Fp12* f, Fp12* scratch) {      // r, f, scratch pairwise distinct
  fp12_cyclotomic_exp_u_to(scratch, f);
  fp12_conj_to(r, scratch);
}

static RB_NOINLINE void final_exponentiation(Fp12* out, const Fp12* in) {
  Fp12 x, a, b, d, e, k, t, sc;              // letters in the comments follow the addition chain of DESIGN.md
  // easy part
  fp12_inv_to(&t, in);
  fp12_conj_to(&a, in);
  fp12_mul_to(&a, &a, &t);                 // f^(p^6-1)
  fp12_frobenius_to(&t, &a, 2);
  fp12_mul_to(&x, &t, &a);                 // ^(p^2+1)
  // hard part
  exp_neg_u(&a, &x, &sc);                  // A = x^-u
  fp12_cyclotomic_sqr_to(&b, &a);          // B = A^2
  fp12_cyclotomic_sqr_to(&a, &b);          // C = B^2          (a := C)
  fp12_mul_to(&d, &a, &b);                 // D = C*B
  exp_neg_u(&e, &d, &sc);                  // E = D^-u
  fp12_cyclotomic_sqr_to(&t, &e);          // F = E^2
  exp_neg_u(&a, &t, &sc);                  // G = F^-u         (a := G)
  fp12_conj_to(&t, &a);                    // I = 1/G
  fp12_mul_to(&t, &t, &e);                 // J = I*E
  fp12_conj_to(&a, &d);                    // H = 1/D          (a := H)
  fp12_mul_to(&k, &t, &a);                 // K = J*H
  fp12_mul_to(&d, &k, &b);                 // L = K*B          (d := L)
  fp12_mul_to(&t, &k, &e);                 // M = K*E
  fp12_mul_to(&t, &t, &x);                 // N = M*x
  fp12_frobenius_to(&a, &d, 1);            // O = L^p
  fp12_mul_to(&t, &a, &t);                 // P = O*N
  fp12_frobenius_to(&a, &k, 2);            // Q = K^(p^2)
  fp12_mul_to(&t, &a, &t);                 // R = Q*P
  fp12_conj_to(&a, &x);                    // S = 1/x
  fp12_mul_to(&a, &a, &d);                 // T = S*L
  fp12_frobenius_to(&b, &a, 3);            // U = T^(p^3)
  fp12_mul_to(out, &b, &t);                // V = U*R
}

// Gt^k by MSB-first square-and-multiply over a canonical (non-Montgomery) scalar
static RB_NOINLINE void fp12_pow(Fp12* r, const Fp12* base, const uint32_t* k) {
  Fp12 acc, b; fp12_set_one(acc); fp12_copy(&b, base);     // private copies: r may alias base
  bool started = false;
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
  for (int i = 255; i >= 0; --i) {
    if (started) fp12_sqr_to(&acc, &acc);
    if ((k[i >> 5] >> (i & 31)) & 1u) {
      if (started) fp12_mul_to(&acc, &acc, &b); else { fp12_copy(&acc, &b); started = true; }
    }
  }
  fp12_copy(r, &acc);
}

}  // namespace rb
