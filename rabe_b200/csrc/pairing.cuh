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

// f <- f * miller(P, Q) for affine, finite P and Q.  `first` = f is known to be one (skips the
// first squaring).  The accumulator may already hold other Miller values ONLY if they were
// accumulated by the same loop (shared squarings); use miller_single + fp12_mul otherwise.
static RB_NOINLINE void miller_single(Fp12* f, const G1Affine* p, const G2Affine* q) {
  const uint64_t loop_lo = 0x9d797039be763ba8ull;    // 6u+2 = 2^64 + loop_lo; top bit consumed by T = Q
  G2Homog t; t.x = q->x; t.y = q->y; t.z = fp2_one();
  Fp2 l0, l3, l4;
  fp12_set_one(*f);
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
  for (int i = 63; i >= 0; --i) {
    if (i != 63) fp12_sqr_to(f, f);
    miller_dbl_step(&t, &l0, &l3, &l4);
    l3 = fp2_mul_fp(l3, p->y); l4 = fp2_mul_fp(l4, p->x);
    fp12_mul_by_line(f, &l0, &l3, &l4);
    if ((loop_lo >> i) & 1) {
      miller_add_step(&t, &q->x, &q->y, &l0, &l3, &l4);
      l3 = fp2_mul_fp(l3, p->y); l4 = fp2_mul_fp(l4, p->x);
      fp12_mul_by_line(f, &l0, &l3, &l4);
    }
  }
  // Frobenius endomorphism steps: Q1 = pi(Q), Q2 = -pi^2(Q)
  Fp2 q1x = fp2_mul(fp2_conj(q->x), FROB1[2]);
  Fp2 q1y = fp2_mul(fp2_conj(q->y), FROB1[3]);
  miller_add_step(&t, &q1x, &q1y, &l0, &l3, &l4);
  l3 = fp2_mul_fp(l3, p->y); l4 = fp2_mul_fp(l4, p->x);
  fp12_mul_by_line(f, &l0, &l3, &l4);
  Fp2 q2x = fp2_mul(q->x, FROB2[2]);
  Fp2 q2y = fp2_neg(fp2_mul(q->y, FROB2[3]));
  miller_add_step(&t, &q2x, &q2y, &l0, &l3, &l4);
  l3 = fp2_mul_fp(l3, p->y); l4 = fp2_mul_fp(l4, p->x);
  fp12_mul_by_line(f, &l0, &l3, &l4);
}

// r = f^(-u) for f in the cyclotomic subgroup
RB_FN void exp_neg_u(Fp12* r, const Fp12* f) {
  Fp12 t;
  fp12_cyclotomic_exp_u_to(&t, f);
  fp12_conj_to(r, &t);
}

static RB_NOINLINE void final_exponentiation(Fp12* out, const Fp12* in) {
  Fp12 x, a, b, c, d, e, g, k, l, t;
  // easy part
  fp12_inv_to(&t, in);
  fp12_conj_to(&a, in);
  fp12_mul_to(&a, &a, &t);                 // f^(p^6-1)
  fp12_frobenius_to(&t, &a, 2);
  fp12_mul_to(&x, &t, &a);                 // ^(p^2+1)
  // hard part
  exp_neg_u(&a, &x);                       // A = x^-u
  fp12_cyclotomic_sqr_to(&b, &a);          // B = A^2
  fp12_cyclotomic_sqr_to(&c, &b);          // C = B^2
  fp12_mul_to(&d, &c, &b);                 // D = C*B
  exp_neg_u(&e, &d);                       // E = D^-u
  fp12_cyclotomic_sqr_to(&t, &e);          // F = E^2
  exp_neg_u(&g, &t);                       // G = F^-u
  fp12_conj_to(&t, &g);                    // I = 1/G
  fp12_mul_to(&t, &t, &e);                 // J = I*E
  fp12_conj_to(&c, &d);                    // H = 1/D
  fp12_mul_to(&k, &t, &c);                 // K = J*H
  fp12_mul_to(&l, &k, &b);                 // L = K*B
  fp12_mul_to(&t, &k, &e);                 // M = K*E
  fp12_mul_to(&t, &t, &x);                 // N = M*x
  fp12_frobenius_to(&c, &l, 1);            // O = L^p
  fp12_mul_to(&t, &c, &t);                 // P = O*N
  fp12_frobenius_to(&c, &k, 2);            // Q = K^(p^2)
  fp12_mul_to(&t, &c, &t);                 // R = Q*P
  fp12_conj_to(&c, &x);                    // S = 1/x
  fp12_mul_to(&c, &c, &l);                 // T = S*L
  fp12_frobenius_to(&d, &c, 3);            // U = T^(p^3)
  fp12_mul_to(out, &d, &t);                // V = U*R
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
