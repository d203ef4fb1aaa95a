// TEST INFRASTRUCTURE ONLY: compiles the device headers of rabe_b200/csrc for the HOST
// (-DRB_HOST_SIM) so that the algorithmic layer (tower, curve, Miller loop, final exponentiation)
// can be checked against the oracle without a GPU.  The PTX Montgomery product is replaced by a
// portable one here, so this does NOT validate the device carry chains -- tests marked `gpu` do.
// Never linked into the product library.
#define RB_HOST_SIM 1
#include "../../rabe_b200/csrc/pairing.cuh"
#include <cstring>
using namespace rb;

extern "C" {
void hs_fp_mul(const uint8_t* a, const uint8_t* b, uint8_t* out) {
  Fp x = fe_to_mont(fe_load_be<ModP>(a)), y = fe_to_mont(fe_load_be<ModP>(b));
  fe_store_be(out, fe_from_mont(x * y));
}
void hs_fr_mul(const uint8_t* a, const uint8_t* b, uint8_t* out) {
  Fr x = fe_to_mont(fe_load_be<ModR>(a)), y = fe_to_mont(fe_load_be<ModR>(b));
  fe_store_be(out, fe_from_mont(x * y));
}
void hs_fp_inv(const uint8_t* a, uint8_t* out) {
  fe_store_be(out, fe_from_mont(fe_inv(fe_to_mont(fe_load_be<ModP>(a)))));
}
static void load_scalar(const uint8_t* k, uint32_t* w) { Fr s = fe_load_be<ModR>(k); memcpy(w, s.v, 32); }
void hs_g1_mul(const uint8_t* base, const uint8_t* k, uint8_t* out) {
  G1Affine b = g1_load_be(base); uint32_t w[8]; load_scalar(k, w);
  G1Xyzz r; xyzz_mul_affine(r, b, w, 256);
  if (xyzz_is_inf(r)) { memset(out, 0, 64); return; }
  Fp inv = fe_inv(r.zz * r.zzz);
  g1_store_be(out, xyzz_to_affine_with(r, inv));
}
void hs_g2_mul(const uint8_t* base, const uint8_t* k, uint8_t* out) {
  G2Affine b = g2_load_be(base); uint32_t w[8]; load_scalar(k, w);
  G2Xyzz r; xyzz_mul_affine(r, b, w, 256);
  if (xyzz_is_inf(r)) { memset(out, 0, 128); return; }
  Fp2 inv = fp2_inv(fp2_mul(r.zz, r.zzz));
  g2_store_be(out, xyzz_to_affine_with(r, inv));
}
void hs_g1_add(const uint8_t* a, const uint8_t* b, uint8_t* out) {
  G1Affine x = g1_load_be(a), y = g1_load_be(b);
  G1Xyzz r, s; xyzz_from_affine(r, x); xyzz_add_affine(r, y);
  xyzz_from_affine(s, x); G1Xyzz t; xyzz_from_affine(t, y); xyzz_add(s, t);
  if (xyzz_is_inf(r)) { memset(out, 0, 64); return; }
  G1Affine ra = xyzz_to_affine_with(r, fe_inv(r.zz * r.zzz));
  G1Affine sa = xyzz_to_affine_with(s, fe_inv(s.zz * s.zzz));
  if (!fe_eq(ra.x, sa.x) || !fe_eq(ra.y, sa.y)) { memset(out, 0xff, 64); return; }
  g1_store_be(out, ra);
}
void hs_pairing(const uint8_t* p1, const uint8_t* q2, uint8_t* out) {
  G1Affine p = g1_load_be(p1); G2Affine q = g2_load_be(q2);
  Fp12 f, r; miller_single(&f, &p, &q); final_exponentiation(&r, &f);
  fp12_store_be(out, r);
}
void hs_pairing_fixed(const uint8_t* p1, const uint8_t* q2, uint8_t* out) {
  G1Affine p = g1_load_be(p1); G2Affine q = g2_load_be(q2);
  static MillerLine lines[MILLER_LINES];
  miller_lines_for(lines, &q);
  Fp12 f, r; miller_fixed(&f, &p, lines); final_exponentiation(&r, &f);
  fp12_store_be(out, r);
}
// FE(miller_pair(pv, qv, pf, lines(qf))) -- must equal e(pv, qv) * e(pf, qf)
void hs_pairing_pair(const uint8_t* pv1, const uint8_t* qv2, const uint8_t* pf1, const uint8_t* qf2, uint8_t* out) {
  G1Affine pv = g1_load_be(pv1), pf = g1_load_be(pf1); G2Affine qv = g2_load_be(qv2), qf = g2_load_be(qf2);
  static MillerLine lines[MILLER_LINES];
  miller_lines_for(lines, &qf);
  Fp12 f, r; miller_pair(&f, &pv, &qv, &pf, lines); final_exponentiation(&r, &f);
  fp12_store_be(out, r);
}
// the same through a line table normalised to l0 = 1 and the cheaper line product (miller_pair(..., unit_lines = true))
void hs_pairing_pair_unit(const uint8_t* pv1, const uint8_t* qv2, const uint8_t* pf1, const uint8_t* qf2, uint8_t* out) {
  G1Affine pv = g1_load_be(pv1), pf = g1_load_be(pf1); G2Affine qv = g2_load_be(qv2), qf = g2_load_be(qf2);
  static MillerLine lines[MILLER_LINES], unit[MILLER_LINES];
  miller_lines_for(lines, &qf);
  if (!miller_lines_normalize(unit, lines, MILLER_LINES)) { memset(out, 0xff, 384); return; }
  Fp12 f, r; miller_pair(&f, &pv, &qv, &pf, unit, true); final_exponentiation(&r, &f);
  fp12_store_be(out, r);
}
// FE(miller_fixed4) over up to four (P_k, Q_k) pairs; mask bit k = pair k present
void hs_pairing_fixed4(const uint8_t* p1s, const uint8_t* q2s, int mask, int unit, uint8_t* out) {
  static MillerLine lines[4][MILLER_LINES], raw[MILLER_LINES];
  G1Affine p[4]; const MillerLine* lp[4]; bool present[4];
  int first = -1;
  for (int k = 0; k < 4; ++k) if ((mask >> k) & 1) { first = (first < 0) ? k : first; }
  for (int k = 0; k < 4; ++k) {
    present[k] = (mask >> k) & 1;
    int src = present[k] ? k : first;
    p[k] = g1_load_be(p1s + 64 * src);
    G2Affine q = g2_load_be(q2s + 128 * src);
    if (unit) {
      miller_lines_for(raw, &q);
      if (!miller_lines_normalize(lines[k], raw, MILLER_LINES)) { memset(out, 0xff, 384); return; }
    } else miller_lines_for(lines[k], &q);
    lp[k] = lines[k];
  }
  Fp12 f, r; miller_fixed4(&f, p, lp, present, unit != 0); final_exponentiation(&r, &f);
  fp12_store_be(out, r);
}
// FE(miller_pair3) over three (variable, fixed) pair couples -- must equal the product of the six pairings
void hs_pairing_pair3(const uint8_t* pv1, const uint8_t* qv2, const uint8_t* pf1, const uint8_t* qf2, uint8_t* out) {
  static MillerLine lines[3 * MILLER_LINES];
  G1Affine pv[3], pf[3]; G2Affine qv[3];
  for (int j = 0; j < 3; ++j) {
    pv[j] = g1_load_be(pv1 + 64 * j); pf[j] = g1_load_be(pf1 + 64 * j); qv[j] = g2_load_be(qv2 + 128 * j);
    G2Affine qf = g2_load_be(qf2 + 128 * j);
    miller_lines_for(lines + j * MILLER_LINES, &qf);
  }
  Fp12 f, r; miller_pair3(&f, pv, qv, pf, lines); final_exponentiation(&r, &f);
  fp12_store_be(out, r);
}
void hs_gt_pow(const uint8_t* a, const uint8_t* k, uint8_t* out) {
  Fp12 x, r; fp12_load_be(x, a); uint32_t w[8]; load_scalar(k, w);
  fp12_pow(&r, &x, w); fp12_store_be(out, r);
}
// Fp-product counts of the device primitives (what one GPU thread executes), for the roofline
// numerators in bench.py / DESIGN.md.  out: see names in tools/gen_op_counts.py.
void hs_op_counts(unsigned long long* out) {
  G1Affine p; p.x = fe_one<ModP>(); p.y = fe_dbl(fe_one<ModP>());
  G2Affine q; q.x = G2_GEN_X; q.y = G2_GEN_Y;
  unsigned long long c0; int i = 0;
#define COUNT(stmt) c0 = g_host_mul_count; stmt; out[i++] = g_host_mul_count - c0;
  Fp a = p.y, b;
  COUNT(b = fe_inv(a));                                                     // 0 fe_inv
  G1Xyzz acc; xyzz_dbl_affine(acc, p);
  COUNT(xyzz_add_affine(acc, p));                                           // 1 g1_madd
  { G1Xyzz t = acc; COUNT(xyzz_dbl(acc, t)); }                              // 2 g1_dbl
  G2Xyzz acc2; xyzz_dbl_affine(acc2, q);
  COUNT(xyzz_add_affine(acc2, q));                                          // 3 g2_madd
  { G2Xyzz t = acc2; COUNT(xyzz_dbl(acc2, t)); }                            // 4 g2_dbl
  Fp2 x2 = q.x, y2;
  COUNT(y2 = fp2_inv(x2));                                                  // 5 fp2_inv
  Fp12 f, g, r;
  COUNT(miller_single(&f, &p, &q));                                         // 6 miller_single
  COUNT(final_exponentiation(&g, &f));                                      // 7 final_exponentiation
  COUNT(fp12_mul_to(&r, &f, &g));                                           // 8 fp12_mul
  COUNT(fp12_sqr_to(&r, &f));                                               // 9 fp12_sqr
  COUNT(fp12_cyclotomic_sqr_to(&r, &g));                                    // 10 fp12_cyclotomic_sqr
  COUNT(fp12_mul_by_line(&r, &q.x, &q.y, &q.x));                            // 11 fp12_mul_by_line
  COUNT(fp12_inv_to(&r, &f));                                               // 12 fp12_inv
  COUNT(b = fe_to_mont(a));                                                 // 13 to_mont (== from_mont)
  COUNT(g1_on_curve(p));                                                    // 14 g1_on_curve
  COUNT(g2_on_curve(q));                                                    // 15 g2_on_curve
  static MillerLine lines[MILLER_LINES];
  COUNT(miller_lines_for(lines, &q));                                       // 16 miller_lines_for
  COUNT(miller_fixed(&f, &p, lines));                                       // 17 miller_fixed
  COUNT(miller_pair(&f, &p, &q, &p, lines));                                // 18 miller_pair
  COUNT(fp12_mul_by_line_pair(&r, &q.x, &q.y, &q.x, &q.y, &q.x, &q.y));     // 19 fp12_mul_by_line_pair
  { G1Affine p4[4] = {p, p, p, p}; const MillerLine* l4[4] = {lines, lines, lines, lines}; bool pr[4] = {true, true, true, true};
    COUNT(miller_fixed4(&f, p4, l4, pr)); }                                 // 20 miller_fixed4 (four pairs)
  { static MillerLine l3[3 * MILLER_LINES]; for (int j = 0; j < 3; ++j) miller_lines_for(l3 + j * MILLER_LINES, &q);
    G1Affine p3[3] = {p, p, p}; G2Affine q3[3] = {q, q, q};
    COUNT(miller_pair3(&f, p3, q3, p3, l3)); }                              // 21 miller_pair3 (three terms)
  COUNT(miller_pair(&f, &p, &q, &p, lines, true));                          // 22 miller_pair_unit (fixed lines normalised to l0 = 1)
  { G1Affine p4[4] = {p, p, p, p}; const MillerLine* l4[4] = {lines, lines, lines, lines}; bool pr[4] = {true, true, true, true};
    COUNT(miller_fixed4(&f, p4, l4, pr, true)); }                           // 23 miller_fixed4_unit
  { static MillerLine unit[MILLER_LINES]; COUNT(miller_lines_normalize(unit, lines, MILLER_LINES)); }   // 24 miller_lines_normalize
  (void)b; (void)y2;
#undef COUNT
}
int hs_on_curve(const uint8_t* p1, const uint8_t* q2) {
  return (g1_on_curve(g1_load_be(p1)) ? 1 : 0) | (g2_on_curve(g2_load_be(q2)) ? 2 : 0);
}
}
