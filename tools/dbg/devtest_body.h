// shared body: deterministic inputs derived from the generators; writes results as canonical bytes
// into out[slot*384 ...].  Compiled for the host (RB_HOST_SIM) and for the device.
#define SLOT(n) (out + 384 * (n))
DT_FN void run_tests(uint8_t* out) {
  G1Affine p; p.x = fe_one<ModP>(); p.y = fe_dbl(fe_one<ModP>());
  G2Affine q; q.x = G2_GEN_X; q.y = G2_GEN_Y;
  Fp2 a = q.x, b = q.y;
  Fp2 t;
  t = fp2_mul(a, b); fe_store_be(SLOT(0), fe_from_mont(t.a)); fe_store_be(SLOT(0) + 32, fe_from_mont(t.b));
  t = fp2_sqr(a); fe_store_be(SLOT(1), fe_from_mont(t.a)); fe_store_be(SLOT(1) + 32, fe_from_mont(t.b));
  t = fp2_inv(a); fe_store_be(SLOT(2), fe_from_mont(t.a)); fe_store_be(SLOT(2) + 32, fe_from_mont(t.b));
  // an Fq12 element with all coefficients non-trivial
  Fp12 x;
  f12c(x, 0) = a; f12c(x, 1) = b; f12c(x, 2) = fp2_mul(a, b); f12c(x, 3) = fp2_sqr(a); f12c(x, 4) = fp2_sqr(b); f12c(x, 5) = fp2_add(a, b);
  Fp12 y = x; f12c(y, 2) = fp2_mul_xi(a); f12c(y, 4) = fp2_neg(b);
  Fp12 r;
  fp12_store_be(SLOT(3), x);
  Fp6 s6; fp6_mul_p(&s6, &x.h[0], &y.h[1]); r = x; r.h[0] = s6; fp12_store_be(SLOT(4), r);
  fp12_mul_to(&r, &x, &y); fp12_store_be(SLOT(5), r);
  fp12_sqr_to(&r, &x); fp12_store_be(SLOT(6), r);
  fp12_inv_to(&r, &x); fp12_store_be(SLOT(7), r);
  fp12_frobenius_to(&r, &x, 1); fp12_store_be(SLOT(8), r);
  fp12_frobenius_to(&r, &x, 2); fp12_store_be(SLOT(9), r);
  fp12_frobenius_to(&r, &x, 3); fp12_store_be(SLOT(10), r);
  fp12_conj_to(&r, &x); fp12_store_be(SLOT(11), r);
  r = x; fp12_mul_by_line(&r, &a, &b, &f12c(y, 2)); fp12_store_be(SLOT(12), r);
  // easy part puts x into the cyclotomic subgroup
  Fp12 c, ci; fp12_inv_to(&ci, &x); fp12_conj_to(&c, &x); fp12_mul_to(&c, &c, &ci); fp12_frobenius_to(&ci, &c, 2); fp12_mul_to(&c, &ci, &c);
  fp12_store_be(SLOT(13), c);
  fp12_cyclotomic_sqr_to(&r, &c); fp12_store_be(SLOT(14), r);
  fp12_cyclotomic_exp_u_to(&r, &c); fp12_store_be(SLOT(15), r);
  // miller steps
  G2Homog th; th.x = q.x; th.y = q.y; th.z = fp2_one();
  Fp2 l0, l3, l4;
  miller_dbl_step(&th, &l0, &l3, &l4);
  f12c(r, 0) = th.x; f12c(r, 1) = th.y; f12c(r, 2) = th.z; f12c(r, 3) = l0; f12c(r, 4) = l3; f12c(r, 5) = l4; fp12_store_be(SLOT(16), r);
  miller_add_step(&th, &q.x, &q.y, &l0, &l3, &l4);
  f12c(r, 0) = th.x; f12c(r, 1) = th.y; f12c(r, 2) = th.z; f12c(r, 3) = l0; f12c(r, 4) = l3; f12c(r, 5) = l4; fp12_store_be(SLOT(17), r);
  Fp12 f; miller_single(&f, &p, &q); fp12_store_be(SLOT(18), f);
  final_exponentiation(&r, &f); fp12_store_be(SLOT(19), r);
  final_exponentiation(&r, &x); fp12_store_be(SLOT(20), r);
  MillerLine lines[MILLER_LINES];
  miller_lines_for(lines, &q);
  Fp12 f2; miller_fixed(&f2, &p, lines); fp12_store_be(SLOT(21), f2);   // must equal slot 18
  { int idxs[4] = {0, 3, 40, MILLER_LINES - 1};
    for (int k = 0; k < 4; ++k) { fp12_set_one(r); f12c(r, 0) = lines[idxs[k]].l0; f12c(r, 1) = lines[idxs[k]].l3; f12c(r, 2) = lines[idxs[k]].l4; fp12_store_be(SLOT(22 + k), r); } }
}
#define N_SLOTS 26
