// Pairing kernels over the lane-paired Fq2 of coop.cuh: two adjacent threads per Miller loop /
// final exponentiation.  Thread 2k owns the real parts, thread 2k+1 the imaginary parts of every
// Fq2 value of work item k.  Control flow is warp-uniform (all 32 lanes reach every exchange):
// threads past the end recompute the last item and skip the store; items with a point at
// infinity are detected with a warp vote and handled by substitution + masking.
#pragma once
#include "kernels.cuh"
#include "coop.cuh"

#ifndef RB_CO_BLOCK
#define RB_CO_BLOCK 128      // threads per block of the Miller kernels (= 64 work items)
#endif
#ifndef RB_CO_FE_BLOCK
#define RB_CO_FE_BLOCK RB_CO_BLOCK   // threads per block of the final exponentiation
#endif
#ifndef RB_CO_MINB
#define RB_CO_MINB 1
#endif
#ifndef RB_CO_FE_MINB
#define RB_CO_FE_MINB RB_CO_MINB   // min resident blocks of the final exponentiation (caps its registers)
#endif

namespace rb {

// this lane's half of a canonical G2 point (x.re | x.im | y.re | y.im, 32 bytes each), validated
__device__ __forceinline__ co::G2Affine load_g2_checked_co(const uint8_t* p, int* err, bool* is_inf) {
  const uint32_t im = co::lane_im();
  co::G2Affine a;
  a.x.v = load_fq_checked(p + 32 * im, err);
  a.y.v = load_fq_checked(p + 64 + 32 * im, err);
  const bool inf = co::pair_all(fe_is_zero(a.x.v) && fe_is_zero(a.y.v));      // ONE vote: no short-circuit around a collective
  co::Fp2 lhs = co::fp2_sqr(a.y);
  co::Fp2 rhs = co::fp2_add(co::fp2_mul(co::fp2_sqr(a.x), a.x), co::pick(TWIST_B));
  const bool on = co::fp2_eq(lhs, rhs);
  if (!inf && !on) { flag_error(err, ERR_NOT_MEMBER); a.x.v = fe_zero<ModP>(); a.y.v = fe_zero<ModP>(); }
  *is_inf = inf || !on;
  return a;
}
__device__ __forceinline__ void store_fp12_co(Fp12* dst, const co::Fp12& f) {      // internal Montgomery layout
  const uint32_t im = co::lane_im();
#pragma unroll
  for (int k = 0; k < 6; ++k) { Fp2& d = f12c(*dst, k); Fp* o = im ? &d.b : &d.a; *o = co::f12c(f, k).v; }
}
__device__ __forceinline__ void load_fp12_co(co::Fp12& f, const Fp12* src) {
  const uint32_t im = co::lane_im();
#pragma unroll
  for (int k = 0; k < 6; ++k) { const Fp2& s = f12c(*src, k); const Fp* o = im ? &s.b : &s.a; co::f12c(f, k).v = *o; }
}

// f = miller(pv, q) * miller(pf, fixed Q of `lj`): the shared-accumulator loop when every item of the
// warp has finite points, otherwise (warp vote) both loops on finite stand-ins with the factors of
// missing pairs replaced by one.  f, g: function-scope objects of the calling kernel.
__device__ __forceinline__ void coop_pair_term(co::Fp12& f, co::Fp12& g, G1Affine pv, co::G2Affine q, bool q_inf, G1Affine pf, const MillerLine* lj, bool unit_lines = false) {
  const bool hv = !(aff_is_inf(pv) || q_inf), hf = !aff_is_inf(pf);
  if (__all_sync(co::FULL, hv && hf)) {
    co::miller_pair(&f, &pv, &q, &pf, lj, unit_lines);      // unit_lines: table normalised to l0 = 1 (loaded AC17 keys)
  } else {
    G1Affine gen1; gen1.x = fe_one<ModP>(); gen1.y = fe_dbl(fe_one<ModP>());
    if (!hv) { pv = gen1; q.x = co::pick(G2_GEN_X); q.y = co::pick(G2_GEN_Y); }
    if (!hf) pf = gen1;
    co::miller_single(&f, &pv, &q);
    co::miller_fixed(&g, &pf, lj);
    if (!hv) co::fp12_set_one(f);
    if (!hf) co::fp12_set_one(g);
    co::fp12_mul_to(&f, &f, &g);
  }
}

// work item (b, j), j < 3: e(-(k_p[j] + prod_h_j), c_0[b][j]) * e(prod_g_j, k_0[j]) -- same pairs as
// k_ac17_dec_miller_pair, two threads per item.
__global__ void __launch_bounds__(RB_CO_BLOCK, RB_CO_MINB) k_ac17_dec_miller_pair_co(const G1Affine* __restrict__ ph, int ph_per_item, const G1Affine* __restrict__ pg,
                                                                 const uint8_t* __restrict__ c_0, const MillerLine* __restrict__ lines, int unit_lines, size_t B,
                                                                 Fp12* out, int* err) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = 3 * B;
  size_t t = tid >> 1;
  const bool live = t < n;
  if (!live) t = n - 1;
  const size_t b = t / 3; const int j = (int)(t % 3);
  co::Fp12 f, g;                                  // function scope (pairing_body.inc note)
  G1Affine pv = ph[(ph_per_item ? 3 * b : 0) + j];
  bool q_inf;
  co::G2Affine q = load_g2_checked_co(c_0 + 128 * t, err, &q_inf);
  G1Affine pf = pg[t];
  coop_pair_term(f, g, pv, q, q_inf, pf, lines + (size_t)j * MILLER_LINES, unit_lines != 0);
  if (live) store_fp12_co(out + t, f);
}

// Variant: one work item per ciphertext -- its three terms share ONE accumulator (miller_pair3: 13 %
// fewer Fq products, a third of the threads).  Selected with -DRB_DEC_ITEM=1 (A/B experiment).
__global__ void __launch_bounds__(RB_CO_BLOCK, RB_CO_MINB) k_ac17_dec_miller_item_co(const G1Affine* __restrict__ ph, int ph_per_item, const G1Affine* __restrict__ pg,
                                                                 const uint8_t* __restrict__ c_0, const MillerLine* __restrict__ lines, size_t B,
                                                                 Fp12* out, int* err) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t t = tid >> 1;
  const bool live = t < B;
  if (!live) t = B - 1;
  co::Fp12 f, g, h;                               // function scope (pairing_body.inc note)
  G1Affine pv[3], pf[3];
  co::G2Affine q[3];
  bool q_inf[3], ok = true;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    pv[j] = ph[(ph_per_item ? 3 * t : 0) + j];
    q[j] = load_g2_checked_co(c_0 + 128 * (3 * t + j), err, &q_inf[j]);
    pf[j] = pg[3 * t + j];
    ok = ok && !(aff_is_inf(pv[j]) || q_inf[j]) && !aff_is_inf(pf[j]);
  }
  if (__all_sync(co::FULL, ok)) {
    co::miller_pair3(&f, pv, q, pf, lines);
  } else {
    co::fp12_set_one(f);
#pragma unroll 1
    for (int j = 0; j < 3; ++j) {
      coop_pair_term(h, g, pv[j], q[j], q_inf[j], pf[j], lines + (size_t)j * MILLER_LINES);
      co::fp12_mul_to(&f, &f, &h);
    }
  }
  if (live) store_fp12_co(out + t, f);
}

// Per-leaf decrypt loops whose G2 arguments (and the scalars, moved onto them) belong to the key:
// line tables `lines[i]` are built once per call for the pruned leaves.
struct LeafArgs {
  const uint8_t* ct_g1; size_t ct_g1_item;     // ciphertext-side G1 points [B][ct_g1_item][64]
  const uint8_t* ct_g2; size_t ct_g2_item;     // ciphertext-side G2 points [B][ct_g2_item][128] (pair kernel only)
  const uint32_t* ct_idx;                      // [nI] position of leaf i in the ciphertext arrays (null: i)
  const uint8_t* ks;                           // [nI][64] key-side G1 points, already scaled (pair kernel only)
  const MillerLine* lines;                     // [nI][MILLER_LINES]
  uint32_t nI;
  uint32_t out_stride, out_off;                // Miller value of work item w of item b -> out[b * out_stride + out_off + w]
  const MillerLine* lines_unit;                // [nI][MILLER_LINES] the same tables divided by their l0 (fixed4 kernel; null: none)
  const int* degenerate;                       // != 0 on the device: some l0 was zero, lines_unit is not usable
};
// BSW (bsw/mod.rs:292-293): work item (b, i) = e(ks[i], ctG2[b][ci]) * e(ctG1[b][ci], Q_i fixed)
__global__ void __launch_bounds__(RB_CO_BLOCK, RB_CO_MINB) k_leaf_pair_co(LeafArgs a, size_t B, Fp12* out, int* err) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = B * a.nI;
  size_t t = tid >> 1;
  const bool live = t < n;
  if (!live) t = n - 1;
  const size_t b = t / a.nI; const uint32_t i = (uint32_t)(t % a.nI);
  const uint32_t ci = a.ct_idx ? a.ct_idx[i] : i;
  co::Fp12 f, g;
  G1Affine pv = load_g1_checked(a.ks + 64 * (size_t)i, err);
  bool q_inf;
  co::G2Affine q = load_g2_checked_co(a.ct_g2 + 128 * (b * a.ct_g2_item + ci), err, &q_inf);
  G1Affine pf = load_g1_checked(a.ct_g1 + 64 * (b * a.ct_g1_item + ci), err);
  coop_pair_term(f, g, pv, q, q_inf, pf, a.lines + (size_t)i * MILLER_LINES);
  if (live) store_fp12_co(out + b * a.out_stride + a.out_off + i, f);
}
// fixed-argument pairs only (lsw/mod.rs:276, the e(c, d) of bsw/mod.rs:308): work item (b, w) folds
// leaves 4w .. 4w+3 on one accumulator (miller_fixed4)
__global__ void __launch_bounds__(RB_CO_BLOCK, RB_CO_MINB) k_leaf_fixed4_co(LeafArgs a, size_t B, Fp12* out, int* err) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t chunks = (a.nI + 3) / 4;
  const size_t n = B * chunks;
  size_t t = tid >> 1;
  const bool live = t < n;
  if (!live) t = n - 1;
  const size_t b = t / chunks; const uint32_t w = (uint32_t)(t % chunks);
  co::Fp12 f;
  const bool unit = a.lines_unit != nullptr && *a.degenerate == 0;      // the same for every thread of the launch
  const MillerLine* const tables = unit ? a.lines_unit : a.lines;
  G1Affine p[4]; const MillerLine* ln[4]; bool present[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t i = 4 * w + k;
    const bool in = i < a.nI;
    const uint32_t ii = in ? i : 4 * w;                                   // a valid stand-in (masked)
    const uint32_t ci = a.ct_idx ? a.ct_idx[ii] : ii;
    p[k] = load_g1_checked(a.ct_g1 + 64 * (b * a.ct_g1_item + ci), err);
    ln[k] = tables + (size_t)ii * MILLER_LINES;
    present[k] = in && !aff_is_inf(p[k]);
  }
  co::miller_fixed4(&f, p, ln, present, unit);
  if (live) store_fp12_co(out + b * a.out_stride + a.out_off + w, f);
}

// one pair per two threads -> Miller value (generic pairing products: rb_pairing_product_batch)
__global__ void __launch_bounds__(RB_CO_BLOCK, RB_CO_MINB) k_miller_co(MillerArgs a, size_t n_pairs, Fp12* out, int* err) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t t = tid >> 1;
  const bool live = t < n_pairs;
  if (!live) t = n_pairs - 1;
  size_t pi = a.p_single ? 0 : (a.p_map ? a.p_map[t] : t);
  size_t qi = a.q_map ? a.q_map[t] : (a.q_period ? t % a.q_period : t);
  co::Fp12 f;
  G1Affine p = a.p_mont ? a.p_mont[pi] : load_g1_checked(a.p_bytes + 64 * pi, err);
  bool q_inf;
  co::G2Affine q = load_g2_checked_co(a.q_bytes + 128 * qi, err, &q_inf);
  const bool has = !(aff_is_inf(p) || q_inf);
  if (!has) { p.x = fe_one<ModP>(); p.y = fe_dbl(fe_one<ModP>()); q.x = co::pick(G2_GEN_X); q.y = co::pick(G2_GEN_Y); }   // finite stand-in, masked below
  co::miller_single(&f, &p, &q);
  if (!has) co::fp12_set_one(f);
  if (live) store_fp12_co(out + (a.out_stride ? t * a.out_stride + a.out_off : t), f);
}

// product t: multiply its Miller values, final exponentiation, optional extra Gt factor, canonical store
__global__ void __launch_bounds__(RB_CO_FE_BLOCK, RB_CO_FE_MINB) k_final_exp_co(const Fp12* __restrict__ miller, const uint32_t* __restrict__ offs, uint32_t fixed_count,
                                                      size_t n_products, const uint8_t* __restrict__ extra, uint8_t* __restrict__ out, int* err) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t t = tid >> 1;
  const bool live = t < n_products;
  if (!live) t = n_products - 1;
  const uint32_t im = co::lane_im();
  size_t lo = offs ? offs[t] : t * fixed_count, hi = offs ? offs[t + 1] : (t + 1) * fixed_count;
  co::Fp12 f, r, g;
  co::fp12_set_one(f);
  // warp-uniform trip count: the longest list of the warp; shorter lists multiply by one
  size_t cnt = hi - lo, maxc = cnt;
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) { size_t o = __shfl_xor_sync(co::FULL, maxc, s); maxc = o > maxc ? o : maxc; }
#pragma unroll 1
  for (size_t jj = 0; jj < maxc; ++jj) {
    if (jj < cnt) load_fp12_co(g, miller + lo + jj); else co::fp12_set_one(g);
    co::fp12_mul_to(&f, &f, &g);
  }
  co::final_exponentiation(&r, &f);
  if (extra) {
#pragma unroll 1
    for (int k = 0; k < 6; ++k) co::f12c(g, k).v = load_fq_checked(extra + 384 * t + 64 * k + 32 * im, err);
    co::fp12_mul_to(&r, &r, &g);
  }
  if (live) {
#pragma unroll 1
    for (int k = 0; k < 6; ++k) fe_store_be(out + 384 * t + 64 * k + 32 * im, fe_from_mont(co::f12c(r, k).v));
  }
}

}  // namespace rb
