// Pairing kernels over the six-lane Fq12 of wide.cuh: six lanes per work item, five items per warp (lanes 30 and
// 31 of every warp idle along: they execute every shuffle on clamped stand-in data and never store).  Control flow
// is warp-uniform; absent pairs (a point at infinity) are replaced by valid stand-ins and their lines by one.
#pragma once
#include "kernels.cuh"
#include "wide.cuh"

#ifndef RB_W6_BLOCK
#define RB_W6_BLOCK 128      // 4 warps = 20 work items per block
#endif
#ifndef RB_W6_MINB
#define RB_W6_MINB 1
#endif

namespace rb {

// lane roles of this thread; `item` is clamped into range, `live` tells whether this lane may store
struct W6Slot { w6::Lane L; size_t item; bool live; };
__device__ __forceinline__ W6Slot w6_slot(size_t n_items) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int slot = lane / w6::LANES;                       // 0..4 work items, 5 = the two idle lanes
  W6Slot s;
  s.L.k = lane - slot * w6::LANES;
  s.L.base = slot * w6::LANES;
  s.item = warp * w6::ITEMS_PER_WARP + slot;
  s.live = slot < w6::ITEMS_PER_WARP && s.item < n_items;
  if (s.item >= n_items) s.item = n_items - 1;
  return s;
}
static inline unsigned w6_grid(size_t n_items, unsigned block) {
  const size_t warps = (n_items + w6::ITEMS_PER_WARP - 1) / w6::ITEMS_PER_WARP;
  return (unsigned)((warps * 32 + block - 1) / block);
}

// AC17 cp_decrypt, the Miller half (ac17/mod.rs:415-416): work item = one ciphertext; its three terms
//   e(-(k_p[j] + prod_h_j), c_0[j]) * e(prod_g_j, k_0[j]),  j < 3
// (variable G2 argument c_0[j], fixed argument k_0[j] with precomputed lines) run on ONE accumulator.
__global__ void __launch_bounds__(RB_W6_BLOCK, RB_W6_MINB) k_ac17_dec_item_w6(const G1Affine* __restrict__ ph, int ph_per_item, const G1Affine* __restrict__ pg,
                                                                              const uint8_t* __restrict__ c_0, const MillerLine* __restrict__ lines, int unit_lines, size_t B,
                                                                              Fp12* out, int* err) {
  const W6Slot w = w6_slot(B);
  const int j = w.L.k >> 1;
  G1Affine pv = ph[(ph_per_item ? 3 * w.item : 0) + j];
  G1Affine pf = pg[3 * w.item + j];
  G2Affine q = load_g2_checked(c_0 + 128 * (3 * w.item + j), err);
  w6::PairState s;
  s.has_v = !(aff_is_inf(pv) || aff_is_inf(q));
  s.has_f = !aff_is_inf(pf);
  G1Affine gen1; gen1.x = fe_one<ModP>(); gen1.y = fe_dbl(fe_one<ModP>());
  if (!s.has_v) { pv = gen1; q.x = G2_GEN_X; q.y = G2_GEN_Y; }
  if (!s.has_f) pf = gen1;
  s.t.x = q.x; s.t.y = q.y; s.t.z = fp2_one(); s.qx = q.x; s.qy = q.y;
  s.xv = pv.x; s.yv = pv.y; s.xf = pf.x; s.yf = pf.y;
  s.lines = lines + (size_t)j * MILLER_LINES;
  s.unit_fixed = unit_lines != 0;
  const Fp2 f = w6::miller_terms(w.L, &s, 3);
  if (w.live) f12c(out[w.item], w6::tower_index(w.L.k)) = f;
}

// product t: multiply its Miller values (stored in the tower layout), one final exponentiation, optional extra Gt
// factor (canonical bytes), canonical store.  Same contract as k_final_exp / k_final_exp_co.
__global__ void __launch_bounds__(RB_W6_BLOCK, RB_W6_MINB) k_final_exp_w6(const Fp12* __restrict__ miller, const uint32_t* __restrict__ offs, uint32_t fixed_count,
                                                                          size_t n_products, const uint8_t* __restrict__ extra, uint8_t* __restrict__ out, int* err) {
  const W6Slot w = w6_slot(n_products);
  const size_t t = w.item;
  const int idx = w6::tower_index(w.L.k);
  const size_t lo = offs ? offs[t] : t * fixed_count, hi = offs ? offs[t + 1] : (t + 1) * fixed_count;
  // warp-uniform trip count: the longest list of the warp; shorter lists multiply by one
  uint32_t cnt = (uint32_t)(hi - lo), maxc = cnt;
#pragma unroll
  for (int sft = 16; sft >= 1; sft >>= 1) { uint32_t o = __shfl_xor_sync(0xffffffffu, maxc, sft); maxc = o > maxc ? o : maxc; }
  Fp2 f = w6::one(w.L);
#pragma unroll 1
  for (uint32_t jj = 0; jj < maxc; ++jj) {
    Fp2 g = w6::one(w.L);
    if (jj < cnt) g = f12c(miller[lo + jj], idx);
    f = (jj == 0) ? g : w6::mul(w.L, f, g);
  }
  Fp2 r = w6::final_exponentiation(w.L, f);
  if (extra) {
    Fp2 g;
    g.a = load_fq_checked(extra + 384 * t + 64 * idx, err);
    g.b = load_fq_checked(extra + 384 * t + 64 * idx + 32, err);
    r = w6::mul(w.L, r, g);
  }
  if (w.live) {
    fe_store_be(out + 384 * t + 64 * idx, fe_from_mont(r.a));
    fe_store_be(out + 384 * t + 64 * idx + 32, fe_from_mont(r.b));
  }
}

// test hook (tests/test_gpu_wide.py): out[i] = sum_t (2 x_it)(2 y_it) / R through the wide accumulator, with the
// doubled factors left unreduced (the largest operands the layer feeds), K <= 6 -- pins the PTX carry chains.
__global__ void k_dbg_wide_dot(const uint8_t* __restrict__ xs, const uint8_t* __restrict__ ys, int K, size_t n, uint8_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  w6::WAcc A; w6::wacc_zero(A);
#pragma unroll 1
  for (int t = 0; t < K; ++t) {
    Fp x = fe_to_mont(fe_load_be<ModP>(xs + 32 * (i * K + t))), y = fe_to_mont(fe_load_be<ModP>(ys + 32 * (i * K + t)));
    w6::wacc_mac(A, w6::add_nr(x, x), w6::add_nr(y, y));
  }
  fe_store_be(out + 32 * i, fe_from_mont(w6::wacc_redc<24>(A)));
}
// test hook (tests/test_gpu_wide.py): the dedicated squaring of fp.cuh (fe_sqr_wide).  mode 0: out = a^2 mod p; 1: a^2 mod r (Fr);
// 2: (a + b)^2 mod p with the sum left UNREDUCED (< 2N, the largest operand fe_sqr accepts)
__global__ void k_dbg_fq_sqr(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int mode, size_t n, uint8_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mode == 1) {
    Fr x = fe_to_mont(fe_load_be<ModR>(a + 32 * i));
    fe_store_be(out + 32 * i, fe_from_mont(fe_sqr_wide(x)));
    return;
  }
  Fp x = fe_to_mont(fe_load_be<ModP>(a + 32 * i));
  if (mode == 2) x = w6::add_nr(x, fe_to_mont(fe_load_be<ModP>(b + 32 * i)));
  fe_store_be(out + 32 * i, fe_from_mont(fe_sqr_wide(x)));
}
// test hook: one Fq12 operation per work item through the six-lane layer (op codes of tests/hostsim/wide_sim.cpp:
// 0 mul, 1 sqr, 2 cyclotomic_sqr, 3 inverse, 4 frobenius(arg), 5 conj, 6 final_exponentiation, 7 mul_line)
__global__ void __launch_bounds__(RB_W6_BLOCK) k_dbg_w6_op(int op, int arg, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n, uint8_t* __restrict__ out, int* err) {
  const W6Slot w = w6_slot(n);
  const int idx = w6::tower_index(w.L.k);
  int* const e_ = err;
  Fp2 f, g, o;
  f.a = load_fq_checked(a + 384 * w.item + 64 * idx, e_); f.b = load_fq_checked(a + 384 * w.item + 64 * idx + 32, e_);
  g.a = load_fq_checked(b + 384 * w.item + 64 * idx, e_); g.b = load_fq_checked(b + 384 * w.item + 64 * idx + 32, e_);
  Fp2 l[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) { l[m].a = load_fq_checked(b + 384 * w.item + 64 * m, e_); l[m].b = load_fq_checked(b + 384 * w.item + 64 * m + 32, e_); }
  switch (op) {
    case 0: o = w6::mul(w.L, f, g); break;
    case 1: o = w6::sqr(w.L, f); break;
    case 2: o = w6::cyclotomic_sqr(w.L, f); break;
    case 3: o = w6::inverse(w.L, f); break;
    case 4: o = w6::frobenius(w.L, f, arg); break;
    case 5: o = w6::conj(w.L, f); break;
    case 6: o = w6::final_exponentiation(w.L, f); break;
    default: o = w6::mul_line(w.L, f, l[0], l[1], l[2]); break;
  }
  if (w.live) {
    fe_store_be(out + 384 * w.item + 64 * idx, fe_from_mont(o.a));
    fe_store_be(out + 384 * w.item + 64 * idx + 32, fe_from_mont(o.b));
  }
}

}  // namespace rb
