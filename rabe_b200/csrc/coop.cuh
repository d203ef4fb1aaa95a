// Lane-paired Fq2 for the pairing kernels (sm_100a): an Fq2 value a + b i is held by TWO adjacent
// lanes of a warp -- the even lane owns a, the odd lane owns b -- so every Fq2/Fq6/Fq12/G2 value of
// a Miller loop or final exponentiation costs each thread half the registers and half the local
// memory of the one-thread layout in tower.cuh, and one pairing keeps two threads busy.  Twice the
// warps fit next to the same cache footprint, which is what the latency-bound one-thread kernels
// lacked (profiles/: 2 resident warps per scheduler at 255 registers, `wait` stalls unhidden).
//
//   add / sub / neg / dbl / mul by Fq   : lane-local
//   mul   (a+bi)(c+di)                  : one exchange of both operands (16 shuffles), then each lane
//                                         computes its component as ONE fused dual product
//                                         re = a*c + (-b)*d,  im = a*d + b*c   (fe_mul2add: two
//                                         products, one Montgomery reduction)
//   sqr                                 : re = (a+b)(a-b), im = (2b)*a -- one product per lane
//   mul by xi = 9+i, conj, inv          : one exchange
//
// The Fq6/Fq12 layer and the Miller loop / final exponentiation are the SAME source as the
// one-thread path (tower_body.inc, pairing_body.inc), compiled here a second time in namespace
// rb::co over this Fp2.  Every routine that exchanges data must be reached by all 32 lanes of the
// warp: kernels keep their control flow warp-uniform (no early exit; exceptional inputs are
// substituted and masked, see k_ac17_dec_miller_pair_co).
// Replaces the same reference call sites as pairing.cuh (`pairing()`, `Gt * Gt`:
// /root/reference/src/schemes/ac17/mod.rs:415-418).
#pragma once
#include "pairing.cuh"

#if defined(__CUDACC__) && !defined(RB_HOST_SIM)
namespace rb {
namespace co {

struct Fp2 { Fp v; };             // this lane's component: real part on even lanes, imaginary on odd lanes

constexpr unsigned FULL = 0xffffffffu;
RB_FN uint32_t lane_im() { return threadIdx.x & 1u; }
RB_FN Fp xchg(const Fp& a) {
  Fp r;
  RB_UNROLL for (int i = 0; i < 8; ++i) r.v[i] = __shfl_xor_sync(FULL, a.v[i], 1);
  return r;
}
// both lanes of the pair agree on a predicate that each evaluated on its own component
RB_FN bool pair_all(bool mine) {
  unsigned m = __ballot_sync(FULL, mine);
  return ((m >> ((threadIdx.x & 31u) & ~1u)) & 3u) == 3u;
}
// component of a one-thread Fq2 constant (both halves are read with a uniform address, then selected)
RB_FN Fp2 pick(const rb::Fp2& c) { return {fe_select(lane_im(), c.a, c.b)}; }
// component of a one-thread Fq2 stored in global memory (only this lane's 32 bytes are read)
RB_FN Fp2 pick_mem(const rb::Fp2& c) { const Fp* p = lane_im() ? &c.b : &c.a; return {*p}; }

RB_FN Fp2 fp2_zero() { return {fe_zero<ModP>()}; }
RB_FN Fp2 fp2_one() { return {fe_select(lane_im(), fe_one<ModP>(), fe_zero<ModP>())}; }
// NOTE: these two vote across the warp -- never put them behind `&&` / `||` / a divergent branch
RB_FN bool fp2_is_zero(const Fp2& x) { return pair_all(fe_is_zero(x.v)); }
RB_FN bool fp2_eq(const Fp2& x, const Fp2& y) { return pair_all(fe_eq(x.v, y.v)); }
RB_FN Fp2 fp2_add(const Fp2& x, const Fp2& y) { return {x.v + y.v}; }
RB_FN Fp2 fp2_sub(const Fp2& x, const Fp2& y) { return {x.v - y.v}; }
RB_FN Fp2 fp2_neg(const Fp2& x) { return {fe_neg(x.v)}; }
RB_FN Fp2 fp2_dbl(const Fp2& x) { return {fe_dbl(x.v)}; }
RB_FN Fp2 fp2_half(const Fp2& x) { return {fe_half(x.v)}; }
RB_FN Fp2 fp2_conj(const Fp2& x) { return {fe_select(lane_im(), x.v, fe_neg(x.v))}; }
#if defined(RB_CO_MULFP_NOINLINE)
static RB_NOINLINE Fp2 fp2_mul_fp_nv(Fp2 x, Fp k) { return {x.v * k}; }       // one shared instance (code size)
RB_FN Fp2 fp2_mul_fp(const Fp2& x, const Fp& k) { return fp2_mul_fp_nv(x, k); }
#else
RB_FN Fp2 fp2_mul_fp(const Fp2& x, const Fp& k) { return {x.v * k}; }
#endif
#if defined(RB_COMPACT)
static RB_NOINLINE Fp2 fp2_mul_xi(Fp2 x) {   // (9a - b) + (9b + a) i
#else
RB_FN Fp2 fp2_mul_xi(const Fp2& x) {   // (9a - b) + (9b + a) i
#endif
  Fp t2 = fe_dbl(x.v), t4 = fe_dbl(t2), t8 = fe_dbl(t4);
  Fp p = xchg(x.v);
  return {t8 + x.v + fe_select(lane_im(), fe_neg(p), p)};
}

// Unreduced helpers for operands that feed exactly one Montgomery product: with both factors below
// 2N the product is below 4N^2/R + N < 2N, so the product's own conditional subtraction restores the
// canonical range and the additions in front of it need none.
RB_FN Fp add_nr(const Fp& a, const Fp& b) {                                                    // a + b        in [0, 2N)
  Fp r;
#if defined(__CUDA_ARCH__)                        // (the carry-chain primitives exist in the device pass only)
  add8(r.v, a.v, b.v);
#else
  r = a + b;
#endif
  return r;
}
RB_FN Fp neg_nr(const Fp& a) {                                                                 // N - a        in (0, N]
  Fp n, r; RB_UNROLL for (int i = 0; i < 8; ++i) n.v[i] = ModP::N(i);
#if defined(__CUDA_ARCH__)
  sub8(r.v, n.v, a.v);
#else
  r = fe_neg(a);
#endif
  return r;
}
RB_FN Fp sub_nr(const Fp& a, const Fp& b) { return add_nr(a, neg_nr(b)); }                      // a - b + N    in (0, 2N)

// out of line like their one-thread counterparts; operands by value (see the note in tower.cuh)
static RB_NOINLINE Fp2 fp2_mul_nv(Fp2 x, Fp2 y) {
  const uint32_t im = lane_im();
  Fp xp = xchg(x.v), yp = xchg(y.v);
  Fp u1 = fe_select(im, x.v, xp);                 // re: a      im: a (the partner's)
  Fp u2 = fe_select(im, neg_nr(xp), x.v);         // re: -b     im: b        (fe_mul2add takes operands <= N)
  return {fe_mul2add(u1, y.v, u2, yp)};           // re: a*c + (-b)*d    im: a*d + b*c
}
static RB_NOINLINE Fp2 fp2_sqr_nv(Fp2 x) {
  const uint32_t im = lane_im();
  Fp xp = xchg(x.v);
  Fp u = add_nr(x.v, fe_select(im, xp, x.v));     // re: a + b  im: 2b       (unreduced, < 2N)
  Fp v = fe_select(im, sub_nr(x.v, xp), xp);      // re: a - b  im: a        (unreduced, < 2N)
  return {u * v};
}
static RB_NOINLINE Fp2 fp2_inv_nv(Fp2 x) {
  Fp s = fe_sqr(x.v);
  Fp n = fe_inv(s + xchg(s));
  Fp r = x.v * n;
  return {fe_select(lane_im(), r, fe_neg(r))};
}
RB_FN Fp2 fp2_mul(const Fp2& x, const Fp2& y) { return fp2_mul_nv(x, y); }
RB_FN Fp2 fp2_sqr(const Fp2& x) { return fp2_sqr_nv(x); }
RB_FN Fp2 fp2_inv(const Fp2& x) { return fp2_inv_nv(x); }
// The same product, inlined: for the few routines whose independent products should be scheduled TOGETHER (the six of
// an Fq6 product, the line products) -- across a call the glue additions cannot overlap the multiply chains.
RB_FN Fp2 fp2_mul_inl(const Fp2& x, const Fp2& y) {
  const uint32_t im = lane_im();
  Fp xp = xchg(x.v), yp = xchg(y.v);
  Fp u1 = fe_select(im, x.v, xp);
  Fp u2 = fe_select(im, neg_nr(xp), x.v);
  return {fe_mul2add(u1, y.v, u2, yp)};
}

typedef rb::Fp2 FullFp2;
typedef rb::MillerLine FullLine;
typedef rb::Affine<Fp2> G2Affine;

#undef RB_K2
#undef RB_KL
#define RB_K2(c) rb::co::pick(c)
#define RB_KL(c) rb::co::pick_mem(c)
#define RB_COOP 1
#undef RB_FP2_MUL_HOT
#if defined(RB_CO_HOT_INLINE)       // A/B switch (tools/co_probe.cu): +2..7 % for a lone batch, -11 % on the pipelined step (instruction footprint) -- off
#define RB_FP2_MUL_HOT(x, y) fp2_mul_inl(x, y)
#else
#define RB_FP2_MUL_HOT(x, y) fp2_mul(x, y)
#endif
#include "tower_body.inc"
#include "pairing_body.inc"
#undef RB_COOP
#undef RB_FP2_MUL_HOT
#define RB_FP2_MUL_HOT(x, y) fp2_mul(x, y)
#undef RB_K2
#undef RB_KL
#define RB_K2(c) (c)
#define RB_KL(c) (c)

}  // namespace co
}  // namespace rb
#endif
