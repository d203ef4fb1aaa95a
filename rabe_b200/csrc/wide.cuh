// Six-lane Fq12 for the pairing kernels (sm_100a): one work item (a ciphertext's Miller product, a final
// exponentiation) is owned by SIX lanes of a warp; lane k holds the Fq2 coefficient of w^k of every Fq12 value
// in the flat basis Fq12 = Fq2[w]/(w^6 - xi) (tower order c0.c0 c0.c1 c0.c2 c1.c0 c1.c1 c1.c2 = w^0 w^2 w^4 w^1
// w^3 w^5).  Five items per warp (lanes 30, 31 idle along).  Everything lives in registers -- an Fq12 value is 16
// registers per lane, there is no per-thread Fq12 in local memory -- and the lanes exchange operands with warp
// shuffles.
//
//   product      c_k = sum_{i+j=k} f_i g_j + xi sum_{i+j=k+6} f_i g_j : six Fq2 products per lane, accumulated
//                UNREDUCED (Karatsuba: three 512-bit sums P = sum a c, Q = sum b d, S = sum (a+b)(c+d)) and
//                reduced once per sum: 18 limb products + 3 Montgomery reductions per lane.
//   line product three Fq2 products per lane (the line l0 + l3 w^3 + l4 w^4 has three coefficients)
//   cyclotomic squaring, Frobenius, conjugation: two products / one product / none per lane
//
// A Miller step keeps the G2 points of an item's pairs on lane PAIRS (lanes 2j, 2j+1 walk pair j: the independent
// Fq2 products of a doubling / addition step are split between the two lanes and exchanged), so a ciphertext's
// three decrypt terms share ONE accumulator f: one f^2 per step for all six pairings.
//
// Replaces the same reference call sites as pairing.cuh (`pairing()`, `Gt * Gt`, `Gt.pow`:
// /root/reference/src/schemes/ac17/mod.rs:415-418, bsw/mod.rs:291-308, lsw/mod.rs:275-280, aw11/mod.rs:340-350).
//
// The header also compiles for the host (RB_HOST_SIM, tests/hostsim/wide_sim.cpp): there a lane is a host thread
// and a shuffle is a barrier-protected exchange, so the whole layer is checked against the one-thread tower
// without a GPU; the PTX carry chains are checked on the device (tests/test_gpu_wide.py).
#pragma once
#include "pairing.cuh"

namespace rb {
namespace w6 {

constexpr int LANES = 6;
#ifndef RB_W6_MUL_UNROLL
#define RB_W6_MUL_UNROLL 2   // rounds of a product pass unrolled together (6 = straight-line code)
#endif
constexpr int MUL_UNROLL = RB_W6_MUL_UNROLL;
constexpr int ITEMS_PER_WARP = 5;

// ------------------------------------------------------------------------------------------ lanes
#if defined(RB_HOST_SIM)
struct Team;                                   // host harness: 6 threads + an exchange buffer
struct Lane { int k; Team* team; };
Fp team_exchange(const Lane& L, const Fp& mine, int src);       // the value lane `src` passed at the same call site
uint32_t team_exchange_u32(const Lane& L, uint32_t mine, int src);
#else
struct Lane { int k; int base; };              // k: lane of the item (0..5); base: first warp lane of the item
#endif

RB_FN Fp shfl_fp(const Lane& L, const Fp& a, int src) {
#if defined(RB_HOST_SIM)
  return team_exchange(L, a, src);
#else
  Fp r;
  RB_UNROLL for (int i = 0; i < 8; ++i) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], L.base + src);
  return r;
#endif
}
RB_FN uint32_t shfl_u32(const Lane& L, uint32_t v, int src) {
#if defined(RB_HOST_SIM)
  return team_exchange_u32(L, v, src);
#else
  return __shfl_sync(0xffffffffu, v, L.base + src);
#endif
}
RB_FN Fp2 shfl_fp2(const Lane& L, const Fp2& a, int src) { return {shfl_fp(L, a.a, src), shfl_fp(L, a.b, src)}; }
RB_FN Fp sel(bool c, const Fp& a, const Fp& b) { return fe_select<ModP>(c ? 1u : 0u, a, b); }      // c ? b : a
RB_FN Fp2 sel2(bool c, const Fp2& a, const Fp2& b) { return {sel(c, a.a, b.a), sel(c, a.b, b.b)}; }

// a + b without the conditional subtraction (< 2N): operand of exactly one wide product
RB_FN Fp add_nr(const Fp& a, const Fp& b) {
  Fp r;
#if defined(__CUDA_ARCH__)
  add8(r.v, a.v, b.v);
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; ++i) { uint64_t x = (uint64_t)a.v[i] + b.v[i] + c; r.v[i] = (uint32_t)x; c = x >> 32; }
#endif
  return r;
}

// ------------------------------------------------------------------------------------------ wide sums
// Unreduced sum of products.  Device: two 16-limb accumulators -- E takes the partial products that start at an
// even limb, O (one limb up) those that start at an odd limb, so every 32x32->64 product is a mad.lo.cc/madc.hi.cc
// pair on an aligned register pair (one IMAD.WIDE) -- plus one deferred-carry counter per chain end: a chain adds
// into limbs that already hold data, so its carry-out is COUNTED (cE / cO) instead of rippled, and folded in once
// per sum.  Bound: up to 6 products of factors < 2N stay below 24 N^2 < 2^512.
struct WAcc {
#if defined(__CUDA_ARCH__)
  uint32_t E[16], O[16];
  uint32_t cE[5], cO[4];       // cE[m]: carries into limb 8+2m ; cO[m]: carries into O index 8+2m (= limb 9+2m)
#else
  uint32_t t[17];
#endif
};

RB_FN void wacc_zero(WAcc& A) {
#if defined(__CUDA_ARCH__)
  RB_UNROLL for (int i = 0; i < 16; ++i) { A.E[i] = 0; A.O[i] = 0; }
  RB_UNROLL for (int i = 0; i < 5; ++i) A.cE[i] = 0;
  RB_UNROLL for (int i = 0; i < 4; ++i) A.cO[i] = 0;
#else
  for (int i = 0; i < 17; ++i) A.t[i] = 0;
#endif
}

// A += x * y   (plain integers: the Montgomery forms as they are)
RB_FN void wacc_mac(WAcc& A, const Fp& x, const Fp& y) {
#if defined(__CUDA_ARCH__)
  const uint32_t* a = x.v;
  RB_UNROLL for (int i = 0; i < 8; ++i) {
    const uint32_t b = y.v[i];
    if ((i & 1) == 0) {
      // x_even * b starts at limb i (even) -> E[i..i+7], carry into limb i+8 ; x_odd * b starts at limb i+1 -> O[i..i+7], carry into O index i+8
      mad_even(A.E + i, A.cE[i / 2], a[0], a[2], a[4], a[6], b);
      mad_even(A.O + i, A.cO[i / 2], a[1], a[3], a[5], a[7], b);
    } else {
      // x_even * b starts at limb i (odd) -> O[i-1..i+6], carry into O index i+7 ; x_odd * b starts at limb i+1 (even) -> E[i+1..i+8], carry into limb i+9
      mad_even(A.O + i - 1, A.cO[(i - 1) / 2], a[0], a[2], a[4], a[6], b);
      mad_even(A.E + i + 1, A.cE[(i + 1) / 2], a[1], a[3], a[5], a[7], b);
    }
  }
#else
#if defined(RB_HOST_SIM)
  ++g_host_mul_count;
#endif
  uint32_t p[16];
  for (int i = 0; i < 16; ++i) p[i] = 0;
  for (int i = 0; i < 8; ++i) {
    uint64_t c = 0;
    for (int j = 0; j < 8; ++j) { uint64_t s = (uint64_t)x.v[j] * y.v[i] + p[i + j] + c; p[i + j] = (uint32_t)s; c = s >> 32; }
    p[i + 8] = (uint32_t)c;
  }
  uint64_t c = 0;
  for (int i = 0; i < 16; ++i) { uint64_t s = (uint64_t)A.t[i] + p[i] + c; A.t[i] = (uint32_t)s; c = s >> 32; }
  A.t[16] += (uint32_t)c;
#endif
}

#if defined(__CUDA_ARCH__)
// Role swap of the two Montgomery accumulators at a row boundary when the row adds no product (y_i = 0): the old
// offset-0 accumulator O (limb 0 already zero) moves two limbs down and becomes the offset-1 accumulator; its
// orphaned limb O[1] is folded into F0.  (mad_odd_swap of fp.cuh with the products removed.)
RB_FN void redc_swap(uint32_t& F0, uint32_t* Sn, const uint32_t* O) {
  asm("add.cc.u32 %0,%0,%9; addc.cc.u32 %1,%10,0; addc.cc.u32 %2,%11,0; addc.cc.u32 %3,%12,0; addc.cc.u32 %4,%13,0;"
      "addc.cc.u32 %5,%14,0; addc.cc.u32 %6,%15,0; addc.u32 %7,0,0; mov.u32 %8,0;"
      : "+r"(F0), "=r"(Sn[0]), "=r"(Sn[1]), "=r"(Sn[2]), "=r"(Sn[3]), "=r"(Sn[4]), "=r"(Sn[5]), "=r"(Sn[6]), "=r"(Sn[7])
      : "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
}
#endif

// lo / R mod N for any 256-bit lo: in [0, N] (REDUCE = false: the caller folds the value into a larger sum that it
// reduces itself) or in [0, N)
template <bool REDUCE>
RB_FN Fp redc8(const uint32_t* lo) {
  Fp r;
#if defined(__CUDA_ARCH__)
  typedef ModP M;
  uint32_t A[8], B[8];
  RB_UNROLL for (int k = 0; k < 8; ++k) { A[k] = lo[k]; B[k] = 0; }
  {
    const uint32_t m = A[0] * M::INV;
    mad_odd(B, M::N(1), M::N(3), M::N(5), M::N(7), m);
    mad_even(A, B[7], M::N(0), M::N(2), M::N(4), M::N(6), m);
  }
  RB_UNROLL for (int i = 1; i < 8; ++i) {
    uint32_t S[8];
    if (i & 1) {
      redc_swap(B[0], S, A);
      const uint32_t m = B[0] * M::INV;
      mad_odd(S, M::N(1), M::N(3), M::N(5), M::N(7), m);
      mad_even(B, S[7], M::N(0), M::N(2), M::N(4), M::N(6), m);
      RB_UNROLL for (int k = 0; k < 8; ++k) A[k] = S[k];
    } else {
      redc_swap(A[0], S, B);
      const uint32_t m = A[0] * M::INV;
      mad_odd(S, M::N(1), M::N(3), M::N(5), M::N(7), m);
      mad_even(A, S[7], M::N(0), M::N(2), M::N(4), M::N(6), m);
      RB_UNROLL for (int k = 0; k < 8; ++k) B[k] = S[k];
    }
  }
  asm("add.cc.u32 %0,%0,%8; addc.cc.u32 %1,%1,%9; addc.cc.u32 %2,%2,%10; addc.cc.u32 %3,%3,%11;"
      "addc.cc.u32 %4,%4,%12; addc.cc.u32 %5,%5,%13; addc.cc.u32 %6,%6,%14; addc.u32 %7,%7,0;"
      : "+r"(A[0]), "+r"(A[1]), "+r"(A[2]), "+r"(A[3]), "+r"(A[4]), "+r"(A[5]), "+r"(A[6]), "+r"(A[7])
      : "r"(B[1]), "r"(B[2]), "r"(B[3]), "r"(B[4]), "r"(B[5]), "r"(B[6]), "r"(B[7]));
  RB_UNROLL for (int k = 0; k < 8; ++k) r.v[k] = A[k];
#else
  uint32_t t[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < 8; ++k) t[k] = lo[k];
  for (int i = 0; i < 8; ++i) {
    uint32_t m = t[0] * ModP::INV;
    uint64_t s = (uint64_t)m * ModP::N(0) + t[0], c = s >> 32;
    for (int j = 1; j < 8; ++j) { s = (uint64_t)m * ModP::N(j) + t[j] + c; t[j - 1] = (uint32_t)s; c = s >> 32; }
    s = (uint64_t)t[8] + c; t[7] = (uint32_t)s; t[8] = (uint32_t)(s >> 32);
  }
  for (int k = 0; k < 8; ++k) r.v[k] = t[k];
#endif
  if (REDUCE) fe_reduce_once<ModP>(r.v);
  return r;
}

// v (9 limbs) -= c * 2^0 when v >= c, for the 9-limb constant k*N
RB_FN void sub9_if_geq(uint32_t* v, int k) {
  uint32_t c[9]; uint64_t cy = 0;
  RB_UNROLL for (int i = 0; i < 8; ++i) { uint64_t s = (uint64_t)ModP::N(i) * (uint32_t)k + cy; c[i] = (uint32_t)s; cy = s >> 32; }
  c[8] = (uint32_t)cy;
  uint32_t d[9];
#if defined(__CUDA_ARCH__)
  uint32_t borrow = sub8(d, v, c);                  // all-ones when v[0..7] < c[0..7]
  const uint32_t hi = v[8] - c[8] - (borrow & 1u);
  const bool neg = (v[8] < c[8]) || (v[8] == c[8] && borrow);
  d[8] = hi;
  RB_UNROLL for (int i = 0; i < 9; ++i) v[i] = neg ? v[i] : d[i];
#else
  uint64_t br = 0;
  for (int i = 0; i < 9; ++i) { uint64_t x = (uint64_t)v[i] - c[i] - br; d[i] = (uint32_t)x; br = (x >> 32) & 1; }
  if (!br) for (int i = 0; i < 9; ++i) v[i] = d[i];
#endif
}

#if defined(__CUDA_ARCH__)
// 8-limb add / subtract with the carry (borrow) passed through registers: every carry chain lives in ONE asm block
// (the flag is implicit state that the compiler does not track across statements).
RB_FN uint32_t addc8(uint32_t* r, const uint32_t* a, const uint32_t* b, uint32_t cin) {
  uint32_t cout;
  asm("{ .reg .u32 t; add.cc.u32 t,%25,0xffffffff; addc.cc.u32 %0,%9,%17; addc.cc.u32 %1,%10,%18; addc.cc.u32 %2,%11,%19; addc.cc.u32 %3,%12,%20;"
      "addc.cc.u32 %4,%13,%21; addc.cc.u32 %5,%14,%22; addc.cc.u32 %6,%15,%23; addc.cc.u32 %7,%16,%24; addc.u32 %8,0,0; }"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(cout)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(cin));
  return cout;
}
// x += b + cin in place
RB_FN uint32_t addc8_ip(uint32_t* x, const uint32_t* b, uint32_t cin) {
  uint32_t cout;
  asm("{ .reg .u32 t; add.cc.u32 t,%17,0xffffffff; addc.cc.u32 %0,%0,%9; addc.cc.u32 %1,%1,%10; addc.cc.u32 %2,%2,%11; addc.cc.u32 %3,%3,%12;"
      "addc.cc.u32 %4,%4,%13; addc.cc.u32 %5,%5,%14; addc.cc.u32 %6,%6,%15; addc.cc.u32 %7,%7,%16; addc.u32 %8,0,0; }"
      : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "=&r"(cout)
      : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(cin));
  return cout;
}
// x -= b + bin in place, returns the borrow (0 / 1)
RB_FN uint32_t subc8_ip(uint32_t* x, const uint32_t* b, uint32_t bin) {
  uint32_t bout;
  asm("{ .reg .u32 t; sub.cc.u32 t,0,%17; subc.cc.u32 %0,%0,%9; subc.cc.u32 %1,%1,%10; subc.cc.u32 %2,%2,%11; subc.cc.u32 %3,%3,%12;"
      "subc.cc.u32 %4,%4,%13; subc.cc.u32 %5,%5,%14; subc.cc.u32 %6,%6,%15; subc.cc.u32 %7,%7,%16; subc.u32 %8,0,0; }"
      : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "=&r"(bout)
      : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(bin));
  return bout & 1u;
}
RB_FN uint32_t subc8(uint32_t* r, const uint32_t* a, const uint32_t* b, uint32_t bin) {
  uint32_t bout;
  asm("{ .reg .u32 t; sub.cc.u32 t,0,%25; subc.cc.u32 %0,%9,%17; subc.cc.u32 %1,%10,%18; subc.cc.u32 %2,%11,%19; subc.cc.u32 %3,%12,%20;"
      "subc.cc.u32 %4,%13,%21; subc.cc.u32 %5,%14,%22; subc.cc.u32 %6,%15,%23; subc.cc.u32 %7,%16,%24; subc.u32 %8,0,0; }"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(bout)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(bin));
  return bout & 1u;
}
#endif

// the accumulated sum as one 512-bit integer (16 limbs)
struct Wide { uint32_t t[16]; };
RB_FN Wide wacc_merge(const WAcc& A) {
  Wide w;
  uint32_t* t = w.t;
#if defined(__CUDA_ARCH__)
  // t = E + (O << 32): limb m takes E[m] + O[m-1]; then the deferred carries: limb 8 += cE0, 9 += cO0, ... 15 += cO3
  // (cE4, O[15] and the carry out of limb 15 are zero by the bound)
  t[0] = A.E[0];
  uint32_t hi[8], zero = 0;
  const uint32_t c1 = addc8(t + 1, A.E + 1, A.O, zero);
  const uint32_t e2[8] = {A.E[9], A.E[10], A.E[11], A.E[12], A.E[13], A.E[14], A.E[15], 0};
  addc8(hi, e2, A.O + 8, c1);
  RB_UNROLL for (int i = 0; i < 7; ++i) t[9 + i] = hi[i];
  const uint32_t cc[8] = {A.cE[0], A.cO[0], A.cE[1], A.cO[1], A.cE[2], A.cO[2], A.cE[3], A.cO[3]};
  addc8_ip(t + 8, cc, zero);
#else
  for (int i = 0; i < 16; ++i) t[i] = A.t[i];
#endif
  return w;
}

// x -= y over 16 limbs (mod 2^512)
RB_FN void wide_sub(Wide& x, const Wide& y) {
#if defined(__CUDA_ARCH__)
  const uint32_t b = subc8_ip(x.t, y.t, 0u);
  subc8_ip(x.t + 8, y.t + 8, b);
#else
  uint64_t br = 0;
  for (int i = 0; i < 16; ++i) { uint64_t v = (uint64_t)x.t[i] - y.t[i] - br; x.t[i] = (uint32_t)v; br = (v >> 32) & 1; }
#endif
}
// x += k * N * 2^256  (keeps a difference of wide sums non-negative; a multiple of N, so invisible after reduction)
RB_FN void wide_add_hiN(Wide& x, int k) {
  uint64_t c = 0, cy = 0;
  RB_UNROLL for (int i = 0; i < 8; ++i) {
    uint64_t kn = (uint64_t)ModP::N(i) * (uint32_t)k + cy; cy = kn >> 32;
    uint64_t v = (uint64_t)x.t[8 + i] + (uint32_t)kn + c; x.t[8 + i] = (uint32_t)v; c = v >> 32;
  }
}

// w / R mod N, fully reduced, for a 512-bit w below BOUND * N^2:  w = hi 2^256 + lo,  w / R = hi + lo / R  (mod N),
// and hi + lo / R + N < (0.19 BOUND + 1) N decides how many conditional subtractions the tail needs.
template <int BOUND>
RB_FN Fp wide_redc(const Wide& w) {
  const uint32_t* t = w.t;
  Fp r = redc8<false>(t);
  uint32_t v[9];
#if defined(__CUDA_ARCH__)
  v[8] = addc8(v, t + 8, r.v, 0u);
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; ++i) { uint64_t s = (uint64_t)t[8 + i] + r.v[i] + c; v[i] = (uint32_t)s; c = s >> 32; }
  v[8] = (uint32_t)c;
#endif
  static_assert(BOUND >= 1 && BOUND <= 36, "wide sum bound");
  if (BOUND * 19 + 100 > 400) sub9_if_geq(v, 4);
  if (BOUND * 19 + 100 > 200) sub9_if_geq(v, 2);
  sub9_if_geq(v, 1);
  RB_UNROLL for (int i = 0; i < 8; ++i) r.v[i] = v[i];
  return r;
}
template <int BOUND> RB_FN Fp wacc_redc(const WAcc& A) { return wide_redc<BOUND>(wacc_merge(A)); }

// Karatsuba recombination in the WIDE domain (two reductions instead of three): with P = sum a c, Q = sum b d (factors
// < N, K terms) and S = sum (a+b)(c+d):   re = P - Q + kN 2^256  (kN 2^256 >= Q keeps it non-negative),  im = S - P - Q >= 0.
template <int K>
RB_FN Fp2 kara_wide(Wide P, const Wide& Q, Wide S) {
  constexpr int KN = (K * 19 + 99) / 100;              // ceil(0.19 K): KN * N * 2^256 >= K N^2 > Q
  wide_sub(S, P); wide_sub(S, Q);                      // S - P - Q = sum (a d + b c) < 2 K N^2
  wide_sub(P, Q); wide_add_hiN(P, KN);                 // < K N^2 + KN N 2^256 <= (K + 5.3 KN) N^2
  return {wide_redc<K + 6 * KN>(P), wide_redc<2 * K>(S)};
}

// ------------------------------------------------------------------------------------------ Fq2 products over wide sums
// An Fq2 dot product sum_t x_t y_t is three real ones (Karatsuba): P = sum a c, Q = sum b d, S = sum (a+b)(c+d) with
// the sums a+b, c+d left unreduced; re = P - Q, im = S - P - Q.  Each real dot product runs on ONE wide accumulator
// (41 registers) -- the three passes run one after the other so that the kernels stay far below the register limit.
struct Trip { Fp a, b, s; };                         // an Fq2 operand prepared for the three passes: re, im, re + im (unreduced)
RB_FN Trip trip(const Fp2& x) { return {x.a, x.b, add_nr(x.a, x.b)}; }
RB_FN Fp2 kara(const Fp& p, const Fp& q, const Fp& s) { return {p - q, s - p - q}; }

// f * g.  Lane k: c_k = sum_t f_t * G_t,  G_t = g_{k-t} for t <= k, xi * g_{k-t+6} for t > k.  In round t lane j's g is
// read by exactly one lane -- k = (j + t) mod 6, wrapped iff j + t >= 6 -- so the SOURCE lane picks the version.
RB_FN Wide mul_pass(const Lane& L, const Fp& x, const Fp& y, const Fp& xy) {
  WAcc A; wacc_zero(A);
#if !defined(RB_HOST_SIM)
#pragma unroll MUL_UNROLL
#endif
  for (int t = 0; t < LANES; ++t) {
    int j = L.k - t; if (j < 0) j += LANES;
    const Fp xt = shfl_fp(L, x, t);
    const Fp yt = shfl_fp(L, sel(L.k + t >= LANES, y, xy), j);
    wacc_mac(A, xt, yt);
  }
  return wacc_merge(A);
}
static RB_NOINLINE Fp2 mul(Lane L, Fp2 f, Fp2 g) {
  const Trip F = trip(f), G = trip(g), X = trip(fp2_mul_xi(g));
  const Wide p = mul_pass(L, F.a, G.a, X.a);
  const Wide q = mul_pass(L, F.b, G.b, X.b);
  const Wide s = mul_pass(L, F.s, G.s, X.s);
  return kara_wide<6>(p, q, s);
}
RB_FN Fp2 sqr(const Lane& L, const Fp2& f) { return mul(L, f, f); }

// x0 y0 + x1 y1 + x2 y2 over Fq2, all operands local
RB_FN Wide dot3_pass(const Fp& x0, const Fp& y0, const Fp& x1, const Fp& y1, const Fp& x2, const Fp& y2) {
  WAcc A; wacc_zero(A);
  wacc_mac(A, x0, y0); wacc_mac(A, x1, y1); wacc_mac(A, x2, y2);
  return wacc_merge(A);
}
RB_FN Fp2 dot3(const Fp2& x0, const Fp2& y0, const Fp2& x1, const Fp2& y1, const Fp2& x2, const Fp2& y2) {
  const Wide p = dot3_pass(x0.a, y0.a, x1.a, y1.a, x2.a, y2.a);
  const Wide q = dot3_pass(x0.b, y0.b, x1.b, y1.b, x2.b, y2.b);
  const Wide s = dot3_pass(add_nr(x0.a, x0.b), add_nr(y0.a, y0.b), add_nr(x1.a, x1.b), add_nr(y1.a, y1.b), add_nr(x2.a, x2.b), add_nr(y2.a, y2.b));
  return kara_wide<3>(p, q, s);
}
RB_FN Wide dot2_pass(const Fp& x0, const Fp& y0, const Fp& x1, const Fp& y1) {
  WAcc A; wacc_zero(A);
  wacc_mac(A, x0, y0); wacc_mac(A, x1, y1);
  return wacc_merge(A);
}
RB_FN Fp2 dot2(const Fp2& x0, const Fp2& y0, const Fp2& x1, const Fp2& y1) {
  const Wide p = dot2_pass(x0.a, y0.a, x1.a, y1.a);
  const Wide q = dot2_pass(x0.b, y0.b, x1.b, y1.b);
  const Wide s = dot2_pass(add_nr(x0.a, x0.b), add_nr(y0.a, y0.b), add_nr(x1.a, x1.b), add_nr(y1.a, y1.b));
  return kara_wide<2>(p, q, s);
}

// f * (y0 + y1 w^j1 + y2 w^j2), 0 < j1 < j2 < 6, with the three coefficients known to every lane (lines: j = 3, 4;
// Fq6 elements: j = 2, 4).  Lane k: c_k = f_k y0 + f_{k-j1} Y1 + f_{k-j2} Y2 with xi on the wrapped terms.
template <int J1, int J2>
static RB_NOINLINE Fp2 mul_sparse(Lane L, Fp2 f, Fp2 y0, Fp2 y1, Fp2 y2) {
  const bool w1 = L.k < J1, w2 = L.k < J2;
  const Fp2 f1 = shfl_fp2(L, f, w1 ? L.k - J1 + LANES : L.k - J1), f2 = shfl_fp2(L, f, w2 ? L.k - J2 + LANES : L.k - J2);
  const Fp2 z1 = sel2(w1, y1, fp2_mul_xi(y1)), z2 = sel2(w2, y2, fp2_mul_xi(y2));
  return dot3(f, y0, f1, z1, f2, z2);
}
RB_FN Fp2 mul_line(const Lane& L, const Fp2& f, const Fp2& l0, const Fp2& l3, const Fp2& l4) { return mul_sparse<3, 4>(L, f, l0, l3, l4); }

RB_FN Fp2 one(const Lane& L) { return sel2(L.k == 0, fp2_zero(), fp2_one()); }
// f^(p^6) = conjugation over Fq6: w -> -w
RB_FN Fp2 conj(const Lane& L, const Fp2& f) { return sel2((L.k & 1) != 0, f, fp2_neg(f)); }
// f^(p^j), j = 1, 2, 3: coefficient of w^k times xi^(k (p^j - 1)/6), conjugated for odd j
RB_FN Fp2 frobenius(const Lane& L, const Fp2& f, int j) {
  const FullFp2* g = (j == 1) ? FROB1 : ((j == 2) ? FROB2 : FROB3);
  Fp2 z = (j & 1) ? fp2_conj(f) : f;
  return fp2_mul(z, g[L.k]);
}

// Granger-Scott squaring in the cyclotomic subgroup.  With z_k the coefficient of w^k, the three pairs (z_k, z_{k+3}),
// k = 0, 1, 2, are Fq4 elements u + v s (s = w^3, s^2 = xi):  A_k = u^2 + xi v^2,  B_k = 2 u v, and
//   w^0: 3 A_0 - 2 z_0   w^3: 3 B_0 + 2 z_3   w^1: 3 xi B_2 + 2 z_1   w^4: 3 A_2 - 2 z_4   w^2: 3 A_1 - 2 z_2   w^5: 3 B_1 + 2 z_5
static RB_NOINLINE Fp2 cyclotomic_sqr(Lane L, Fp2 f) {
  // which Fq4 pair feeds this lane: lanes 0,3 <- pair 0 ; lanes 2,5 <- pair 1 ; lanes 1,4 <- pair 2
  const int k = L.k;
  const int pr = (k == 0 || k == 3) ? 0 : ((k == 2 || k == 5) ? 1 : 2);
  const bool wantA = (k == 0 || k == 2 || k == 4);
  const Fp2 u = shfl_fp2(L, f, pr), v = shfl_fp2(L, f, pr + 3);
  // A = u*u + v*(xi v) ; B = u*v (doubled afterwards): one code path, operands selected
  const Fp2 xv = fp2_mul_xi(v);
  const Fp2 y0 = sel2(wantA, v, u), x1 = sel2(wantA, fp2_zero(), v);
  const Wide p = dot2_pass(u.a, y0.a, x1.a, xv.a);
  const Wide q = dot2_pass(u.b, y0.b, x1.b, xv.b);
  const Wide sm = dot2_pass(add_nr(u.a, u.b), add_nr(y0.a, y0.b), add_nr(x1.a, x1.b), add_nr(xv.a, xv.b));
  Fp2 r = kara_wide<2>(p, q, sm);
  if (!wantA) r = fp2_dbl(r);
  const Fp2 rx = fp2_mul_xi(r);
  r = sel2(k == 1, r, rx);
  const Fp2 r3 = fp2_add(fp2_dbl(r), r), z2 = fp2_dbl(f);
  return sel2(wantA, fp2_add(r3, z2), fp2_sub(r3, z2));
}

// 1 / f:  f^-1 = conj(f) / (f conj(f)), the norm lies in Fq6 = even coefficients; its inverse is computed by every
// lane from the three broadcast coefficients (one Fq inversion per lane, all alike).
static RB_NOINLINE Fp2 inverse(Lane L, Fp2 f) {
  const Fp2 cf = conj(L, f);
  const Fp2 n = mul(L, f, cf);
  const Fp2 n0 = shfl_fp2(L, n, 0), n1 = shfl_fp2(L, n, 2), n2 = shfl_fp2(L, n, 4);
  const Fp2 t0 = fp2_sub(fp2_sqr(n0), fp2_mul_xi(fp2_mul(n1, n2)));
  const Fp2 t1 = fp2_sub(fp2_mul_xi(fp2_sqr(n2)), fp2_mul(n0, n1));
  const Fp2 t2 = fp2_sub(fp2_sqr(n1), fp2_mul(n0, n2));
  const Fp2 d = fp2_inv(fp2_add(fp2_mul(n0, t0), fp2_mul_xi(fp2_add(fp2_mul(n2, t1), fp2_mul(n1, t2)))));
  return mul_sparse<2, 4>(L, cf, fp2_mul(t0, d), fp2_mul(t1, d), fp2_mul(t2, d));
}

// x^u for the BN parameter u (63 bits), x in the cyclotomic subgroup: width-3 signed digits (tower_body.inc), a negative
// digit multiplies by the conjugate
static RB_NOINLINE Fp2 cyclotomic_exp_u(Lane L, Fp2 x) {
  const Fp2 x3 = mul(L, cyclotomic_sqr(L, x), x);
  Fp2 r = (U_WNAF[U_WNAF_LEN - 1] == 3) ? x3 : x;
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
  for (int i = U_WNAF_LEN - 2; i >= 0; --i) {
    r = cyclotomic_sqr(L, r);
    const int d = U_WNAF[i];
    if (d != 0) {
      Fp2 s = (d == 1 || d == -1) ? x : x3;
      if (d < 0) s = conj(L, s);
      r = mul(L, r, s);
    }
  }
  return r;
}
RB_FN Fp2 exp_neg_u(const Lane& L, const Fp2& x) { return conj(L, cyclotomic_exp_u(L, x)); }

// f^((p^6-1)(p^2+1)) then the lineage's hard part (same addition chain as pairing_body.inc)
static RB_NOINLINE Fp2 final_exponentiation(Lane L, Fp2 in) {
  Fp2 t = inverse(L, in);
  Fp2 a = mul(L, conj(L, in), t);                         // f^(p^6-1)
  const Fp2 x = mul(L, frobenius(L, a, 2), a);            // ^(p^2+1)
  a = exp_neg_u(L, x);                                    // A = x^-u
  const Fp2 b = cyclotomic_sqr(L, a);                     // B = A^2
  a = cyclotomic_sqr(L, b);                               // C = B^2
  Fp2 d = mul(L, a, b);                                   // D = C*B
  const Fp2 e = exp_neg_u(L, d);                          // E = D^-u
  t = cyclotomic_sqr(L, e);                               // F = E^2
  a = exp_neg_u(L, t);                                    // G = F^-u
  t = mul(L, conj(L, a), e);                              // J = (1/G)*E
  const Fp2 k = mul(L, t, conj(L, d));                    // K = J*(1/D)
  d = mul(L, k, b);                                       // L = K*B
  t = mul(L, mul(L, k, e), x);                            // N = K*E*x
  t = mul(L, frobenius(L, d, 1), t);                      // P = L^p * N
  t = mul(L, frobenius(L, k, 2), t);                      // R = K^(p^2) * P
  a = mul(L, conj(L, x), d);                              // T = (1/x)*L
  return mul(L, frobenius(L, a, 3), t);                   // V = T^(p^3) * R
}

// tower-order index of the coefficient of w^k (struct order c0.c0 c0.c1 c0.c2 c1.c0 c1.c1 c1.c2)
RB_FN int tower_index(int k) { return (k & 1) ? 3 + (k >> 1) : (k >> 1); }

// ------------------------------------------------------------------------------------------ Miller steps on lane pairs
// The G2 point T of a pair lives on two lanes (r = 0 / 1): both keep X, Y, Z; the Fq2 products of a step are split
// between them (one code path: operands are selected by r) and the halves exchanged.  Line: l0 + l3 w^3 + l4 w^4,
// returned with l3 * yP and l4 * xP already applied.
struct G2H { Fp2 x, y, z; };
// a line as its pair of lanes keeps it for the broadcast: lane r = 0 holds l3, l4, lane r = 1 holds xi l3, xi l4 in the
// SAME variables, so a reader lane picks the version it needs (wrapped terms take xi) by picking the source lane.
struct Line { Fp2 l0, l3, l4; };
RB_FN void line_finish(const Lane& L, Line* ln, bool present) {
  const bool r = (L.k & 1) != 0;
  ln->l0 = sel2(present, fp2_one(), ln->l0);
  ln->l3 = sel2(present, fp2_zero(), sel2(r, ln->l3, fp2_mul_xi(ln->l3)));
  ln->l4 = sel2(present, fp2_zero(), sel2(r, ln->l4, fp2_mul_xi(ln->l4)));
}

RB_FN Fp2 xchg2(const Lane& L, const Fp2& a) { return shfl_fp2(L, a, L.k ^ 1); }
RB_FN Fp2 half(const Fp2& a) { return fp2_half(a); }

static RB_NOINLINE void pair_dbl_step(Lane L, G2H* t, Fp xp, Fp yp, Line* out) {
  const bool r = (L.k & 1) != 0;
  const Fp2 X = t->x, Y = t->y, Z = t->z;
  const Fp2 u1 = fp2_mul(sel2(r, X, Y), Y);                                 // r0: A' = X Y      r1: B = Y^2
  const Fp2 u2 = fp2_sqr(sel2(r, Z, X));                                    // r0: C = Z^2       r1: J = X^2
  const Fp2 o2 = xchg2(L, u2);
  const Fp2 C = sel2(r, u2, o2), J = sel2(r, o2, u2);
  const Fp2 c3 = fp2_add(fp2_dbl(C), C), yz = fp2_add(Y, Z);
  const Fp2 u3 = fp2_mul(sel2(r, yz, RB_K2(TWIST_B)), sel2(r, yz, c3));     // r0: H' = (Y+Z)^2  r1: E = b' 3C
  const Fp2 o1 = xchg2(L, u1), o3 = xchg2(L, u3);
  const Fp2 A = half(sel2(r, u1, o1)), B = sel2(r, o1, u1);
  const Fp2 Hs = sel2(r, u3, o3), E = sel2(r, o3, u3);
  const Fp2 F = fp2_add(fp2_dbl(E), E);
  const Fp2 G = half(fp2_add(B, F));
  const Fp2 H = fp2_sub(Hs, fp2_add(B, C));
  const Fp2 u4 = fp2_mul(sel2(r, A, G), sel2(r, fp2_sub(B, F), G));         // r0: X' = A (B - F)   r1: G^2
  const Fp2 u5 = fp2_mul(sel2(r, B, E), sel2(r, H, E));                     // r0: Z' = B H         r1: E^2
  const Fp2 o4 = xchg2(L, u4), o5 = xchg2(L, u5);
  const Fp2 G2 = sel2(r, o4, u4), E2 = sel2(r, o5, u5);
  t->x = sel2(r, u4, o4);
  t->y = fp2_sub(G2, fp2_add(fp2_dbl(E2), E2));
  t->z = sel2(r, u5, o5);
  // line: l0 = xi (E - B), l3 = -H (times yP), l4 = 3 J (times xP); r0 scales l3, r1 scales l4
  const Fp2 u6 = fp2_mul_fp(sel2(r, fp2_neg(H), fp2_add(fp2_dbl(J), J)), sel(r, yp, xp));
  const Fp2 o6 = xchg2(L, u6);
  out->l0 = fp2_mul_xi(fp2_sub(E, B));
  out->l3 = sel2(r, u6, o6);
  out->l4 = sel2(r, o6, u6);
}

// T <- T + Q for the affine point Q = (qx, qy)
static RB_NOINLINE void pair_add_step(Lane L, G2H* t, Fp2 qx, Fp2 qy, Fp xp, Fp yp, Line* out) {
  const bool r = (L.k & 1) != 0;
  const Fp2 X = t->x, Y = t->y, Z = t->z;
  const Fp2 u1 = fp2_mul(sel2(r, qx, qy), Z);                               // r0: qx Z   r1: qy Z
  const Fp2 o1 = xchg2(L, u1);
  const Fp2 D = fp2_sub(X, sel2(r, u1, o1)), E = fp2_sub(Y, sel2(r, o1, u1));
  const Fp2 u2 = fp2_sqr(sel2(r, D, E));                                    // r0: F = D^2   r1: G = E^2
  const Fp2 u3 = fp2_mul(sel2(r, E, D), sel2(r, qx, qy));                   // r0: E qx      r1: D qy
  const Fp2 o2 = xchg2(L, u2), o3 = xchg2(L, u3);
  const Fp2 F = sel2(r, u2, o2), G = sel2(r, o2, u2);
  const Fp2 u4 = fp2_mul(sel2(r, D, X), F);                                 // r0: Hh = D F  r1: I = X F
  const Fp2 u5 = fp2_mul(Z, sel2(r, G, G));                                 // both: Z G (needed by both halves of round 6)
  const Fp2 o4 = xchg2(L, u4);
  const Fp2 Hh = sel2(r, u4, o4), I = sel2(r, o4, u4);
  const Fp2 Jv = fp2_sub(fp2_add(Hh, u5), fp2_dbl(I));
  const Fp2 u6 = fp2_mul(sel2(r, D, E), sel2(r, Jv, fp2_sub(I, Jv)));       // r0: X' = D J  r1: E (I - J)
  const Fp2 u7 = fp2_mul(Hh, sel2(r, Y, Z));                                // r0: Hh Y      r1: Z' = Z Hh
  const Fp2 o6 = xchg2(L, u6), o7 = xchg2(L, u7);
  t->x = sel2(r, u6, o6);
  t->y = fp2_sub(sel2(r, o6, u6), sel2(r, u7, o7));
  t->z = sel2(r, o7, u7);
  const Fp2 u8 = fp2_mul_fp(sel2(r, D, fp2_neg(E)), sel(r, yp, xp));         // r0: l3 = D yP   r1: l4 = -E xP
  const Fp2 o8 = xchg2(L, u8);
  out->l0 = fp2_mul_xi(fp2_sub(sel2(r, u3, o3), sel2(r, o3, u3)));          // xi (E qx - D qy)
  out->l3 = sel2(r, u8, o8);
  out->l4 = sel2(r, o8, u8);
}

// a precomputed line of a fixed G2 argument, scaled by the pair's G1 point: r0 scales l3, r1 scales l4
RB_FN void pair_fixed_line(const Lane& L, const FullLine* ln, const Fp& xp, const Fp& yp, Line* out) {
  const bool r = (L.k & 1) != 0;
  const FullFp2* src = r ? &ln->l4 : &ln->l3;
  const Fp2 u = fp2_mul_fp(*src, sel(r, yp, xp));
  const Fp2 o = xchg2(L, u);
  out->l0 = ln->l0;
  out->l3 = sel2(r, u, o);
  out->l4 = sel2(r, o, u);
}

// f *= line of pair j (held by lanes 2j, 2j+1 after line_finish).  Lane k: c_k = f_k l0 + f_{k-3} L3 + f_{k-4} L4 with
// xi on the wrapped terms (k < 3, k < 4): those lanes read the odd lane of the pair.
static RB_NOINLINE Fp2 mul_pair_line(Lane L, Fp2 f, const Line* mine, int j) {
  const bool w3 = L.k < 3, w4 = L.k < 4;
  const Fp2 l0 = shfl_fp2(L, mine->l0, 2 * j), z3 = shfl_fp2(L, mine->l3, 2 * j + (w3 ? 1 : 0)), z4 = shfl_fp2(L, mine->l4, 2 * j + (w4 ? 1 : 0));
  const Fp2 f3 = shfl_fp2(L, f, w3 ? L.k + 3 : L.k - 3), f4 = shfl_fp2(L, f, w4 ? L.k + 2 : L.k - 4);
  return dot3(f, l0, f3, z3, f4, z4);
}

// The same for a line of a table normalised to l0 = 1 (loaded AC17 keys, k_lines_normalize): c_k = f_k + f_{k-3} L3 + f_{k-4} L4,
// two products per lane instead of three.  An absent pair (line_finish: l3 = l4 = 0) leaves f as it is.
static RB_NOINLINE Fp2 mul_pair_line_unit(Lane L, Fp2 f, const Line* mine, int j) {
  const bool w3 = L.k < 3, w4 = L.k < 4;
  const Fp2 z3 = shfl_fp2(L, mine->l3, 2 * j + (w3 ? 1 : 0)), z4 = shfl_fp2(L, mine->l4, 2 * j + (w4 ? 1 : 0));
  const Fp2 f3 = shfl_fp2(L, f, w3 ? L.k + 3 : L.k - 3), f4 = shfl_fp2(L, f, w4 ? L.k + 2 : L.k - 4);
  return fp2_add(f, dot2(f3, z3, f4, z4));
}

// One item = up to three terms; term j pairs (pv[j], q[j]) -- variable G2 argument, walked here -- with
// (pf[j], fixed argument of lines[j]) -- precomputed line table; both kinds may be absent (has_v / has_f false:
// they contribute one; their inputs must still be valid stand-ins).  Lanes 2j, 2j+1 hold term j's points.
//   f = prod_j miller(pv_j, q_j) * miller(pf_j, Q_j)
// All six Miller loops of the item run on ONE accumulator (one f^2 per doubling step).
struct PairState {
  G2H t; Fp2 qx, qy;            // the walking point and the affine Q of this lane's term
  Fp xv, yv, xf, yf;            // G1 points of the variable / fixed pair
  const FullLine* lines;        // line table of the fixed argument
  bool has_v, has_f;
  bool unit_fixed;              // the fixed arguments' line tables are normalised to l0 = 1
};

static RB_NOINLINE Fp2 miller_terms(Lane L, PairState* s, int n_terms) {
  Fp2 f = one(L);
  const Fp2 nqy = fp2_neg(s->qy);
  Line lv, lf;
  int li = 0;

#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
  for (int i = ATE_NAF_LEN - 2; i >= -2; --i) {                 // i == -1, -2: the two Frobenius chords
    if (i >= 0 && i != ATE_NAF_LEN - 2) f = sqr(L, f);
    const int d = (i >= 0) ? ATE_NAF[i] : 0;
    const int nl = (i >= 0 && d != 0) ? 2 : 1;
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
    for (int st = 0; st < nl; ++st, ++li) {
      if (i >= 0 && st == 0) pair_dbl_step(L, &s->t, s->xv, s->yv, &lv);
      else {
        Fp2 ax, ay;
        if (i >= 0) { ax = s->qx; ay = d > 0 ? s->qy : nqy; }
        else if (i == -1) { ax = fp2_mul(fp2_conj(s->qx), RB_K2(FROB1[2])); ay = fp2_mul(fp2_conj(s->qy), RB_K2(FROB1[3])); }
        else { ax = fp2_mul(s->qx, RB_K2(FROB2[2])); ay = fp2_neg(fp2_mul(s->qy, RB_K2(FROB2[3]))); }
        pair_add_step(L, &s->t, ax, ay, s->xv, s->yv, &lv);
      }
      pair_fixed_line(L, s->lines + li, s->xf, s->yf, &lf);
      line_finish(L, &lv, s->has_v);
      line_finish(L, &lf, s->has_f);
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
      for (int j = 0; j < n_terms; ++j) {
        f = mul_pair_line(L, f, &lv, j);
        f = s->unit_fixed ? mul_pair_line_unit(L, f, &lf, j) : mul_pair_line(L, f, &lf, j);
      }
    }
  }
  return f;
}

}  // namespace w6
}  // namespace rb
