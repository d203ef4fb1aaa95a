// 254-bit prime-field arithmetic for BN254 on sm_100a: 8 x 32-bit limbs held in registers,
// Montgomery form with R = 2^256, values kept fully reduced in [0, N).
//
// Replaces the Fq / Fr types of the external crate rabe-bn 0.4.23 that rabe's schemes use
// (/root/reference/Cargo.toml:33; call sites: SURVEY.md section 2, "L1 operator call-site
// inventory").  The multiplication is an operand-scanning Montgomery product that keeps two
// accumulators -- one for products of even-indexed limbs, one (shifted by a limb) for odd-indexed
// limbs -- so that every partial product is a mad.lo.cc/madc.hi.cc pair on one unbroken carry
// chain (ptxas fuses each pair into one IMAD.WIDE.U32[.X]); the accumulators swap roles at every
// row instead of being shifted.
//
// The same header compiles for the host when RB_HOST_SIM is defined (tests/hostsim only; the
// product library never builds that path).
#pragma once
#include <stdint.h>

#if defined(RB_HOST_SIM)
#define RB_FN inline
#define RB_NOINLINE __attribute__((noinline))
#define RB_UNROLL
#else
#define RB_FN __device__ __forceinline__
#define RB_NOINLINE __device__ __noinline__
#define RB_UNROLL _Pragma("unroll")
#endif

// Miller doubling / addition steps: inlined into the loops by default; RB_STEP_NOINLINE makes them
// real functions (smaller register live ranges and code at the price of a call), for occupancy sweeps.
#if defined(RB_STEP_NOINLINE)
#define RB_STEP_FN static RB_NOINLINE
#else
#define RB_STEP_FN RB_FN
#endif
// Fq6 additions / multiplication by v / by xi: inline by default, real functions with RB_COMPACT
// (instruction-cache footprint of the pairing kernels when several kernels share an SM).
#if defined(RB_COMPACT) && !defined(RB_HOST_SIM)
#define RB_SMALL_FN static RB_NOINLINE
#else
#define RB_SMALL_FN RB_FN
#endif

namespace rb {

struct ModP {   // base field Fq
  static RB_FN constexpr uint32_t N(int i) {
    constexpr uint32_t t[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return t[i];
  }
  static RB_FN constexpr uint32_t R1(int i) {
    constexpr uint32_t t[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return t[i];
  }
  static RB_FN constexpr uint32_t R2(int i) {
    constexpr uint32_t t[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return t[i];
  }
  static constexpr uint32_t INV = 0xe4866389u;
};

struct ModR {   // scalar field Fr
  static RB_FN constexpr uint32_t N(int i) {
    constexpr uint32_t t[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return t[i];
  }
  static RB_FN constexpr uint32_t R1(int i) {
    constexpr uint32_t t[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return t[i];
  }
  static RB_FN constexpr uint32_t R2(int i) {
    constexpr uint32_t t[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return t[i];
  }
  static constexpr uint32_t INV = 0xefffffffu;
};

template <class M>
struct Fe {
  uint32_t v[8];
};
#if defined(RB_HOST_SIM)
static unsigned long long g_host_mul_count = 0;
#endif
typedef Fe<ModP> Fp;
typedef Fe<ModR> Fr;

// ------------------------------------------------------------------------------------------
template <class M> RB_FN Fe<M> fe_zero() { Fe<M> r; RB_UNROLL for (int i = 0; i < 8; ++i) r.v[i] = 0; return r; }
template <class M> RB_FN Fe<M> fe_one() { Fe<M> r; RB_UNROLL for (int i = 0; i < 8; ++i) r.v[i] = M::R1(i); return r; }
template <class M> RB_FN bool fe_is_zero(const Fe<M>& a) {
  uint32_t o = 0; RB_UNROLL for (int i = 0; i < 8; ++i) o |= a.v[i]; return o == 0;
}
template <class M> RB_FN bool fe_eq(const Fe<M>& a, const Fe<M>& b) {
  uint32_t o = 0; RB_UNROLL for (int i = 0; i < 8; ++i) o |= a.v[i] ^ b.v[i]; return o == 0;
}

#if defined(__CUDA_ARCH__)
// --- device carry-chain primitives (each chain lives in ONE asm block: CC.CF is implicit state)

// r = a + b (8 limbs), returns nothing; the sum of two values < N < 2^254 cannot carry out.
RB_FN void add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  asm("add.cc.u32 %0,%8,%16; addc.cc.u32 %1,%9,%17; addc.cc.u32 %2,%10,%18; addc.cc.u32 %3,%11,%19;"
      "addc.cc.u32 %4,%12,%20; addc.cc.u32 %5,%13,%21; addc.cc.u32 %6,%14,%22; addc.u32 %7,%15,%23;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
}
// r = a - b (8 limbs); returns the borrow as an all-ones / all-zero mask.
RB_FN uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t m;
  asm("sub.cc.u32 %0,%9,%17; subc.cc.u32 %1,%10,%18; subc.cc.u32 %2,%11,%19; subc.cc.u32 %3,%12,%20;"
      "subc.cc.u32 %4,%13,%21; subc.cc.u32 %5,%14,%22; subc.cc.u32 %6,%15,%23; subc.cc.u32 %7,%16,%24;"
      "subc.u32 %8,0,0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(m)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  return m;
}
// F[0..7] += (x0, x2, x4, x6) * y at limb offsets 0,2,4,6 ; the carry out is added to top.
RB_FN void mad_even(uint32_t* F, uint32_t& top, uint32_t x0, uint32_t x2, uint32_t x4, uint32_t x6, uint32_t y) {
  asm("mad.lo.cc.u32 %0,%9,%13,%0; madc.hi.cc.u32 %1,%9,%13,%1; madc.lo.cc.u32 %2,%10,%13,%2; madc.hi.cc.u32 %3,%10,%13,%3;"
      "madc.lo.cc.u32 %4,%11,%13,%4; madc.hi.cc.u32 %5,%11,%13,%5; madc.lo.cc.u32 %6,%12,%13,%6; madc.hi.cc.u32 %7,%12,%13,%7;"
      "addc.u32 %8,%8,0;"
      : "+r"(F[0]), "+r"(F[1]), "+r"(F[2]), "+r"(F[3]), "+r"(F[4]), "+r"(F[5]), "+r"(F[6]), "+r"(F[7]), "+r"(top)
      : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
}
// S[0..7] += (x1, x3, x5, x7) * y at limb offsets 0,2,4,6 of S (S itself sits one limb up).
RB_FN void mad_odd(uint32_t* S, uint32_t x1, uint32_t x3, uint32_t x5, uint32_t x7, uint32_t y) {
  asm("mad.lo.cc.u32 %0,%8,%12,%0; madc.hi.cc.u32 %1,%8,%12,%1; madc.lo.cc.u32 %2,%9,%12,%2; madc.hi.cc.u32 %3,%9,%12,%3;"
      "madc.lo.cc.u32 %4,%10,%12,%4; madc.hi.cc.u32 %5,%10,%12,%5; madc.lo.cc.u32 %6,%11,%12,%6; madc.hi.u32 %7,%11,%12,%7;"
      : "+r"(S[0]), "+r"(S[1]), "+r"(S[2]), "+r"(S[3]), "+r"(S[4]), "+r"(S[5]), "+r"(S[6]), "+r"(S[7])
      : "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(y));
}
// Role swap at a row boundary: the old offset-0 accumulator O (whose limb 0 is now zero) becomes
// the new offset-1 accumulator, two limbs down; its orphaned limb O[1] is folded into F[0] and the
// carry of that fold rides into the first limb of the new chain.  Then adds (x1,x3,x5,x7)*y.
RB_FN void mad_odd_swap(uint32_t& F0, uint32_t* Sn, const uint32_t* O, uint32_t x1, uint32_t x3, uint32_t x5, uint32_t x7, uint32_t y) {
  asm("add.cc.u32 %0,%0,%9;"
      "madc.lo.cc.u32 %1,%16,%20,%10; madc.hi.cc.u32 %2,%16,%20,%11; madc.lo.cc.u32 %3,%17,%20,%12; madc.hi.cc.u32 %4,%17,%20,%13;"
      "madc.lo.cc.u32 %5,%18,%20,%14; madc.hi.cc.u32 %6,%18,%20,%15; madc.lo.cc.u32 %7,%19,%20,0; madc.hi.u32 %8,%19,%20,0;"
      : "+r"(F0), "=r"(Sn[0]), "=r"(Sn[1]), "=r"(Sn[2]), "=r"(Sn[3]), "=r"(Sn[4]), "=r"(Sn[5]), "=r"(Sn[6]), "=r"(Sn[7])
      : "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]),
        "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(y));
}
#endif

// r = a if mask == 0 else b
template <class M> RB_FN Fe<M> fe_select(uint32_t mask, const Fe<M>& a, const Fe<M>& b) {
  Fe<M> r; RB_UNROLL for (int i = 0; i < 8; ++i) r.v[i] = mask ? b.v[i] : a.v[i]; return r;
}

// conditional final subtraction: t in [0, 2N) -> [0, N)
template <class M> RB_FN void fe_reduce_once(uint32_t* t) {
#if defined(__CUDA_ARCH__)
  uint32_t n[8], d[8];
  RB_UNROLL for (int i = 0; i < 8; ++i) n[i] = M::N(i);
  uint32_t borrow = sub8(d, t, n);
  RB_UNROLL for (int i = 0; i < 8; ++i) t[i] = borrow ? t[i] : d[i];
#else
  uint32_t d[8]; uint64_t br = 0;
  for (int i = 0; i < 8; ++i) { uint64_t x = (uint64_t)t[i] - M::N(i) - br; d[i] = (uint32_t)x; br = (x >> 32) & 1; }
  if (!br) for (int i = 0; i < 8; ++i) t[i] = d[i];
#endif
}

template <class M> RB_FN Fe<M> fe_add(const Fe<M>& a, const Fe<M>& b) {
  Fe<M> r;
#if defined(__CUDA_ARCH__)
  add8(r.v, a.v, b.v);
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; ++i) { uint64_t x = (uint64_t)a.v[i] + b.v[i] + c; r.v[i] = (uint32_t)x; c = x >> 32; }
#endif
  fe_reduce_once<M>(r.v);
  return r;
}

template <class M> RB_FN Fe<M> fe_sub(const Fe<M>& a, const Fe<M>& b) {
  Fe<M> r;
#if defined(__CUDA_ARCH__)
  uint32_t d[8], e[8], n[8];
  uint32_t borrow = sub8(d, a.v, b.v);
  RB_UNROLL for (int i = 0; i < 8; ++i) n[i] = M::N(i);
  add8(e, d, n);
  RB_UNROLL for (int i = 0; i < 8; ++i) r.v[i] = borrow ? e[i] : d[i];
#else
  uint64_t br = 0;
  for (int i = 0; i < 8; ++i) { uint64_t x = (uint64_t)a.v[i] - b.v[i] - br; r.v[i] = (uint32_t)x; br = (x >> 32) & 1; }
  if (br) { uint64_t c = 0; for (int i = 0; i < 8; ++i) { uint64_t x = (uint64_t)r.v[i] + M::N(i) + c; r.v[i] = (uint32_t)x; c = x >> 32; } }
#endif
  return r;
}

template <class M> RB_FN Fe<M> fe_neg(const Fe<M>& a) {
  Fe<M> n; RB_UNROLL for (int i = 0; i < 8; ++i) n.v[i] = M::N(i);
  Fe<M> z = fe_zero<M>();
  if (fe_is_zero(a)) return z;
  Fe<M> r;
#if defined(__CUDA_ARCH__)
  sub8(r.v, n.v, a.v);
#else
  uint64_t br = 0;
  for (int i = 0; i < 8; ++i) { uint64_t x = (uint64_t)n.v[i] - a.v[i] - br; r.v[i] = (uint32_t)x; br = (x >> 32) & 1; }
#endif
  return r;
}

template <class M> RB_FN Fe<M> fe_dbl(const Fe<M>& a) { return fe_add(a, a); }
// a / 2 mod N: (a + (a odd ? N : 0)) >> 1 -- the Montgomery form halves like the value it stands for
template <class M> RB_FN Fe<M> fe_half(const Fe<M>& a) {
  uint32_t t[9]; uint64_t c = 0;
  const uint32_t odd = 0u - (a.v[0] & 1u);
  RB_UNROLL for (int i = 0; i < 8; ++i) { uint64_t x = (uint64_t)a.v[i] + (M::N(i) & odd) + c; t[i] = (uint32_t)x; c = x >> 32; }
  t[8] = (uint32_t)c;
  Fe<M> r;
  RB_UNROLL for (int i = 0; i < 8; ++i) r.v[i] = (t[i] >> 1) | (t[i + 1] << 31);
  return r;
}

// Montgomery product a*b/R mod N
template <class M> RB_FN Fe<M> fe_mul(const Fe<M>& a, const Fe<M>& b) {
  Fe<M> r;
#if defined(__CUDA_ARCH__)
  uint32_t A[8], B[8];      // the two accumulators; which one is "offset 0" alternates per row
  const uint32_t* x = a.v;
  {
    const uint32_t y = b.v[0];
    RB_UNROLL for (int k = 0; k < 4; ++k) {
      uint64_t e = (uint64_t)x[2 * k] * y;       A[2 * k] = (uint32_t)e; A[2 * k + 1] = (uint32_t)(e >> 32);
      uint64_t o = (uint64_t)x[2 * k + 1] * y;   B[2 * k] = (uint32_t)o; B[2 * k + 1] = (uint32_t)(o >> 32);
    }
    const uint32_t m = A[0] * M::INV;
    mad_odd(B, M::N(1), M::N(3), M::N(5), M::N(7), m);
    mad_even(A, B[7], M::N(0), M::N(2), M::N(4), M::N(6), m);
  }
  RB_UNROLL for (int i = 1; i < 8; ++i) {
    const uint32_t y = b.v[i];
    if (i & 1) {           // offset-0 accumulator is B, A is demoted
      uint32_t S[8];
      mad_odd_swap(B[0], S, A, x[1], x[3], x[5], x[7], y);
      mad_even(B, S[7], x[0], x[2], x[4], x[6], y);
      const uint32_t m = B[0] * M::INV;
      mad_odd(S, M::N(1), M::N(3), M::N(5), M::N(7), m);
      mad_even(B, S[7], M::N(0), M::N(2), M::N(4), M::N(6), m);
      RB_UNROLL for (int k = 0; k < 8; ++k) A[k] = S[k];
    } else {               // offset-0 accumulator is A, B is demoted
      uint32_t S[8];
      mad_odd_swap(A[0], S, B, x[1], x[3], x[5], x[7], y);
      mad_even(A, S[7], x[0], x[2], x[4], x[6], y);
      const uint32_t m = A[0] * M::INV;
      mad_odd(S, M::N(1), M::N(3), M::N(5), M::N(7), m);
      mad_even(A, S[7], M::N(0), M::N(2), M::N(4), M::N(6), m);
      RB_UNROLL for (int k = 0; k < 8; ++k) B[k] = S[k];
    }
  }
  // after row 7 (odd): offset-0 accumulator B has B[0] == 0, offset-1 accumulator is A.
  // result = B[1..7] + A
  asm("add.cc.u32 %0,%0,%8; addc.cc.u32 %1,%1,%9; addc.cc.u32 %2,%2,%10; addc.cc.u32 %3,%3,%11;"
      "addc.cc.u32 %4,%4,%12; addc.cc.u32 %5,%5,%13; addc.cc.u32 %6,%6,%14; addc.u32 %7,%7,0;"
      : "+r"(A[0]), "+r"(A[1]), "+r"(A[2]), "+r"(A[3]), "+r"(A[4]), "+r"(A[5]), "+r"(A[6]), "+r"(A[7])
      : "r"(B[1]), "r"(B[2]), "r"(B[3]), "r"(B[4]), "r"(B[5]), "r"(B[6]), "r"(B[7]));
  RB_UNROLL for (int k = 0; k < 8; ++k) r.v[k] = A[k];
#else
#if defined(RB_HOST_SIM)
  ++g_host_mul_count;      // test-only instrumentation: Fp/Fr products executed (tools/gen_op_counts.py)
#endif
  uint32_t t[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 8; ++i) {
    uint64_t c = 0;
    for (int j = 0; j < 8; ++j) { uint64_t s = (uint64_t)a.v[j] * b.v[i] + t[j] + c; t[j] = (uint32_t)s; c = s >> 32; }
    uint64_t s = (uint64_t)t[8] + c; t[8] = (uint32_t)s; t[9] = (uint32_t)(s >> 32);
    uint32_t m = t[0] * M::INV;
    s = (uint64_t)m * M::N(0) + t[0]; c = s >> 32;
    for (int j = 1; j < 8; ++j) { s = (uint64_t)m * M::N(j) + t[j] + c; t[j - 1] = (uint32_t)s; c = s >> 32; }
    s = (uint64_t)t[8] + c; t[7] = (uint32_t)s; t[8] = t[9] + (uint32_t)(s >> 32);
  }
  for (int k = 0; k < 8; ++k) r.v[k] = t[k];
#endif
  fe_reduce_once<M>(r.v);
  return r;
}

#if defined(__CUDA_ARCH__)
// ---- dedicated squaring (device): 28 cross products + 8 squares + one Montgomery reduction of the 512-bit result = 108
// IMAD-pipe instructions instead of the 136 of fe_mul(a, a).
// F[0..2K-1] += (x_0 .. x_{K-1}) * y at limb offsets 0, 2, ..; the carry out is added to top (K = 3, 2, 1; K = 4 is mad_even)
RB_FN void mad_k3(uint32_t* F, uint32_t& top, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t y) {
  asm("mad.lo.cc.u32 %0,%7,%10,%0; madc.hi.cc.u32 %1,%7,%10,%1; madc.lo.cc.u32 %2,%8,%10,%2; madc.hi.cc.u32 %3,%8,%10,%3;"
      "madc.lo.cc.u32 %4,%9,%10,%4; madc.hi.cc.u32 %5,%9,%10,%5; addc.u32 %6,%6,0;"
      : "+r"(F[0]), "+r"(F[1]), "+r"(F[2]), "+r"(F[3]), "+r"(F[4]), "+r"(F[5]), "+r"(top)
      : "r"(x0), "r"(x1), "r"(x2), "r"(y));
}
RB_FN void mad_k2(uint32_t* F, uint32_t& top, uint32_t x0, uint32_t x1, uint32_t y) {
  asm("mad.lo.cc.u32 %0,%5,%7,%0; madc.hi.cc.u32 %1,%5,%7,%1; madc.lo.cc.u32 %2,%6,%7,%2; madc.hi.cc.u32 %3,%6,%7,%3; addc.u32 %4,%4,0;"
      : "+r"(F[0]), "+r"(F[1]), "+r"(F[2]), "+r"(F[3]), "+r"(top)
      : "r"(x0), "r"(x1), "r"(y));
}
RB_FN void mad_k1(uint32_t* F, uint32_t& top, uint32_t x0, uint32_t y) {
  asm("mad.lo.cc.u32 %0,%3,%4,%0; madc.hi.cc.u32 %1,%3,%4,%1; addc.u32 %2,%2,0;"
      : "+r"(F[0]), "+r"(F[1]), "+r"(top)
      : "r"(x0), "r"(y));
}
// r[0..7] = a + b + cin (8 limbs), returns the carry out
RB_FN uint32_t add8c(uint32_t* r, const uint32_t* a, const uint32_t* b, uint32_t cin) {
  uint32_t cout;
  asm("{ .reg .u32 t; add.cc.u32 t,%25,0xffffffff; addc.cc.u32 %0,%9,%17; addc.cc.u32 %1,%10,%18; addc.cc.u32 %2,%11,%19; addc.cc.u32 %3,%12,%20;"
      "addc.cc.u32 %4,%13,%21; addc.cc.u32 %5,%14,%22; addc.cc.u32 %6,%15,%23; addc.cc.u32 %7,%16,%24; addc.u32 %8,0,0; }"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(cout)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(cin));
  return cout;
}
// Role swap of the two Montgomery accumulators at a row boundary of a pure reduction (mad_odd_swap without products)
RB_FN void redc_swap8(uint32_t& F0, uint32_t* Sn, const uint32_t* O) {
  asm("add.cc.u32 %0,%0,%9; addc.cc.u32 %1,%10,0; addc.cc.u32 %2,%11,0; addc.cc.u32 %3,%12,0; addc.cc.u32 %4,%13,0;"
      "addc.cc.u32 %5,%14,0; addc.cc.u32 %6,%15,0; addc.u32 %7,0,0; mov.u32 %8,0;"
      : "+r"(F0), "=r"(Sn[0]), "=r"(Sn[1]), "=r"(Sn[2]), "=r"(Sn[3]), "=r"(Sn[4]), "=r"(Sn[5]), "=r"(Sn[6]), "=r"(Sn[7])
      : "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
}
// t (512 bits, below 4 N^2) / R mod N, fully reduced:  t = hi 2^256 + lo,  t / R = hi + lo / R (mod N) with lo / R in [0, N]
// from eight reduction rows, and hi + lo / R < 0.76 N + N < 2 N: one conditional subtraction.
template <class M> RB_FN Fe<M> fe_redc_wide(const uint32_t* t) {
  uint32_t A[8], B[8];
  RB_UNROLL for (int k = 0; k < 8; ++k) { A[k] = t[k]; B[k] = 0; }
  {
    const uint32_t m = A[0] * M::INV;
    mad_odd(B, M::N(1), M::N(3), M::N(5), M::N(7), m);
    mad_even(A, B[7], M::N(0), M::N(2), M::N(4), M::N(6), m);
  }
  RB_UNROLL for (int i = 1; i < 8; ++i) {
    uint32_t S[8];
    if (i & 1) {
      redc_swap8(B[0], S, A);
      const uint32_t m = B[0] * M::INV;
      mad_odd(S, M::N(1), M::N(3), M::N(5), M::N(7), m);
      mad_even(B, S[7], M::N(0), M::N(2), M::N(4), M::N(6), m);
      RB_UNROLL for (int k = 0; k < 8; ++k) A[k] = S[k];
    } else {
      redc_swap8(A[0], S, B);
      const uint32_t m = A[0] * M::INV;
      mad_odd(S, M::N(1), M::N(3), M::N(5), M::N(7), m);
      mad_even(A, S[7], M::N(0), M::N(2), M::N(4), M::N(6), m);
      RB_UNROLL for (int k = 0; k < 8; ++k) B[k] = S[k];
    }
  }
  // lo / R = B[1..7] + A  (after row 7 the offset-0 accumulator is B with B[0] == 0), in [0, N]
  uint32_t lo[8], z[8] = {B[1], B[2], B[3], B[4], B[5], B[6], B[7], 0};
  add8c(lo, A, z, 0u);
  Fe<M> r;
  add8c(r.v, lo, t + 8, 0u);                        // < 2 N < 2^255: no carry out
  fe_reduce_once<M>(r.v);
  return r;
}
// a^2 / R mod N for a < 2 N (Montgomery form in, Montgomery form out, fully reduced).  Measured on a B200 (round 2, VERDICT r1
// item 14): as fe_sqr everywhere the pipelined AC17 step goes 560 -> 547 k round trips/s (k_ac17_enc_rows 1.92 -> 2.08 ms: the
// longer ALU tail of the 512-bit assembly and the separate reduction cost more than the 28 products saved); inside the
// inversion's 254 squarings only, 558 k (k_g1_gather_sum 0.68 -> 0.67 ms) -- no gain either way.  fe_sqr therefore stays the
// product; this routine is kept behind -DRB_FE_SQR_WIDE=1 and under test (tests/test_gpu_wide.py).
template <class M> RB_FN Fe<M> fe_sqr_wide(const Fe<M>& x) {
  const uint32_t* a = x.v;
  // cross products a_i a_j, i < j, in two accumulators: E takes those that start at an even limb, O (one limb up) the others,
  // so every 32x32 product is a mad.lo.cc / madc.hi.cc pair on an aligned register pair.  Rows ascend, which makes every
  // carry target either untouched or a small carry count when the carry arrives (no ripple needed).
  uint32_t E[16], O[16];
  RB_UNROLL for (int k = 0; k < 16; ++k) { E[k] = 0; O[k] = 0; }
  mad_k3(E + 2, E[8], a[2], a[4], a[6], a[0]);    mad_even(O + 0, O[8], a[1], a[3], a[5], a[7], a[0]);
  mad_k3(E + 4, E[10], a[3], a[5], a[7], a[1]);   mad_k3(O + 2, O[8], a[2], a[4], a[6], a[1]);
  mad_k2(E + 6, E[10], a[4], a[6], a[2]);         mad_k3(O + 4, O[10], a[3], a[5], a[7], a[2]);
  mad_k2(E + 8, E[12], a[5], a[7], a[3]);         mad_k2(O + 6, O[10], a[4], a[6], a[3]);
  mad_k1(E + 10, E[12], a[6], a[4]);              mad_k2(O + 8, O[12], a[5], a[7], a[4]);
  mad_k1(E + 12, E[14], a[7], a[5]);              mad_k1(O + 10, O[12], a[6], a[5]);
                                                  mad_k1(O + 12, O[14], a[7], a[6]);
  // C = E + (O << 32)   (C[0] = E[0] = 0; below 2^511)
  uint32_t C[16];
  C[0] = E[0];
  {
    const uint32_t c1 = add8c(C + 1, E + 1, O, 0u);
    const uint32_t e2[8] = {E[9], E[10], E[11], E[12], E[13], E[14], E[15], 0};
    uint32_t hi[8];
    add8c(hi, e2, O + 8, c1);
    RB_UNROLL for (int k = 0; k < 7; ++k) C[9 + k] = hi[k];
  }
  // T = 2 C + sum_i a_i^2 2^(64 i)
  uint32_t S[16], D[16], T[16];
  S[0] = C[0] << 1;
  RB_UNROLL for (int k = 1; k < 16; ++k) S[k] = __funnelshift_l(C[k - 1], C[k], 1);
  RB_UNROLL for (int i = 0; i < 8; ++i) { const uint64_t d = (uint64_t)a[i] * a[i]; D[2 * i] = (uint32_t)d; D[2 * i + 1] = (uint32_t)(d >> 32); }
  const uint32_t c2 = add8c(T, S, D, 0u);
  add8c(T + 8, S + 8, D + 8, c2);
  return fe_redc_wide<M>(T);
}
#else
template <class M> RB_FN Fe<M> fe_sqr_wide(const Fe<M>& a) { return fe_mul(a, a); }
#endif
template <class M> RB_FN Fe<M> fe_sqr(const Fe<M>& a) { return fe_mul(a, a); }

// (a*b + c*d) / R mod N, fully reduced: two operand-scanning products share one Montgomery
// reduction (a row adds x*y_i AND z*w_i before the row's m*N).  Needs a, c <= N and b, d < N:
// the running value stays below 3N < 2^256 and the result below 2N^2/R + N < 2N, so the single
// conditional subtraction of fe_mul is enough.  192 + 8 IMAD-pipe instructions instead of 2 x 136;
// used by the lane-paired Fq2 product (coop.cuh), where each lane owns one component.
template <class M> RB_FN Fe<M> fe_mul2add(const Fe<M>& a, const Fe<M>& b, const Fe<M>& c, const Fe<M>& d) {
  Fe<M> r;
#if defined(__CUDA_ARCH__)
  uint32_t A[8], B[8];
  const uint32_t* x = a.v;
  const uint32_t* z = c.v;
  {
    const uint32_t y = b.v[0], w = d.v[0];
    RB_UNROLL for (int k = 0; k < 4; ++k) {
      uint64_t e = (uint64_t)x[2 * k] * y;       A[2 * k] = (uint32_t)e; A[2 * k + 1] = (uint32_t)(e >> 32);
      uint64_t o = (uint64_t)x[2 * k + 1] * y;   B[2 * k] = (uint32_t)o; B[2 * k + 1] = (uint32_t)(o >> 32);
    }
    mad_odd(B, z[1], z[3], z[5], z[7], w);
    mad_even(A, B[7], z[0], z[2], z[4], z[6], w);
    const uint32_t m = A[0] * M::INV;
    mad_odd(B, M::N(1), M::N(3), M::N(5), M::N(7), m);
    mad_even(A, B[7], M::N(0), M::N(2), M::N(4), M::N(6), m);
  }
  RB_UNROLL for (int i = 1; i < 8; ++i) {
    const uint32_t y = b.v[i], w = d.v[i];
    if (i & 1) {
      uint32_t S[8];
      mad_odd_swap(B[0], S, A, x[1], x[3], x[5], x[7], y);
      mad_even(B, S[7], x[0], x[2], x[4], x[6], y);
      mad_odd(S, z[1], z[3], z[5], z[7], w);
      mad_even(B, S[7], z[0], z[2], z[4], z[6], w);
      const uint32_t m = B[0] * M::INV;
      mad_odd(S, M::N(1), M::N(3), M::N(5), M::N(7), m);
      mad_even(B, S[7], M::N(0), M::N(2), M::N(4), M::N(6), m);
      RB_UNROLL for (int k = 0; k < 8; ++k) A[k] = S[k];
    } else {
      uint32_t S[8];
      mad_odd_swap(A[0], S, B, x[1], x[3], x[5], x[7], y);
      mad_even(A, S[7], x[0], x[2], x[4], x[6], y);
      mad_odd(S, z[1], z[3], z[5], z[7], w);
      mad_even(A, S[7], z[0], z[2], z[4], z[6], w);
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
#if defined(RB_HOST_SIM)
  g_host_mul_count += 2;
#endif
  uint32_t t[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 8; ++i) {
    uint64_t cy = 0;
    for (int j = 0; j < 8; ++j) { uint64_t s = (uint64_t)a.v[j] * b.v[i] + t[j] + cy; t[j] = (uint32_t)s; cy = s >> 32; }
    uint64_t s = (uint64_t)t[8] + cy; t[8] = (uint32_t)s; t[9] = (uint32_t)(s >> 32);
    cy = 0;
    for (int j = 0; j < 8; ++j) { uint64_t s2 = (uint64_t)c.v[j] * d.v[i] + t[j] + cy; t[j] = (uint32_t)s2; cy = s2 >> 32; }
    s = (uint64_t)t[8] + cy; t[8] = (uint32_t)s; t[9] += (uint32_t)(s >> 32);
    uint32_t m = t[0] * M::INV;
    s = (uint64_t)m * M::N(0) + t[0]; cy = s >> 32;
    for (int j = 1; j < 8; ++j) { s = (uint64_t)m * M::N(j) + t[j] + cy; t[j - 1] = (uint32_t)s; cy = s >> 32; }
    s = (uint64_t)t[8] + cy; t[7] = (uint32_t)s; t[8] = t[9] + (uint32_t)(s >> 32); t[9] = 0;
  }
  for (int k = 0; k < 8; ++k) r.v[k] = t[k];
#endif
  fe_reduce_once<M>(r.v);
  return r;
}

// into / out of Montgomery form
template <class M> RB_FN Fe<M> fe_to_mont(const Fe<M>& a) {
  Fe<M> r2; RB_UNROLL for (int i = 0; i < 8; ++i) r2.v[i] = M::R2(i);
  return fe_mul(a, r2);
}
template <class M> RB_FN Fe<M> fe_from_mont(const Fe<M>& a) {
  Fe<M> one = fe_zero<M>(); one.v[0] = 1;
  return fe_mul(a, one);
}

// a^(N-2) by plain square-and-multiply (used once per warp-level batch, see batch_inverse)
template <class M> RB_NOINLINE Fe<M> fe_inv(const Fe<M>& a) {
  Fe<M> acc = fe_one<M>();
#if !defined(RB_HOST_SIM)
#pragma unroll 1
#endif
  for (int i = 253; i >= 0; --i) {
#if defined(RB_FE_SQR_WIDE)
    acc = fe_sqr_wide(acc);
#else
    acc = fe_sqr(acc);
#endif
    uint32_t w = M::N(0);
    // bit i of N-2; only limb 0 differs from N
    uint32_t limb;
    switch (i >> 5) {
      case 0: limb = M::N(0) - 2u; break;
      case 1: limb = M::N(1); break;
      case 2: limb = M::N(2); break;
      case 3: limb = M::N(3); break;
      case 4: limb = M::N(4); break;
      case 5: limb = M::N(5); break;
      case 6: limb = M::N(6); break;
      default: limb = M::N(7); break;
    }
    (void)w;
    if ((limb >> (i & 31)) & 1u) acc = fe_mul(acc, a);
  }
  return acc;
}

// 32 canonical big-endian bytes <-> limbs (not Montgomery).  `p` need only be byte aligned on the
// host; on the device it must be 4-byte aligned.
template <class M> RB_FN Fe<M> fe_load_be(const uint8_t* p) {
  Fe<M> r;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
  RB_UNROLL for (int i = 0; i < 8; ++i) {
    uint32_t x = w[7 - i];
#if defined(__CUDA_ARCH__)
    r.v[i] = __byte_perm(x, 0, 0x0123);
#else
    r.v[i] = __builtin_bswap32(x);
#endif
  }
  return r;
}
template <class M> RB_FN void fe_store_be(uint8_t* p, const Fe<M>& a) {
  uint32_t* w = reinterpret_cast<uint32_t*>(p);
  RB_UNROLL for (int i = 0; i < 8; ++i) {
#if defined(__CUDA_ARCH__)
    w[7 - i] = __byte_perm(a.v[i], 0, 0x0123);
#else
    w[7 - i] = __builtin_bswap32(a.v[i]);
#endif
  }
}
// value >= N ?
template <class M> RB_FN bool fe_geq_modulus(const Fe<M>& a) {
  uint64_t br = 0;
  RB_UNROLL for (int i = 0; i < 8; ++i) { uint64_t x = (uint64_t)a.v[i] - M::N(i) - br; br = (x >> 32) & 1; }
  return br == 0;
}

// short names
RB_FN Fp operator+(const Fp& a, const Fp& b) { return fe_add(a, b); }
RB_FN Fp operator-(const Fp& a, const Fp& b) { return fe_sub(a, b); }
RB_FN Fp operator*(const Fp& a, const Fp& b) { return fe_mul(a, b); }
RB_FN Fr operator+(const Fr& a, const Fr& b) { return fe_add(a, b); }
RB_FN Fr operator-(const Fr& a, const Fr& b) { return fe_sub(a, b); }
RB_FN Fr operator*(const Fr& a, const Fr& b) { return fe_mul(a, b); }

}  // namespace rb
