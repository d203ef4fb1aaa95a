// CUDA kernels of the rabe_b200 engine (sm_100a).  One thread owns one (or a short run of)
// independent group operation(s); all field arithmetic runs in registers on the integer pipes,
// HBM is touched only for the canonical inputs/outputs and L2 for the fixed-base tables.
// Kernel <-> reference map is in DESIGN.md section 3.
#pragma once
#include <cuda_runtime.h>
#include "pairing.cuh"

#ifndef RB_PAIR_MINB
#define RB_PAIR_MINB 1       // min resident blocks per SM for the pairing kernels (caps registers)
#endif
#ifndef RB_G1_MINB
#define RB_G1_MINB 1         // min resident 128-thread blocks per SM for the G1 fixed-base kernels (caps registers)
#endif
#ifndef RB_FE_BLOCK
#define RB_FE_BLOCK 128      // threads per block, final exponentiation
#endif
#ifndef RB_ML_BLOCK
#define RB_ML_BLOCK 128      // threads per block, Miller loops
#endif

namespace rb {

enum { ERR_NOT_MEMBER = 1, ERR_POLICY = 2 };   // device error flags (engine.cu map_flags: RB_ENOTMEMBER, RB_EPOLICY)

__device__ __forceinline__ void flag_error(int* err, int code) { atomicOr(err, code); }

// canonical big-endian Fr -> limbs (non-Montgomery); flags values >= r
__device__ __forceinline__ Fr load_scalar(const uint8_t* p, int* err) {
  Fr k = fe_load_be<ModR>(p);
  if (fe_geq_modulus(k)) { flag_error(err, ERR_NOT_MEMBER); k = fe_zero<ModR>(); }
  return k;
}
__device__ __forceinline__ Fp load_fq_checked(const uint8_t* p, int* err) {
  Fp x = fe_load_be<ModP>(p);
  if (fe_geq_modulus(x)) { flag_error(err, ERR_NOT_MEMBER); x = fe_zero<ModP>(); }
  return fe_to_mont(x);
}
__device__ __forceinline__ G1Affine load_g1_checked(const uint8_t* p, int* err) {
  G1Affine a; a.x = load_fq_checked(p, err); a.y = load_fq_checked(p + 32, err);
  if (!g1_on_curve(a)) { flag_error(err, ERR_NOT_MEMBER); a.x = fe_zero<ModP>(); a.y = fe_zero<ModP>(); }
  return a;
}
__device__ __forceinline__ G2Affine load_g2_checked(const uint8_t* p, int* err) {
  G2Affine a;
  a.x.a = load_fq_checked(p, err); a.x.b = load_fq_checked(p + 32, err);
  a.y.a = load_fq_checked(p + 64, err); a.y.b = load_fq_checked(p + 96, err);
  if (!g2_on_curve(a)) { flag_error(err, ERR_NOT_MEMBER); a.x = fp2_zero(); a.y = fp2_zero(); }
  return a;
}
__device__ __forceinline__ void load_gt_checked(Fp12& r, const uint8_t* p, int* err) {
#pragma unroll 1
  for (int k = 0; k < 6; ++k) { f12c(r, k).a = load_fq_checked(p + 64 * k, err); f12c(r, k).b = load_fq_checked(p + 64 * k + 32, err); }
}

// G2 subgroup membership for untrusted inputs.  BN254's twist has a large cofactor (2p - r); the
// zcash-bn lineage behind rabe-bn rejects a twist point outside the order-r subgroup when a G2
// value is constructed / deserialised ([r]Q == O, surfacing as FieldError::NotMember -> RabeError,
// /root/reference/src/error.rs:60-69).  The same set is tested here with the BN endomorphism
//     [u+1]Q + psi([u]Q) + psi^2([u]Q) == psi^3([2u]Q),     psi = twist o Frobenius o untwist,
// i.e. one 63-bit scalar multiplication instead of a 254-bit one (checked against [r]Q == O in
// tests/test_oracle_pin.py and on the device in tests/test_gpu_edge_cases.py).
__device__ __forceinline__ void g2_psi(G2Xyzz& r, const G2Xyzz& p) {
  r.x = fp2_mul(fp2_conj(p.x), FROB1[2]); r.y = fp2_mul(fp2_conj(p.y), FROB1[3]);
  r.zz = fp2_conj(p.zz); r.zzz = fp2_conj(p.zzz);
}
__device__ __forceinline__ bool xyzz_equal(const G2Xyzz& a, const G2Xyzz& b) {
  const bool ia = xyzz_is_inf(a), ib = xyzz_is_inf(b);
  if (ia || ib) return ia && ib;
  return fp2_eq(fp2_mul(a.x, b.zz), fp2_mul(b.x, a.zz)) && fp2_eq(fp2_mul(a.y, b.zzz), fp2_mul(b.y, a.zzz));
}
static __device__ __noinline__ bool g2_in_subgroup(const G2Affine* q) {     // q finite and on the twist
  const uint32_t u[8] = {0x4a6909f1u, 0x44e992b4u, 0, 0, 0, 0, 0, 0};     // BN parameter u = 4965661367192848881
  G2Xyzz uq, lhs, t, rhs;
  xyzz_mul_affine(uq, *q, u, 63);
  lhs = uq; xyzz_add_affine(lhs, *q);                                      // [u+1]Q
  g2_psi(t, uq); xyzz_add(lhs, t);                                         // + psi([u]Q)
  g2_psi(rhs, t); xyzz_add(lhs, rhs);                                      // + psi^2([u]Q)
  xyzz_dbl(t, uq);                                                         // [2u]Q
  g2_psi(rhs, t); g2_psi(t, rhs); g2_psi(rhs, t);                          // psi^3
  return xyzz_equal(lhs, rhs);
}
// one thread per point; element i sits at q + stride * i (canonical bytes)
__global__ void __launch_bounds__(64) k_g2_subgroup_check(const uint8_t* __restrict__ q, size_t stride, size_t n, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G2Affine a = load_g2_checked(q + stride * i, err);                       // range + on-curve (flags on failure, yields infinity)
  if (aff_is_inf(a)) return;
  if (!g2_in_subgroup(&a)) flag_error(err, ERR_NOT_MEMBER);
}

// 16-byte vector copy helpers for table entries (tables are 64/128/384-byte aligned)
template <class T> __device__ __forceinline__ T ldg_struct(const T* p) {
  static_assert(sizeof(T) % 16 == 0, "vector load");
  T r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);     // every T here is a pure aggregate of uint32_t limbs
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); ++i) {
    uint4 v = __ldg(s + i);
    d[4 * i] = v.x; d[4 * i + 1] = v.y; d[4 * i + 2] = v.z; d[4 * i + 3] = v.w;
  }
  return r;
}

// one-thread decoders (canonical bytes -> Montgomery) used when a handle is created
__global__ void k_decode_g2(const uint8_t* in, G2Affine* out, int* err) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *out = load_g2_checked(in, err);
}
__global__ void k_decode_gt(const uint8_t* in, Fp12* out, int* err) {
  if (blockIdx.x == 0 && threadIdx.x == 0) load_gt_checked(*out, in, err);
}

// ------------------------------------------------------------------------------------------
// micro kernels: element-wise products (parity of the Montgomery core) and the dependent-chain
// benchmark that measures the achievable Fp-mul rate (roofline denominator)
template <class M>
__global__ void k_fe_mul(const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<M> x = fe_load_be<M>(a + 32 * i), y = fe_load_be<M>(b + 32 * i);
  if (fe_geq_modulus(x) || fe_geq_modulus(y)) { flag_error(err, ERR_NOT_MEMBER); x = fe_zero<M>(); }
  fe_store_be(out + 32 * i, fe_from_mont(fe_mul(fe_to_mont(x), fe_to_mont(y))));
}

template <int ILP>
__global__ void k_fq_mul_chain(const uint8_t* a, const uint8_t* b, size_t n, int iters, uint8_t* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp x[ILP], y = fe_load_be<ModP>(b + 32 * i);
#pragma unroll
  for (int j = 0; j < ILP; ++j) { x[j] = fe_load_be<ModP>(a + 32 * i); x[j].v[0] ^= j; }
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = fe_mul(x[j], y);
  }
  Fp acc = x[0];
#pragma unroll
  for (int j = 1; j < ILP; ++j) acc = acc + x[j];
  fe_store_be(out + 32 * i, acc);
}

// ------------------------------------------------------------------------------------------
// fixed-base tables.  tab[w][d] = d * 2^(W*w) * base (affine, Montgomery), d = 0 unused.
// phase 1: one thread per window computes the window base; phase 2: one thread per (w, d >= 2).
template <class F>
__device__ __forceinline__ F f_inverse(const F& x);
template <> __device__ __forceinline__ Fp f_inverse<Fp>(const Fp& x) { return fe_inv(x); }
template <> __device__ __forceinline__ Fp2 f_inverse<Fp2>(const Fp2& x) { return fp2_inv(x); }

template <class F>
__device__ __forceinline__ Affine<F> xyzz_normalize(const Xyzz<F>& p) {
  Affine<F> a;
  if (xyzz_is_inf(p)) { f_set_zero(a.x); f_set_zero(a.y); return a; }
  F inv = f_inverse<F>(f_mul(p.zz, p.zzz));
  return xyzz_to_affine_with(p, inv);
}

template <class F>
__global__ void k_table_window_bases(Affine<F> base, int W, int nwin, Affine<F>* tab, size_t stride) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwin) return;
  Xyzz<F> acc; xyzz_from_affine(acc, base);
#pragma unroll 1
  for (int i = 0; i < W * w; ++i) { Xyzz<F> t = acc; xyzz_dbl(acc, t); }
  tab[(size_t)w * stride + 1] = xyzz_normalize(acc);
  Affine<F> z; f_set_zero(z.x); f_set_zero(z.y);
  tab[(size_t)w * stride] = z;
}

template <class F>
__global__ void k_table_fill(int W, int nwin, Affine<F>* tab) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t per = (size_t)1 << W;
  if (t >= per * nwin) return;
  uint32_t d = (uint32_t)(t & (per - 1));
  size_t w = t >> W;
  if (d < 2) return;
  Affine<F> b = tab[(w << W) + 1];
  uint32_t k[8] = {d, 0, 0, 0, 0, 0, 0, 0};
  Xyzz<F> acc;
  xyzz_mul_affine(acc, b, k, W);
  tab[t] = xyzz_normalize(acc);
}

// Large-window tables (W > 12; up to 2^24 entries per window, 11.8 GB for a G1 base): one thread
// fills TABLE_CHUNK consecutive digits of one window by repeated mixed addition from
// (chunk start) * window base, then normalises its run with ONE field inversion (Montgomery trick),
// instead of a double-and-add and an inversion per entry.
constexpr int TABLE_CHUNK = 32;
template <class F>
__global__ void __launch_bounds__(64) k_table_fill_chunked(int W, int nwin, Affine<F>* tab, size_t stride, size_t count) {
  // stride: entries per window; count: digits 0 .. count-1 are filled (2^W, or 2^(W-1) + TABLE_CHUNK for signed digits)
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t chunks_per_win = count / TABLE_CHUNK;
  if (t >= chunks_per_win * nwin) return;
  const size_t w = t / chunks_per_win;
  const uint32_t d0 = (uint32_t)(t % chunks_per_win) * TABLE_CHUNK;
  Affine<F>* row = tab + w * stride;
  const Affine<F> b = row[1];                       // written by k_table_window_bases
  Xyzz<F> pts[TABLE_CHUNK];
  F pre[TABLE_CHUNK];
  Xyzz<F> acc;
  uint32_t k[8] = {d0, 0, 0, 0, 0, 0, 0, 0};
  xyzz_mul_affine(acc, b, k, W);
  F run; f_set_one(run);
#pragma unroll 1
  for (int j = 0; j < TABLE_CHUNK; ++j) {
    pts[j] = acc;
    pre[j] = run;
    if (!xyzz_is_inf(acc)) run = f_mul(run, f_mul(acc.zz, acc.zzz));
    xyzz_add_affine(acc, b);
  }
  F inv = f_inverse<F>(run);
#pragma unroll 1
  for (int j = TABLE_CHUNK - 1; j >= 0; --j) {
    const uint32_t d = d0 + (uint32_t)j;
    Affine<F> a;
    if (xyzz_is_inf(pts[j])) { f_set_zero(a.x); f_set_zero(a.y); }
    else {
      F zi = f_mul(inv, pre[j]);
      inv = f_mul(inv, f_mul(pts[j].zz, pts[j].zzz));
      a = xyzz_to_affine_with(pts[j], zi);
    }
    if (d != 1) row[d] = a;                         // d == 1 is the window base itself (being read by other threads)
  }
}

// window bases of n tables at once (thread = (table, window)); bases are canonical bytes
__global__ void k_g2_multi_window_bases(const uint8_t* __restrict__ bases, uint32_t n, int W, int nwin, G2Affine* tab, int* err) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * (uint32_t)nwin) return;
  uint32_t tb = t / nwin, w = t % nwin;
  G2Affine base = load_g2_checked(bases + 128 * (size_t)tb, err);
  G2Xyzz acc; xyzz_from_affine(acc, base);
#pragma unroll 1
  for (int i = 0; i < W * (int)w; ++i) { G2Xyzz u = acc; xyzz_dbl(acc, u); }
  tab[((size_t)t << W) + 1] = xyzz_normalize(acc);
  G2Affine z; f_set_zero(z.x); f_set_zero(z.y);
  tab[(size_t)t << W] = z;
}
__global__ void k_gt_multi_window_bases(const uint8_t* __restrict__ bases, uint32_t n, int W, int nwin, Fp12* tab, int* err) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * (uint32_t)nwin) return;
  uint32_t tb = t / nwin, w = t % nwin;
  Fp12 acc; load_gt_checked(acc, bases + 384 * (size_t)tb, err);
#pragma unroll 1
  for (int i = 0; i < W * (int)w; ++i) fp12_sqr_to(&acc, &acc);
  tab[((size_t)t << W) + 1] = acc;
  Fp12 one; fp12_set_one(one);
  tab[(size_t)t << W] = one;
}

// Gt tables: tab[w][d] = base^(d * 2^(W*w)); d = 0 holds one.
__global__ void k_gt_table_window_bases(const Fp12* base, int W, int nwin, Fp12* tab) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwin) return;
  Fp12 acc = *base;
#pragma unroll 1
  for (int i = 0; i < W * w; ++i) fp12_sqr_to(&acc, &acc);
  tab[((size_t)w << W) + 1] = acc;
  Fp12 one; fp12_set_one(one);
  tab[(size_t)w << W] = one;
}
__global__ void k_gt_table_fill(int W, int nwin, Fp12* tab) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t per = (size_t)1 << W;
  if (t >= per * nwin) return;
  uint32_t d = (uint32_t)(t & (per - 1));
  size_t w = t >> W;
  if (d < 2) return;
  Fp12 b = tab[(w << W) + 1], r;
  uint32_t k[8] = {d, 0, 0, 0, 0, 0, 0, 0};
  fp12_pow(&r, &b, k);
  tab[t] = r;
}

// large-window Gt tables: thread (w, chunk) walks its digits with one Fq12 product per entry
__global__ void __launch_bounds__(64) k_gt_table_fill_chunked(int W, int nwin, Fp12* tab) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t chunks_per_win = ((size_t)1 << W) / TABLE_CHUNK;
  if (t >= chunks_per_win * nwin) return;
  const size_t w = t / chunks_per_win;
  const uint32_t d0 = (uint32_t)(t % chunks_per_win) * TABLE_CHUNK;
  Fp12* row = tab + (w << W);
  Fp12 b = row[1], acc;
  uint32_t k[8] = {d0, 0, 0, 0, 0, 0, 0, 0};
  fp12_pow(&acc, &b, k);                            // d0 == 0 gives one
#pragma unroll 1
  for (int j = 0; j < TABLE_CHUNK; ++j) {
    const uint32_t d = d0 + (uint32_t)j;
    if (d != 1) row[d] = acc;
    fp12_mul_to(&acc, &acc, &b);
  }
}

// ------------------------------------------------------------------------------------------
// acc = k * base through the window table (k canonical limbs, < r < 2^254)
template <class F>
__device__ __forceinline__ void fixed_base_mul(Xyzz<F>& acc, const Affine<F>* __restrict__ tab, int W, int nwin, const uint32_t* k) {
  xyzz_set_inf(acc);
  uint32_t d = scalar_window(k, 0, W);
  Affine<F> e = ldg_struct(tab + d);
#pragma unroll 1
  for (int w = 0; w < nwin; ++w) {
    Affine<F> cur = e;
    uint32_t dcur = d;
    if (w + 1 < nwin) {          // fetch the next entry while this addition runs
      int bit = (w + 1) * W;
      int width = (bit + W <= 256) ? W : 256 - bit;
      d = scalar_window(k, bit, width);
      e = ldg_struct(tab + (((size_t)(w + 1)) << W) + d);
    }
    if (dcur) xyzz_add_affine(acc, cur);
  }
}

// Signed-digit walk of a G1 table (windows wider than 12 bits): the digits are recoded into [-2^(W-1), 2^(W-1)] with a
// carry into the next window, the table keeps only d = 0 .. 2^(W-1) per window (half the memory of the unsigned table at
// the same number of additions) and a negative digit adds the entry with its y negated.  k < r < 2^254 keeps the top
// window far below 2^(W-1), so the last carry is absorbed.  stride = entries per window.
__device__ __forceinline__ void fixed_base_mul_signed(G1Xyzz& acc, const G1Affine* __restrict__ tab, int W, int nwin, size_t stride, const uint32_t* k) {
  xyzz_set_inf(acc);
  const uint32_t half = 1u << (W - 1);
  uint32_t d = scalar_window(k, 0, W), carry = 0;
  bool neg = d > half;
  if (neg) { d = (1u << W) - d; carry = 1; }
  G1Affine e = ldg_struct(tab + d), first;
  first.x = fe_zero<ModP>(); first.y = fe_zero<ModP>();
#pragma unroll 1
  for (int w = 0; w < nwin; ++w) {
    G1Affine cur = e;
    const uint32_t dcur = d; const bool ncur = neg;
    if (w + 1 < nwin) {          // fetch the next entry while this addition runs
      int bit = (w + 1) * W;
      int width = (bit + W <= 256) ? W : 256 - bit;
      d = scalar_window(k, bit, width) + carry;
      neg = d > half; carry = 0;
      if (neg) { d = (1u << W) - d; carry = 1; }
      e = ldg_struct(tab + (size_t)(w + 1) * stride + d);
    }
    if (ncur) cur.y = fe_neg(cur.y);
    if (!dcur) { cur.x = fe_zero<ModP>(); cur.y = fe_zero<ModP>(); }          // digit 0: the point at infinity
    // the first two entries are both affine: 6 products instead of a copy and a 10-product mixed addition
    if (w == 0) first = cur;
    else if (w == 1) xyzz_from_two_affine(acc, first, cur);
    else xyzz_add_affine(acc, cur);
  }
  if (nwin == 1) xyzz_from_affine(acc, first);
}

// Montgomery-trick tail shared by the G1 kernels: given the running product `run` of the
// z-values zs[0..cnt) (with prefix products in pre[]), write the affine results.
template <int M>
__device__ __forceinline__ void g1_batch_store(const G1Xyzz* pts, const Fp* zs, const Fp* pre, Fp run, int cnt, uint8_t* out) {
  Fp inv = fe_inv(run);
#pragma unroll 1
  for (int j = cnt - 1; j >= 0; --j) {
    uint8_t* o = out + 64 * (size_t)j;
    if (xyzz_is_inf(pts[j])) {
      uint4 z = make_uint4(0, 0, 0, 0);
      uint4* q = reinterpret_cast<uint4*>(o);
      q[0] = z; q[1] = z; q[2] = z; q[3] = z;
      continue;
    }
    Fp zi = inv * pre[j];
    inv = inv * zs[j];
    g1_store_be(o, xyzz_to_affine_with(pts[j], zi));
  }
}

// out[i] = k[i] * base, M consecutive outputs per thread, one field inversion per thread
template <int M>
__global__ void __launch_bounds__(128, RB_G1_MINB) k_g1_mul_fixed(const G1Affine* __restrict__ tab, int W, int nwin, const uint8_t* __restrict__ k,
                                                       size_t n, uint8_t* __restrict__ out, int* err, size_t stride, int opt) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t o0 = t * opt;                       // opt <= M outputs per thread, chosen by the launch (g1_outputs_per_thread)
  if (o0 >= n) return;
  int cnt = (int)((n - o0 < (size_t)opt) ? (n - o0) : opt);
  G1Xyzz pts[M]; Fp zs[M], pre[M];
  Fp run = fe_one<ModP>();
#pragma unroll 1
  for (int j = 0; j < cnt; ++j) {
    Fr s = load_scalar(k + 32 * (o0 + j), err);
    if (stride != ((size_t)1 << W)) fixed_base_mul_signed(pts[j], tab, W, nwin, stride, s.v);
    else fixed_base_mul(pts[j], tab, W, nwin, s.v);
    Fp z = xyzz_is_inf(pts[j]) ? fe_one<ModP>() : pts[j].zz * pts[j].zzz;
    zs[j] = z; pre[j] = run; run = run * z;
  }
  g1_batch_store<M>(pts, zs, pre, run, cnt, out + 64 * o0);
}

// AC17 cp_encrypt rows (ac17/mod.rs:330-356 restructured): output o = (item, row, l) gets
//   c = g * (s0 * A[row][l][0] + s1 * A[row][l][1]),  A in Montgomery form so that the Montgomery
// product with the canonical s lands directly on the canonical scalar.
template <int M>
__global__ void __launch_bounds__(128, RB_G1_MINB) k_ac17_enc_rows(const G1Affine* __restrict__ tab, int W, int nwin, const Fr* __restrict__ A,
                                                        const uint8_t* __restrict__ s, uint32_t rows3, size_t total,
                                                        uint8_t* __restrict__ out, int* err, size_t a_item_stride, size_t stride, int opt) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t o0 = t * opt;
  if (o0 >= total) return;
  int cnt = (int)((total - o0 < (size_t)opt) ? (total - o0) : opt);
  G1Xyzz pts[M]; Fp zs[M], pre[M];
  Fp run = fe_one<ModP>();
  size_t item = o0 / rows3;
  uint32_t r = (uint32_t)(o0 - item * rows3);
  Fr s0 = load_scalar(s + 64 * item, err), s1 = load_scalar(s + 64 * item + 32, err);
#pragma unroll 1
  for (int j = 0; j < cnt; ++j) {
    const Fr* Ai = A + item * a_item_stride;          // a_item_stride == 0: one policy for the whole batch
    Fr a0 = ldg_struct(Ai + 2 * (size_t)r), a1 = ldg_struct(Ai + 2 * (size_t)r + 1);
    Fr kk = s0 * a0 + s1 * a1;
    if (stride != ((size_t)1 << W)) fixed_base_mul_signed(pts[j], tab, W, nwin, stride, kk.v);
    else fixed_base_mul(pts[j], tab, W, nwin, kk.v);
    Fp z = xyzz_is_inf(pts[j]) ? fe_one<ModP>() : pts[j].zz * pts[j].zzz;
    zs[j] = z; pre[j] = run; run = run * z;
    if (++r == rows3) {
      r = 0; ++item;
      if (j + 1 < cnt) { s0 = load_scalar(s + 64 * item, err); s1 = load_scalar(s + 64 * item + 32, err); }
    }
  }
  g1_batch_store<M>(pts, zs, pre, run, cnt, out + 64 * o0);
}

// A[i][l][t] = h_row[i][l][t] + sum_j m[i][j] * h_col[j][l][t]   (Fr, stored in Montgomery form)
// (n_pol policies of the same shape: policy p uses m + p*n1*n2, h_row + p*n1*192, h_col + p*n2*192 -- or the one
//  shared h_col: the column labels "0"+(j+1)+l+t of ac17:305-328 do not depend on the policy)
__global__ void k_ac17_fold_msp(uint32_t n1, uint32_t n2, const int8_t* __restrict__ m, const uint8_t* __restrict__ h_row,
                                const uint8_t* __restrict__ h_col, Fr* A, int* err, size_t n_pol, int h_col_shared = 0) {
  size_t tt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tt >= n_pol * n1 * 6) return;
  size_t p = tt / ((size_t)n1 * 6);
  uint32_t t = (uint32_t)(tt - p * n1 * 6);
  uint32_t i = t / 6, lt = t % 6;
  m += p * n1 * n2; h_row += 192 * p * n1;
  if (!h_col_shared) h_col += 192 * p * n2;          // shared: one [n2][3][2] column table for every policy
  Fr acc = load_scalar(h_row + 32 * (size_t)t, err);
#pragma unroll 1
  for (uint32_t j = 0; j < n2; ++j) {
    int8_t v = m[(size_t)i * n2 + j];
    if (v == 0) continue;
    if (v < -1 || v > 1) { flag_error(err, ERR_POLICY); continue; }       // an MSP entry is -1, 0 or 1 (msp.rs:86-137)
    Fr h = load_scalar(h_col + 32 * ((size_t)j * 6 + lt), err);
    acc = (v > 0) ? acc + h : acc - h;
  }
  A[tt] = fe_to_mont(acc);
}

// Several tables of one width laid out back to back (one per attribute, aw11 authority keys):
// output i uses table map[i % mod] (map == null: i % mod); mod == 0: the single table.
struct TabSel { uint32_t mod; const uint32_t* map; };
__device__ __forceinline__ size_t tab_offset(TabSel ts, size_t i, int W, int nwin) {
  if (!ts.mod) return 0;
  uint32_t j = (uint32_t)(i % ts.mod);
  if (ts.map) j = ts.map[j];
  return ((size_t)j * nwin) << W;
}

// out[i] = k[i] * base over G2 (one inversion per output)
__global__ void __launch_bounds__(128) k_g2_mul_fixed(const G2Affine* __restrict__ tab, TabSel ts, int W, int nwin, const uint8_t* __restrict__ k,
                                                       size_t n, uint8_t* __restrict__ out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = load_scalar(k + 32 * i, err);
  G2Xyzz acc;
  fixed_base_mul(acc, tab + tab_offset(ts, i, W, nwin), W, nwin, s.v);
  g2_store_be(out + 128 * i, xyzz_normalize(acc));
}

// AC17 c_0 (ac17/mod.rs:297-302): thread (item, i): h_a[i] * (s_i | s0+s1)
struct G2Tab3 { const G2Affine* t[3]; };
__global__ void __launch_bounds__(128) k_ac17_enc_c0(G2Tab3 tabs, int W, int nwin, const uint8_t* __restrict__ s, size_t B,
                                                      uint8_t* __restrict__ out, int* err) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * B) return;
  size_t item = t / 3; int i = (int)(t % 3);
  Fr k;
  if (i < 2) k = load_scalar(s + 64 * item + 32 * i, err);
  else k = load_scalar(s + 64 * item, err) + load_scalar(s + 64 * item + 32, err);
  G2Xyzz acc;
  fixed_base_mul(acc, tabs.t[i], W, nwin, k.v);
  g2_store_be(out + 128 * t, xyzz_normalize(acc));
}

// r = base^k through a Gt window table
__device__ __forceinline__ void gt_fixed_pow(Fp12* acc, bool* started, const Fp12* __restrict__ tab, int W, int nwin, const uint32_t* k, Fp12* t) {
#pragma unroll 1
  for (int w = 0; w < nwin; ++w) {
    int bit = w * W;
    int width = (bit + W <= 256) ? W : 256 - bit;
    uint32_t d = scalar_window(k, bit, width);
    if (!d) continue;
    const Fp12* e = tab + (((size_t)w) << W) + d;
    *t = ldg_struct(e);                       // t: function-scope scratch of the caller
    if (*started) fp12_mul_to(acc, acc, t);
    else { fp12_copy(acc, t); *started = true; }
  }
}
__global__ void __launch_bounds__(64) k_gt_pow_fixed(const Fp12* __restrict__ tab, TabSel ts, int W, int nwin, const uint8_t* __restrict__ k, size_t n,
                                                      uint8_t* __restrict__ out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = load_scalar(k + 32 * i, err);
  Fp12 acc, scratch; bool started = false;
  gt_fixed_pow(&acc, &started, tab + tab_offset(ts, i, W, nwin), W, nwin, s.v, &scratch);
  if (!started) fp12_set_one(acc);
  fp12_store_be(out + 384 * i, acc);
}
// AC17 c_p (ac17/mod.rs:357-368): e_gh_ka[0]^s0 * e_gh_ka[1]^s1 * msg.
// 16 threads per item: thread `part` multiplies the table entries of 4 windows of one base
// (parts 0-7: base 0, parts 8-15: base 1); the 16 partial products are then folded pairwise through
// shared memory (4 levels) and thread 0 of the item applies msg.  Critical path: 3 + 4 + 1 Fq12
// products instead of 64.
constexpr int CP_PARTS = 16;
#ifndef RB_CP_ITEMS
#define RB_CP_ITEMS 8
#endif
constexpr int CP_ITEMS_PER_BLOCK = RB_CP_ITEMS;
__global__ void __launch_bounds__(CP_PARTS * CP_ITEMS_PER_BLOCK) k_ac17_enc_cp(const Fp12* __restrict__ tab0, const Fp12* __restrict__ tab1, int W, int nwin,
                                                     const uint8_t* __restrict__ s, const uint8_t* __restrict__ msg, size_t B,
                                                     uint8_t* __restrict__ out, int* err) {
  extern __shared__ uint4 cp_smem_raw[];
  Fp12* slots = reinterpret_cast<Fp12*>(cp_smem_raw);
  const int part = threadIdx.x % CP_PARTS;
  const size_t item = (size_t)blockIdx.x * CP_ITEMS_PER_BLOCK + threadIdx.x / CP_PARTS;
  const bool live = item < B;
  Fp12* mine = slots + threadIdx.x;
  Fp12 t, m;                                   // function-scope operands (see pairing.cuh note)
  fp12_set_one(*mine);
  if (live) {
    const int base = part / (CP_PARTS / 2), sub = part % (CP_PARTS / 2);
    const int per = (nwin + CP_PARTS / 2 - 1) / (CP_PARTS / 2);
    Fr k = load_scalar(s + 64 * item + 32 * base, err);
    const Fp12* tab = base ? tab1 : tab0;
    bool started = false;
#pragma unroll 1
    for (int w = sub * per; w < nwin && w < (sub + 1) * per; ++w) {
      int bit = w * W;
      int width = (bit + W <= 256) ? W : 256 - bit;
      uint32_t d = scalar_window(k.v, bit, width);
      if (!d) continue;
      t = ldg_struct(tab + (((size_t)w) << W) + d);
      if (started) fp12_mul_to(mine, mine, &t); else { fp12_copy(mine, &t); started = true; }
    }
  }
#pragma unroll 1
  for (int stride = 1; stride < CP_PARTS; stride <<= 1) {
    __syncthreads();
    if (live && (part % (2 * stride)) == 0) fp12_mul_to(mine, mine, mine + stride);
  }
  if (live && part == 0) {
    load_gt_checked(m, msg + 384 * item, err);
    fp12_mul_to(&m, &m, mine);
    fp12_store_be(out + 384 * item, m);
  }
}

// ------------------------------------------------------------------------------------------
// Operand indexing of the element-wise kernels: element i reads operand (i / div) % mod (mod == 0:
// no wrap).  {1,0} element-wise, {1,1} one shared value, {1,n} a per-leaf table tiled over the
// batch, {n,0} one value per batch item broadcast over its n leaves.
struct OpIdx { size_t div, mod; };
__device__ __forceinline__ size_t op_index(OpIdx o, size_t i) { size_t j = i / o.div; return o.mod ? j % o.mod : j; }

// variable-base operators
__global__ void __launch_bounds__(128) k_g1_mul_var(const uint8_t* __restrict__ p, OpIdx pi, const uint8_t* __restrict__ k, OpIdx ki, size_t n,
                                                     uint8_t* __restrict__ out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine b = load_g1_checked(p + 64 * op_index(pi, i), err);
  Fr s = load_scalar(k + 32 * op_index(ki, i), err);
  G1Xyzz acc; xyzz_mul_affine(acc, b, s.v, 254);
  g1_store_be(out + 64 * i, xyzz_normalize(acc));
}
__global__ void __launch_bounds__(128) k_g2_mul_var(const uint8_t* __restrict__ p, OpIdx pi, const uint8_t* __restrict__ k, OpIdx ki, size_t n,
                                                     uint8_t* __restrict__ out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G2Affine b = load_g2_checked(p + 128 * op_index(pi, i), err);
  Fr s = load_scalar(k + 32 * op_index(ki, i), err);
  G2Xyzz acc; xyzz_mul_affine(acc, b, s.v, 254);
  g2_store_be(out + 128 * i, xyzz_normalize(acc));
}
__global__ void k_iota(uint32_t* p, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}
// out[i] = -q[i] over G2 (canonical bytes; the point at infinity stays): the `-d` of bsw/mod.rs:308 without a scalar multiplication
__global__ void k_g2_neg(const uint8_t* __restrict__ q, size_t n, uint8_t* __restrict__ out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G2Affine a = load_g2_checked(q + 128 * i, err);
  a.y = fp2_neg(a.y);
  g2_store_be(out + 128 * i, a);
}
__global__ void __launch_bounds__(64) k_gt_pow_var(const uint8_t* __restrict__ a, OpIdx ai, const uint8_t* __restrict__ k, OpIdx ki, size_t n,
                                                    uint8_t* __restrict__ out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp12 x, r; load_gt_checked(x, a + 384 * op_index(ai, i), err);
  Fr s = load_scalar(k + 32 * op_index(ki, i), err);
  fp12_pow(&r, &x, s.v);
  fp12_store_be(out + 384 * i, r);
}
__global__ void __launch_bounds__(64) k_gt_mul(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, OpIdx bi, size_t n,
                                                uint8_t* __restrict__ out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp12 x, y; load_gt_checked(x, a + 384 * i, err); load_gt_checked(y, b + 384 * op_index(bi, i), err);
  fp12_mul_to(&x, &x, &y);
  fp12_store_be(out + 384 * i, x);
}
__global__ void __launch_bounds__(64) k_gt_inverse(const uint8_t* __restrict__ a, size_t n, uint8_t* __restrict__ out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp12 x, y; load_gt_checked(x, a + 384 * i, err);
  fp12_inv_to(&y, &x);
  fp12_store_be(out + 384 * i, y);
}

// ------------------------------------------------------------------------------------------
// gather sums.  Output o adds points[base_of(o) + idx[j]] for j in its list; optionally one extra
// point (`extra`, may be null) and an optional negation; result affine in Montgomery form
// (internal) or canonical bytes.
struct GatherArgs {
  const uint8_t* points;      // canonical G1, [.][64]
  const uint32_t* idx;
  const uint32_t* offs;       // per-list offsets, or null: single shared list [0, n_idx)
  uint32_t n_idx;
  uint32_t lists_per_group;   // outputs are (group, lane): list = group (or shared), point row = idx*stride + lane
  uint32_t lanes;             // e.g. 3 for AC17's [row][3] layout
  size_t group_stride;        // points per group (0: all groups read the same point array)
  const uint8_t* extra;       // [lanes][64] canonical, added to every output of that lane (or null)
  int negate;
};
__global__ void __launch_bounds__(128) k_g1_gather_sum(GatherArgs a, size_t n_groups, G1Affine* out_mont, uint8_t* out_bytes, int* err) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_groups * a.lanes) return;
  const unsigned live = __activemask();
  size_t grp = t / a.lanes; uint32_t lane = (uint32_t)(t % a.lanes);
  uint32_t lo = a.offs ? a.offs[grp] : 0, hi = a.offs ? a.offs[grp + 1] : a.n_idx;
  const uint8_t* base = a.points + 64 * (grp * a.group_stride);
  G1Xyzz acc; xyzz_set_inf(acc);
  if (a.extra) { G1Affine e = load_g1_checked(a.extra + 64 * lane, err); xyzz_add_affine(acc, e); }
#pragma unroll 1
  for (uint32_t j = lo; j < hi; ++j) {
    G1Affine p = load_g1_checked(base + 64 * ((size_t)a.idx[j] * a.lanes + lane), err);
    xyzz_add_affine(acc, p);
  }
  // Per-item lists have different lengths inside a warp.  Without this barrier the lanes that leave the
  // loop early run the inversion below on their own, one length class after the other (measured: 8.8x the
  // warp instructions of the shared-list launch); with it the warp inverts once, together.
  __syncwarp(live);
  G1Affine r = xyzz_normalize(acc);
  if (a.negate) r = aff_neg(r);
  if (out_mont) out_mont[t] = r;
  if (out_bytes) g1_store_be(out_bytes + 64 * t, r);
}

// ------------------------------------------------------------------------------------------
// pairing products.  k_miller: one thread per pair -> Miller value (Montgomery limbs, internal);
// k_final_exp: one thread per product: multiply its Miller values, final exponentiation, optional
// extra Gt factor, canonical store.
struct MillerArgs {
  const G1Affine* p_mont;   // internal affine G1 (Montgomery) or null
  const uint8_t* p_bytes;   // canonical G1 or null
  const uint8_t* q_bytes;   // canonical G2
  size_t q_period;          // 0: q index == pair index; else q index = pair % q_period (shared G2 arguments)
  const uint32_t* p_map;    // optional: pair -> p index
  const uint32_t* q_map;    // optional: pair -> q index
  int p_single;             // every pair uses P[0]
  uint32_t out_stride, out_off;   // out index = out_stride ? pair * out_stride + out_off : pair   (lane-paired kernel only)
};
__global__ void __launch_bounds__(RB_ML_BLOCK, RB_PAIR_MINB) k_miller(MillerArgs a, size_t n_pairs, Fp12* out, int* err) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pairs) return;
  size_t pi = a.p_map ? a.p_map[t] : t;
  size_t qi = a.q_map ? a.q_map[t] : (a.q_period ? t % a.q_period : t);
  G1Affine p = a.p_mont ? a.p_mont[pi] : load_g1_checked(a.p_bytes + 64 * pi, err);
  G2Affine q = load_g2_checked(a.q_bytes + 128 * qi, err);
  Fp12 f;
  if (aff_is_inf(p) || aff_is_inf(q)) fp12_set_one(f);
  else miller_single(&f, &p, &q);
  out[t] = f;
}
__global__ void __launch_bounds__(RB_FE_BLOCK, RB_PAIR_MINB) k_final_exp(const Fp12* __restrict__ miller, const uint32_t* __restrict__ offs, uint32_t fixed_count,
                                                   size_t n_products, const uint8_t* __restrict__ extra, uint8_t* __restrict__ out, int* err) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_products) return;
  size_t lo = offs ? offs[t] : t * fixed_count, hi = offs ? offs[t + 1] : (t + 1) * fixed_count;
  Fp12 f, r, g;                                // function-scope operands (see pairing.cuh note)
  if (lo == hi) fp12_set_one(f);
  else {
    f = miller[lo];
#pragma unroll 1
    for (size_t j = lo + 1; j < hi; ++j) { g = miller[j]; fp12_mul_to(&f, &f, &g); }
  }
  final_exponentiation(&r, &f);
  if (extra) { load_gt_checked(g, extra + 384 * t, err); fp12_mul_to(&r, &r, &g); }
  fp12_store_be(out + 384 * t, r);
}

// ------------------------------------------------------------------------------------------
// helpers of the per-leaf decrypt loops (bsw/mod.rs:282-308, lsw/mod.rs:247-280; kernels in coop_kernels.cuh)
__device__ __forceinline__ void copy_bytes16(uint8_t* d, const uint8_t* s, int n16) {
  const uint4* a = reinterpret_cast<const uint4*>(s); uint4* b = reinterpret_cast<uint4*>(d);
  for (int i = 0; i < n16; ++i) b[i] = a[i];
}
// dst[i] = src[idx[i]]  (elements of 16*n16 bytes)
__global__ void k_gather_rows(const uint8_t* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t n, int n16, uint8_t* __restrict__ dst) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  copy_bytes16(dst + (size_t)16 * n16 * i, src + (size_t)16 * n16 * idx[i], n16);
}

// ------------------------------------------------------------------------------------------
// SHA3-256 -> Fr (utils/hash/mod.rs:23-31: `Fr::from_slice(Sha3_256(msg))`, big-endian integer
// reduced mod r) for a batch of byte strings: one thread per message, Keccak-f[1600] in registers.
// This is the scalar behind every `sha3_hash(g, s)` = g * H(s) of the schemes (hash/mod.rs:10-20).
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }
__device__ __forceinline__ void keccak_f1600(uint64_t* st) {
  const uint64_t RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
                           0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
                           0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
                           0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
                           0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  const int ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
  const int PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
#pragma unroll 1
  for (int round = 0; round < 24; ++round) {
    uint64_t bc[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      uint64_t t = bc[(i + 4) % 5] ^ rotl64(bc[(i + 1) % 5], 1);
#pragma unroll
      for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
    }
    uint64_t t = st[1];
#pragma unroll
    for (int i = 0; i < 24; ++i) { int j = PIL[i]; uint64_t b = st[j]; st[j] = rotl64(t, ROT[i]); t = b; }
#pragma unroll
    for (int j = 0; j < 25; j += 5) {
#pragma unroll
      for (int i = 0; i < 5; ++i) bc[i] = st[j + i];
#pragma unroll
      for (int i = 0; i < 5; ++i) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
    }
    st[0] ^= RC[round];
  }
}
__global__ void __launch_bounds__(128) k_sha3_fr(const uint8_t* __restrict__ data, const uint32_t* __restrict__ offs, size_t n, uint8_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* m = data + offs[i];
  uint32_t len = offs[i + 1] - offs[i];
  uint64_t st[25];
#pragma unroll
  for (int k = 0; k < 25; ++k) st[k] = 0;
  const uint32_t RATE = 136;                                   // SHA3-256
  uint32_t pos = 0;
#pragma unroll 1
  for (uint32_t k = 0; k < len; ++k) {
    st[pos >> 3] ^= (uint64_t)m[k] << (8 * (pos & 7));
    if (++pos == RATE) { keccak_f1600(st); pos = 0; }
  }
  st[pos >> 3] ^= (uint64_t)0x06 << (8 * (pos & 7));           // SHA3 domain separation + pad10*1
  st[(RATE - 1) >> 3] ^= (uint64_t)0x80 << (8 * ((RATE - 1) & 7));
  keccak_f1600(st);
  // digest bytes d[0..31] = little-endian bytes of st[0..3]; as a big-endian integer limb j (LSW first) = bswap of bytes 28-4j .. 31-4j
  Fr x;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int byte0 = 28 - 4 * j;                                    // most significant byte of limb j
    uint32_t w = (uint32_t)(st[byte0 >> 3] >> (8 * (byte0 & 7)));   // bytes byte0..byte0+3, little-endian in w
    x.v[j] = __byte_perm(w, 0, 0x0123);
  }
  // x < 2^256 < 6r: five conditional subtractions of r
#pragma unroll 1
  for (int k = 0; k < 5; ++k) fe_reduce_once<ModR>(x.v);
  fe_store_be(out + 32 * i, x);
}

// ------------------------------------------------------------------------------------------
// aw11::decrypt (aw11/mod.rs:327-349) pair lists and Gt factors, built on the device as canonical bytes.
// Item b, pruned leaf i (ciphertext row ci = ct_idx[i]):
//   pair 2i   = ( hn[i] , c3[b][ci] )      hn[i] = -c_i * H(gid) g1     (key side, scaled once per call)
//   pair 2i+1 = ( kc[i] , c2[b][ci] )      kc[i] =  c_i * K_i
//   G[b][i]   = c1[b][ci]                  (raised to -c_i afterwards)
struct Aw11Gather {
  const uint8_t* hn; const uint8_t* kc;              // [nI][64]
  const uint8_t* c1; const uint8_t* c2; const uint8_t* c3;   // [B][n][384] / [B][n][128] / [B][n][128]
  const uint32_t* ct_idx; uint32_t nI, n;
};
__global__ void __launch_bounds__(128) k_aw11_gather(Aw11Gather a, size_t B, uint8_t* P, uint8_t* Q, uint8_t* G) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * a.nI) return;
  size_t b = t / a.nI; uint32_t i = (uint32_t)(t % a.nI);
  size_t row = b * a.n + a.ct_idx[i];
  copy_bytes16(P + 64 * (2 * t), a.hn + 64 * (size_t)i, 4);
  copy_bytes16(Q + 128 * (2 * t), a.c3 + 128 * row, 8);
  copy_bytes16(P + 64 * (2 * t + 1), a.kc + 64 * (size_t)i, 4);
  copy_bytes16(Q + 128 * (2 * t + 1), a.c2 + 128 * row, 8);
  copy_bytes16(G + 384 * t, a.c1 + 384 * row, 24);
}
// out[b] = base[b] * prod_{i < cnt} f[b][i]   (Gt, canonical bytes)
__global__ void __launch_bounds__(64) k_gt_prod_rows(const uint8_t* __restrict__ f, uint32_t cnt, const uint8_t* __restrict__ base, size_t B,
                                                      uint8_t* __restrict__ out, int* err) {
  size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  Fp12 acc, x;                                       // function scope (pairing_body.inc note)
  load_gt_checked(acc, base + 384 * b, err);
#pragma unroll 1
  for (uint32_t i = 0; i < cnt; ++i) { load_gt_checked(x, f + 384 * (b * cnt + i), err); fp12_mul_to(&acc, &acc, &x); }
  fp12_store_be(out + 384 * b, acc);
}

// ------------------------------------------------------------------------------------------
// AC17 setup (ac17/mod.rs:141-188), one-off: a single thread walks the reference statements.
// rnd = rho_g, rho_h, a0, b0, a1, b1, k0, k1, k2 (canonical Fr).
__global__ void k_ac17_setup(const uint8_t* __restrict__ rnd, uint8_t* __restrict__ pk, uint8_t* __restrict__ msk, int* err) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Fr rho_g = load_scalar(rnd, err), rho_h = load_scalar(rnd + 32, err);
  Fr a[2], b[2], kk[3];
  for (int i = 0; i < 2; ++i) { a[i] = load_scalar(rnd + 64 + 64 * i, err); b[i] = load_scalar(rnd + 96 + 64 * i, err); }
  for (int i = 0; i < 3; ++i) kk[i] = load_scalar(rnd + 192 + 32 * i, err);
  G1Affine g1gen; g1gen.x = fe_one<ModP>(); g1gen.y = fe_dbl(fe_one<ModP>());
  G2Affine g2gen; g2gen.x = G2_GEN_X; g2gen.y = G2_GEN_Y;
  G1Xyzz t1; G2Xyzz t2;
  xyzz_mul_affine(t1, g1gen, rho_g.v, 254); G1Affine g = xyzz_normalize(t1);
  xyzz_mul_affine(t2, g2gen, rho_h.v, 254); G2Affine h = xyzz_normalize(t2);
  g1_store_be(pk, g); g1_store_be(msk, g); g2_store_be(msk + 64, h);
  for (int i = 0; i < 2; ++i) { xyzz_mul_affine(t2, h, a[i].v, 254); g2_store_be(pk + 64 + 128 * i, xyzz_normalize(t2)); }
  g2_store_be(pk + 64 + 256, h);
  for (int i = 0; i < 3; ++i) { xyzz_mul_affine(t1, g, kk[i].v, 254); g1_store_be(msk + 192 + 64 * i, xyzz_normalize(t1)); }
  Fp12 f, e_gh, r;
  if (aff_is_inf(g) || aff_is_inf(h)) fp12_set_one(e_gh);
  else { miller_single(&f, &g, &h); final_exponentiation(&e_gh, &f); }
  Fr k2m = fe_to_mont(kk[2]);
  for (int i = 0; i < 2; ++i) {
    Fr ex = fe_from_mont(fe_to_mont(kk[i]) * fe_to_mont(a[i]) + k2m);
    fp12_pow(&r, &e_gh, ex.v);
    fp12_store_be(pk + 448 + 384 * i, r);
  }
  for (int i = 0; i < 2; ++i) { fe_store_be(msk + 384 + 32 * i, a[i]); fe_store_be(msk + 448 + 32 * i, b[i]); }
}

// msk-derived constants kept on the device: 1/a_t (Montgomery) and b_t (Montgomery)
struct Ac17MskConsts { Fr a_inv[2]; Fr b[2]; };
__global__ void k_ac17_msk_consts(const uint8_t* __restrict__ msk, Ac17MskConsts* out, int* err) {
  if (blockIdx.x != 0 || threadIdx.x >= 2) return;
  int t = threadIdx.x;
  Fr a = fe_to_mont(load_scalar(msk + 384 + 32 * t, err));
  out->a_inv[t] = fe_inv(a);                       // ac17/mod.rs:229,249 `a[_t].inverse().unwrap()`
  out->b[t] = fe_to_mont(load_scalar(msk + 448 + 32 * t, err));
}

// AC17 cp_keygen scalars (ac17/mod.rs:206-261 restructured).  Thread (key, x), x in [0, n]:
//   x < n : row of attribute x      sc[key][x][t<2] = (sum_l H[x][l][t]*br[l] + sigma_x)/a_t ; sc[..][2] = -sigma_x
//   x == n: the k_p row, with H = h_01 and sigma = rnd[n+2]; this thread also writes br -> sc_k0[key][3]
// rnd per key: r0, r1, sigma_attr[0..n), sigma.
__global__ void __launch_bounds__(128) k_ac17_keygen_scalars(const Ac17MskConsts* __restrict__ mc, uint32_t n, const uint8_t* __restrict__ h_attr,
                                                              const uint8_t* __restrict__ h_01, const uint8_t* __restrict__ rnd, size_t B,
                                                              uint8_t* __restrict__ sc, uint8_t* __restrict__ sc_k0, int* err) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * (n + 1)) return;
  size_t key = t / (n + 1); uint32_t x = (uint32_t)(t % (n + 1));
  const uint8_t* rk = rnd + 32 * key * (n + 3);
  Fr r0 = fe_to_mont(load_scalar(rk, err)), r1 = fe_to_mont(load_scalar(rk + 32, err));
  Fr br[3] = {mc->b[0] * r0, mc->b[1] * r1, r0 + r1};
  const uint8_t* H = (x < n) ? h_attr + 192 * (size_t)x : h_01;
  Fr sigma = fe_to_mont(load_scalar(rk + 64 + 32 * (size_t)x, err));   // x == n lands on the final sigma
  uint8_t* o = sc + 96 * t;
  for (int tt = 0; tt < 2; ++tt) {
    Fr acc = sigma;
    for (int l = 0; l < 3; ++l) acc = acc + fe_to_mont(load_scalar(H + 32 * (l * 2 + tt), err)) * br[l];
    fe_store_be(o + 32 * tt, fe_from_mont(acc * mc->a_inv[tt]));
  }
  fe_store_be(o + 64, fe_from_mont(fe_neg(sigma)));
  if (x == n) for (int i = 0; i < 3; ++i) fe_store_be(sc_k0 + 96 * key + 32 * i, fe_from_mont(br[i]));
}

// ------------------------------------------------------------------------------------------
// Fr element-wise operators (`Fr + Fr`, `-`, `*`, `.inverse()`, `.neg()`: secretsharing/mod.rs:25-28,
// 66,218; bsw/mod.rs:103,147; lsw/mod.rs:94,143-152,204-206).  op: 0 add, 1 sub, 2 mul, 3 inverse(a), 4 neg(a)
__global__ void __launch_bounds__(128) k_fr_op(int op, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, OpIdx bi, size_t n,
                                                uint8_t* __restrict__ out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr x = load_scalar(a + 32 * i, err), r;
  Fr y = (op <= 2) ? load_scalar(b + 32 * op_index(bi, i), err) : fe_zero<ModR>();
  switch (op) {
    case 0: r = x + y; break;
    case 1: r = x - y; break;
    case 2: r = fe_from_mont(fe_to_mont(x) * fe_to_mont(y)); break;
    case 3: r = fe_is_zero(x) ? x : fe_from_mont(fe_inv(fe_to_mont(x))); if (fe_is_zero(x)) flag_error(err, ERR_NOT_MEMBER); break;
    default: r = fe_neg(x); break;
  }
  fe_store_be(out + 32 * i, r);
}

// element-wise group additions (`G1 + G1`, `G2 + G2`: bsw/mod.rs:147-148,197-198; aw11/mod.rs:224,276)
__global__ void __launch_bounds__(128) k_g1_add(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, OpIdx bi, size_t n,
                                                 uint8_t* __restrict__ out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p = load_g1_checked(a + 64 * i, err), q = load_g1_checked(b + 64 * op_index(bi, i), err);
  G1Xyzz acc; xyzz_from_affine(acc, p); xyzz_add_affine(acc, q);
  g1_store_be(out + 64 * i, xyzz_normalize(acc));
}
__global__ void __launch_bounds__(128) k_g2_add(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, OpIdx bi, size_t n,
                                                 uint8_t* __restrict__ out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G2Affine p = load_g2_checked(a + 128 * i, err), q = load_g2_checked(b + 128 * op_index(bi, i), err);
  G2Xyzz acc; xyzz_from_affine(acc, p); xyzz_add_affine(acc, q);
  g2_store_be(out + 128 * i, xyzz_normalize(acc));
}

// lsw::encrypt scalars (lsw/mod.rs:193-206).  draws [B][n] are the n values the reference pushes as sx[1..n].
// Its `sx[0] = sx[0] - sx[_i]` runs first with _i = 0 (sx[0] = secret - secret = 0) and then subtracts
// sx[1..n-1]; ej[i] uses sx[i] for i < n.  So sx[0] = -(sx[1] + ... + sx[n-1]), sx[i] = draws[i-1], and the
// last draw is never used -- reproduced here, not repaired.
// thread (b, i):  k1 = H_i * secret_b   k2 = sx_i   k3 = sx_i * H_i      (canonical Fr)
__global__ void __launch_bounds__(128) k_lsw_enc_scalars(const uint8_t* __restrict__ secret, const uint8_t* __restrict__ draws,
                                                          const uint8_t* __restrict__ attr_hash, uint32_t n, size_t B,
                                                          uint8_t* __restrict__ k1, uint8_t* __restrict__ k2, uint8_t* __restrict__ k3, int* err) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * n) return;
  size_t b = t / n; uint32_t i = (uint32_t)(t % n);
  Fr h = fe_to_mont(load_scalar(attr_hash + 32 * (size_t)i, err));
  Fr sec = load_scalar(secret + 32 * b, err);
  Fr sx;
  if (i == 0) {
    sx = fe_zero<ModR>();
#pragma unroll 1
    for (uint32_t j = 1; j < n; ++j) sx = sx - load_scalar(draws + 32 * (b * n + j - 1), err);
  } else {
    sx = load_scalar(draws + 32 * (b * n + i - 1), err);
  }
  fe_store_be(k1 + 32 * t, sec * h);                 // canonical * Montgomery -> canonical
  fe_store_be(k2 + 32 * t, sx);
  fe_store_be(k3 + 32 * t, sx * h);
}

// ------------------------------------------------------------------------------------------
// Secret sharing over a policy tree (secretsharing/mod.rs:82-141,215-221).  The share of a leaf is
//   secret + sum over the AND gates on its root path of  sum_{i=1}^{k-1} a_{gate,i} * (child+1)^i
// so a policy is flattened on the host into per-leaf term lists (coefficient index, x, i); the
// powers x^i are computed here once per plan, the shares per (item, leaf).
struct ShareTerm { uint32_t coef; uint32_t x; uint32_t e; uint32_t pad; };
__global__ void k_share_consts(const ShareTerm* __restrict__ terms, uint32_t n_terms, Fr* consts) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_terms) return;
  Fr x = fe_to_mont(Fr{{terms[t].x, 0, 0, 0, 0, 0, 0, 0}});
  Fr acc = fe_one<ModR>();
  uint32_t e = terms[t].e;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) { acc = fe_sqr(acc); if ((e >> bit) & 1u) acc = acc * x; }
  consts[t] = acc;                                   // Montgomery form of x^e
}
// shares[item][leaf] = secret[item] + sum_t coeffs[item][terms[t].coef] * consts[t],  t in leaf's range
__global__ void __launch_bounds__(128) k_shares(const ShareTerm* __restrict__ terms, const Fr* __restrict__ consts,
                                                 const uint32_t* __restrict__ leaf_offs, uint32_t n_leaves, uint32_t n_coefs,
                                                 const uint8_t* __restrict__ secret, const uint8_t* __restrict__ coeffs, size_t B,
                                                 uint8_t* __restrict__ out, int* err) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * n_leaves) return;
  size_t item = t / n_leaves; uint32_t leaf = (uint32_t)(t % n_leaves);
  Fr acc = load_scalar(secret + 32 * item, err);
#pragma unroll 1
  for (uint32_t j = leaf_offs[leaf]; j < leaf_offs[leaf + 1]; ++j) {
    Fr a = load_scalar(coeffs + 32 * (item * n_coefs + terms[j].coef), err);
    acc = acc + a * consts[j];                       // canonical * Montgomery -> canonical
  }
  fe_store_be(out + 32 * t, acc);
}

// Reconstruction coefficients (secretsharing/mod.rs:9-72): the coefficient of a leaf is the product
// over the AND gates on its path of the Lagrange-at-0 weight of its child position among the
// points 1..k.  terms: (x = child+1, e = k) per AND gate on the path.
__global__ void __launch_bounds__(128) k_lagrange_coeffs(const ShareTerm* __restrict__ terms, const uint32_t* __restrict__ leaf_offs,
                                                          uint32_t n_leaves, uint8_t* __restrict__ out) {
  uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= n_leaves) return;
  Fr acc = fe_one<ModR>();
#pragma unroll 1
  for (uint32_t j = leaf_offs[leaf]; j < leaf_offs[leaf + 1]; ++j) {
    uint32_t xi = terms[j].x, k = terms[j].e;
    Fr num = fe_one<ModR>(), den = fe_one<ModR>();
    Fr fxi = fe_to_mont(Fr{{xi, 0, 0, 0, 0, 0, 0, 0}});
#pragma unroll 1
    for (uint32_t xj = 1; xj <= k; ++xj) {
      if (xj == xi) continue;
      Fr fxj = fe_to_mont(Fr{{xj, 0, 0, 0, 0, 0, 0, 0}});
      num = num * fe_neg(fxj);                        // (0 - x_j)
      den = den * (fxi - fxj);                        // (x_i - x_j)
    }
    acc = acc * num * fe_inv(den);
  }
  fe_store_be(out + 32 * leaf, fe_from_mont(acc));
}

// AC17 kp_keygen scalars (ac17/mod.rs:450-541 restructured).  Thread (key, row i):
//   sc[key][i][t<2] = (sum_l H_row[i][l][t] br_l + sigma_i)/a_t
//                     + sum_{j'=1}^{c-1} w_{i,j'} (sum_l H_col[j'-1][l][t] br_l / a_t - sigma'_{j'-1}),   w_{i,j'} = sum_{j>=j'} M_ij
//   sc[key][i][2]   = -sigma_i - sum_{j=1}^{c-1} M_ij sigma'_{j-1}
// (w is the suffix sum because the reference's `_temp`, declared before its `_j` loop at :491, accumulates
// across j.)  The group part M_i0 * g_k[t] is added by k_ac17_kp_finish.
// rnd per key: r0, r1, sigma'[0..c-2], sigma_attr[0..n1).
__global__ void __launch_bounds__(128) k_ac17_kp_keygen_scalars(const Ac17MskConsts* __restrict__ mc, uint32_t n1, uint32_t c, const int8_t* __restrict__ m,
                                                                 const uint8_t* __restrict__ h_row, const uint8_t* __restrict__ h_col,
                                                                 const uint8_t* __restrict__ rnd, size_t B, uint8_t* __restrict__ sc,
                                                                 uint8_t* __restrict__ sc_k0, int* err) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * n1) return;
  size_t key = t / n1; uint32_t i = (uint32_t)(t % n1);
  const size_t per_key = 2 + (size_t)(c - 1) + n1;
  const uint8_t* rk = rnd + 32 * key * per_key;
  Fr r0 = fe_to_mont(load_scalar(rk, err)), r1 = fe_to_mont(load_scalar(rk + 32, err));
  Fr br[3] = {mc->b[0] * r0, mc->b[1] * r1, r0 + r1};
  Fr sigma = fe_to_mont(load_scalar(rk + 32 * (2 + (size_t)(c - 1) + i), err));
  uint8_t* o = sc + 96 * t;
#pragma unroll 1
  for (int tt = 0; tt < 2; ++tt) {
    Fr u[3] = {br[0] * mc->a_inv[tt], br[1] * mc->a_inv[tt], br[2] * mc->a_inv[tt]};
    Fr acc = sigma * mc->a_inv[tt];
    for (int l = 0; l < 3; ++l) acc = acc + fe_to_mont(load_scalar(h_row + 32 * ((size_t)i * 6 + l * 2 + tt), err)) * u[l];
    int w = 0;
#pragma unroll 1
    for (uint32_t j = c - 1; j >= 1; --j) {
      w += m[(size_t)i * c + j];
      if (w == 0) continue;
      Fr term = fe_neg(fe_to_mont(load_scalar(rk + 32 * (2 + (size_t)(j - 1)), err)));
      for (int l = 0; l < 3; ++l) term = term + fe_to_mont(load_scalar(h_col + 32 * ((size_t)(j - 1) * 6 + l * 2 + tt), err)) * u[l];
      Fr wf = fe_to_mont(Fr{{(uint32_t)(w < 0 ? -w : w), 0, 0, 0, 0, 0, 0, 0}});
      term = term * wf;
      acc = (w > 0) ? acc + term : acc - term;
    }
    fe_store_be(o + 32 * tt, fe_from_mont(acc));
  }
  Fr acc = fe_neg(sigma);
#pragma unroll 1
  for (uint32_t j = 1; j < c; ++j) {
    int8_t v = m[(size_t)i * c + j];
    if (v == 0) continue;
    Fr sp = fe_to_mont(load_scalar(rk + 32 * (2 + (size_t)(j - 1)), err));
    acc = (v > 0) ? acc - sp : acc + sp;
  }
  fe_store_be(o + 64, fe_from_mont(acc));
  if (i == 0) for (int x = 0; x < 3; ++x) fe_store_be(sc_k0 + 96 * key + 32 * x, fe_from_mont(br[x]));
}
// k[key][i][t] = pts[key][i][t] + M_i0 * g_k[t]
__global__ void __launch_bounds__(128) k_ac17_kp_finish(const uint8_t* __restrict__ pts, const uint8_t* __restrict__ g_k, uint32_t n1, uint32_t c,
                                                         const int8_t* __restrict__ m, size_t B, uint8_t* __restrict__ out, int* err) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * n1 * 3) return;
  uint32_t lane = (uint32_t)(t % 3); uint32_t i = (uint32_t)((t / 3) % n1);
  G1Affine p = load_g1_checked(pts + 64 * t, err);
  int8_t v = m[(size_t)i * c];
  G1Xyzz acc; xyzz_from_affine(acc, p);
  if (v != 0) {
    G1Affine gk = load_g1_checked(g_k + 64 * lane, err);
    if (v < 0) gk = aff_neg(gk);
    xyzz_add_affine(acc, gk);
  }
  g1_store_be(out + 64 * t, xyzz_normalize(acc));
}

}  // namespace rb
