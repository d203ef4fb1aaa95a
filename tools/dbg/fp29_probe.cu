// development probe: carry-free 9 x 29-bit signed-limb Montgomery product vs the 8 x 32 carry-chain one
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../rabe_b200/csrc/fp.cuh"
using namespace rb;

struct F29 { int32_t v[9]; };
__device__ __constant__ int32_t N29[9] = {0x187cfd47, 0x10460b6, 0x1c72a34f, 0x2d522d0, 0x1585d978, 0x2db40c0, 0xa6e141, 0xe5c2634, 0x30644e};
#define NINV29 0x1a866389u   // placeholder, set below by host? (computed offline)

template <int DUMMY>
__device__ __forceinline__ F29 mul29(const F29& a, const F29& b, uint32_t ninv) {
  int64_t t[18];
#pragma unroll
  for (int k = 0; k < 18; ++k) t[k] = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < 9; ++j) t[i + j] += (int64_t)a.v[i] * b.v[j];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    uint32_t m = ((uint32_t)t[i] * ninv) & 0x1fffffffu;
#pragma unroll
    for (int j = 0; j < 9; ++j) t[i + j] += (int64_t)(int32_t)m * N29[j];
    t[i + 1] += t[i] >> 29;
  }
  F29 r;
#pragma unroll
  for (int k = 9; k < 17; ++k) { r.v[k - 9] = (int32_t)(t[k] & 0x1fffffff); t[k + 1] += t[k] >> 29; }
  r.v[8] = (int32_t)t[17];
  return r;
}

template <int ILP>
__global__ void k29(const int32_t* in, int iters, uint32_t ninv, int32_t* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  F29 x[ILP], y;
  for (int k = 0; k < 9; ++k) { y.v[k] = in[k] ^ (int)(i & 7); for (int j = 0; j < ILP; ++j) x[j].v[k] = in[9 + k] + j; }
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = mul29<0>(x[j], y, ninv);
  }
  int32_t acc = 0;
  for (int j = 0; j < ILP; ++j) for (int k = 0; k < 9; ++k) acc ^= x[j].v[k];
  out[i] = acc;
}
template <int ILP>
__global__ void k32(const uint32_t* in, int iters, uint32_t* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Fp x[ILP], y;
  for (int k = 0; k < 8; ++k) { y.v[k] = in[k]; for (int j = 0; j < ILP; ++j) x[j].v[k] = in[8 + k] ^ j; }
  y.v[7] &= 0x0fffffff;
  for (int j = 0; j < ILP; ++j) x[j].v[7] &= 0x0fffffff;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = fe_mul(x[j], y);
  }
  uint32_t acc = 0;
  for (int j = 0; j < ILP; ++j) for (int k = 0; k < 8; ++k) acc ^= x[j].v[k];
  out[i] = acc;
}

template <class F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e9f;
  for (int r = 0; r < 3; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
  return best;
}

int main() {
  int32_t h[18]; for (int k = 0; k < 18; ++k) h[k] = 0x0abcdef1 + 977 * k;
  h[8] = 0x123456; h[17] = 0x234567;
  int32_t *din, *dout; cudaMalloc(&din, sizeof h); cudaMemcpy(din, h, sizeof h, cudaMemcpyHostToDevice);
  cudaMalloc(&dout, 4 * 148 * 2048);
  const int iters = 2000;
  int wps[] = {1, 2, 4, 16};
  for (int wi = 0; wi < 4; ++wi) {
    int threads = 592 * 32 * wps[wi], blocks = threads / 128;
    float t;
    t = timeit([&] { k32<1><<<blocks, 128>>>((uint32_t*)din, iters, (uint32_t*)dout); });
    printf("warps/smsp %2d  sat32 ilp1 %7.1f G/s", wps[wi], threads * (double)iters * 1 / t / 1e6);
    t = timeit([&] { k32<2><<<blocks, 128>>>((uint32_t*)din, iters, (uint32_t*)dout); });
    printf("  ilp2 %7.1f", threads * (double)iters * 2 / t / 1e6);
    t = timeit([&] { k29<1><<<blocks, 128>>>(din, iters, 0x1a866389u, dout); });
    printf(" | unsat29 ilp1 %7.1f", threads * (double)iters * 1 / t / 1e6);
    t = timeit([&] { k29<2><<<blocks, 128>>>(din, iters, 0x1a866389u, dout); });
    printf("  ilp2 %7.1f", threads * (double)iters * 2 / t / 1e6);
    t = timeit([&] { k29<3><<<blocks, 128>>>(din, iters, 0x1a866389u, dout); });
    printf("  ilp3 %7.1f G/s\n", threads * (double)iters * 3 / t / 1e6);
  }
  return 0;
}
