// Development probe (not part of the product): saturated throughput of the Fq12 primitives of the three pairing
// layouts -- one thread per value (tower.cuh), two lanes (coop.cuh), six lanes (wide.cuh) -- as chains of dependent
// operations per work item at several occupancies (launch bounds), against the plain Fq product chain of the same
// run.  Prints one JSON line per (primitive, layout, min blocks per SM); `frac` = algorithmic Fq products (one-thread
// Karatsuba tower counts: mul 54, sqr 36, cyclotomic sqr 18, line 39) per second / the Fq product rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xptxas -v -o build/tower_probe tools/tower_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../rabe_b200/csrc/kernels.cuh"
#include "../rabe_b200/csrc/coop_kernels.cuh"
#include "../rabe_b200/csrc/wide_kernels.cuh"

using namespace rb;

__device__ __forceinline__ Fp seed_fp(uint32_t s) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = (s * 2654435761u + i * 40503u) ^ (s >> 3);
  r.v[7] &= 0x0fffffffu;
  return r;
}

// ---- plain Fq product chain (denominator)
__global__ void __launch_bounds__(128) p_fq(int iters, uint32_t* out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  Fp a = seed_fp(tid), b = seed_fp(tid + 77), c = seed_fp(tid + 99), d = seed_fp(tid + 5);
#pragma unroll 1
  for (int i = 0; i < iters; ++i) { a = a * b; c = c * d; b = b * a; d = d * c; }
  Fp s = a + b + c + d;
  out[tid] = s.v[0];
}

// ---- six lanes
template <int OP, int MINB>
__global__ void __launch_bounds__(128, MINB) p_w6(int iters, uint32_t* out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const W6Slot w = w6_slot((size_t)1 << 40);
  Fp2 f = {seed_fp(tid), seed_fp(tid + 1)}, g = {seed_fp(tid + 2), seed_fp(tid + 3)};
  const Fp2 l0 = {seed_fp(w.item * 3), seed_fp(w.item * 3 + 1)}, l3 = {seed_fp(w.item * 5), seed_fp(w.item * 5 + 1)}, l4 = {seed_fp(w.item * 7), seed_fp(w.item * 7 + 1)};
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) f = w6::mul(w.L, f, g);
    if (OP == 1) f = w6::sqr(w.L, f);
    if (OP == 2) f = w6::cyclotomic_sqr(w.L, f);
    if (OP == 3) f = w6::mul_line(w.L, f, l0, l3, l4);
  }
  out[tid] = f.a.v[0] ^ f.b.v[1];
}

// ---- two lanes
template <int OP, int MINB>
__global__ void __launch_bounds__(128, MINB) p_co(int iters, uint32_t* out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  co::Fp12 f, g;
#pragma unroll
  for (int k = 0; k < 6; ++k) { co::f12c(f, k).v = seed_fp(tid * 6 + k); co::f12c(g, k).v = seed_fp(tid * 6 + k + 1000); }
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) co::fp12_mul_to(&f, &f, &g);
    if (OP == 1) co::fp12_sqr_to(&f, &f);
    if (OP == 2) co::fp12_cyclotomic_sqr_to(&f, &f);
  }
  out[tid] = co::f12c(f, 0).v.v[0] ^ co::f12c(f, 5).v.v[1];
}

// ---- one thread
template <int OP, int MINB>
__global__ void __launch_bounds__(128, MINB) p_one(int iters, uint32_t* out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  Fp12 f, g;
#pragma unroll
  for (int k = 0; k < 6; ++k) { f12c(f, k) = {seed_fp(tid * 12 + k), seed_fp(tid * 12 + k + 6)}; f12c(g, k) = {seed_fp(tid * 12 + k + 100), seed_fp(tid * 12 + k + 106)}; }
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) fp12_mul_to(&f, &f, &g);
    if (OP == 1) fp12_sqr_to(&f, &f);
    if (OP == 2) fp12_cyclotomic_sqr_to(&f, &f);
  }
  out[tid] = f12c(f, 0).a.v[0] ^ f12c(f, 5).b.v[1];
}

static double g_fq_rate = 0;
static const char* OPN[] = {"mul", "sqr", "cyc_sqr", "line"};
static const double OPC[] = {54, 36, 18, 39};

template <typename K>
static void run(const char* layout, int op, int minb, K kern, int lanes_per_item, int items_per_warp_x, int blocks_per_sm, int iters, uint32_t* out) {
  // blocks_per_sm resident blocks of 128 threads per SM, one wave
  const int blocks = 148 * blocks_per_sm;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
  int resident = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, 128, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, 128>>>(2, out); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); kern<<<blocks, 128>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  cudaError_t e = cudaDeviceSynchronize();
  // work items: threads / lanes_per_item (six-lane: 5 items per warp)
  double items = (lanes_per_item == 6) ? (double)blocks * 4 * 5 : (double)blocks * 128 / lanes_per_item;
  double rate = items * iters * OPC[op] / (best * 1e-3);
  printf("{\"layout\": \"%s\", \"op\": \"%s\", \"minb\": %d, \"regs\": %d, \"local_bytes\": %zu, \"resident_blocks\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, \"gfpmul_s\": %.2f, \"frac\": %.3f%s}\n",
         layout, OPN[op], minb, fa.numRegs, (size_t)fa.localSizeBytes, resident, blocks_per_sm, best, rate / 1e9, g_fq_rate > 0 ? rate / g_fq_rate : 0.0,
         e == cudaSuccess ? "" : ", \"error\": true");
  fflush(stdout);
}

int main() {
  uint32_t* out; cudaMalloc(&out, (size_t)148 * 64 * 128 * 4);
  cudaDeviceSetLimit(cudaLimitStackSize, 32 * 1024);
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 16, iters = 500;
    p_fq<<<blocks, 128>>>(2, out); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0); p_fq<<<blocks, 128>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    g_fq_rate = (double)blocks * 128 * iters * 4 / (best * 1e-3);
    printf("{\"layout\": \"fq\", \"op\": \"mul\", \"gfpmul_s\": %.2f}\n", g_fq_rate / 1e9);
  }
  const int IT = 300;
#define W6(OP, MB) run("w6", OP, MB, p_w6<OP, MB>, 6, 5, MB, IT, out)
#define CO(OP, MB) run("co", OP, MB, p_co<OP, MB>, 2, 0, MB, IT, out)
#define ONE(OP, MB) run("one", OP, MB, p_one<OP, MB>, 1, 0, MB, IT, out)
  W6(0, 1); W6(0, 2); W6(0, 3); W6(0, 4); W6(0, 5); W6(0, 6); W6(0, 8);
  W6(1, 2); W6(1, 4);
  W6(2, 2); W6(2, 3); W6(2, 4); W6(2, 6); W6(2, 8);
  W6(3, 2); W6(3, 3); W6(3, 4); W6(3, 6); W6(3, 8);
  CO(0, 1); CO(0, 2); CO(0, 3); CO(0, 4); CO(0, 6);
  CO(1, 2); CO(1, 3); CO(1, 4);
  CO(2, 2); CO(2, 3); CO(2, 4);
  ONE(0, 2); ONE(0, 3); ONE(0, 4);
  ONE(1, 2); ONE(2, 2);
  return 0;
}
