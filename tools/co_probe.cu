// Development probe (not part of the product): saturated and single-batch throughput of the two-lane (coop.cuh)
// pairing primitives -- Fq12 product / squaring / cyclotomic squaring / product by two lines / the whole shared-
// accumulator Miller loop -- at several occupancies, against the plain Fq product chain of the same run.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo [-DRB_CO_HOT_INLINE] -o build/co_probe tools/co_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../rabe_b200/csrc/coop.cuh"

using namespace rb;

__device__ __forceinline__ Fp seed_fp(uint32_t s) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = (s * 2654435761u + i * 40503u) ^ (s >> 3);
  r.v[7] &= 0x0fffffffu;
  return r;
}

__global__ void __launch_bounds__(128) p_fq(int iters, uint32_t* out, const MillerLine*) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  Fp a = seed_fp(tid), b = seed_fp(tid + 77), c = seed_fp(tid + 99), d = seed_fp(tid + 5);
#pragma unroll 1
  for (int i = 0; i < iters; ++i) { a = a * b; c = c * d; b = b * a; d = d * c; }
  Fp s = a + b + c + d;
  out[tid] = s.v[0];
}

template <int OP, int MINB>
__global__ void __launch_bounds__(128, MINB) p_co(int iters, uint32_t* out, const MillerLine* lines) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  co::Fp12 f, g;
#pragma unroll
  for (int k = 0; k < 6; ++k) { co::f12c(f, k).v = seed_fp(tid * 6 + k); co::f12c(g, k).v = seed_fp(tid * 6 + k + 1000); }
  co::Fp2 l0 = {seed_fp(tid + 1)}, l3 = {seed_fp(tid + 2)}, l4 = {seed_fp(tid + 3)}, m0 = {seed_fp(tid + 4)}, m3 = {seed_fp(tid + 5)}, m4 = {seed_fp(tid + 6)};
  if (OP == 4) {
    G1Affine pv, pf; pv.x = seed_fp(tid >> 1); pv.y = seed_fp((tid >> 1) + 9); pf.x = seed_fp((tid >> 1) + 19); pf.y = seed_fp((tid >> 1) + 29);
    co::G2Affine q; q.x.v = seed_fp(tid + 100); q.y.v = seed_fp(tid + 200);
#pragma unroll 1
    for (int i = 0; i < iters; ++i) { co::miller_pair(&f, &pv, &q, &pf, lines); q.x = co::f12c(f, 1); }
  } else {
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
      if (OP == 0) co::fp12_mul_to(&f, &f, &g);
      if (OP == 1) co::fp12_sqr_to(&f, &f);
      if (OP == 2) co::fp12_cyclotomic_sqr_to(&f, &f);
      if (OP == 3) co::fp12_mul_by_line_pair(&f, &l0, &l3, &l4, &m0, &m3, &m4);
    }
  }
  out[tid] = co::f12c(f, 0).v.v[0] ^ co::f12c(f, 5).v.v[1];
}

static double g_fq_rate = 0;
static const char* OPN[] = {"mul", "sqr", "cyc_sqr", "line_pair", "miller_pair"};
static const double OPC[] = {54, 36, 18, 69, 11763};

template <typename K>
static void run(int op, int minb, K kern, int blocks, int iters, uint32_t* out, const MillerLine* lines, const char* tag, int bs = 128) {
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
  int resident = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, 128, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, bs>>>(1, out, lines); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); kern<<<blocks, bs>>>(iters, out, lines); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  cudaError_t e = cudaDeviceSynchronize();
  double rate = (double)blocks * (bs / 2) * iters * OPC[op] / (best * 1e-3);
  printf("{\"layout\": \"co\", \"op\": \"%s\", \"grid\": \"%s\", \"minb\": %d, \"regs\": %d, \"local_bytes\": %zu, \"resident_blocks\": %d, \"blocks\": %d, \"ms\": %.3f, \"gfpmul_s\": %.2f, \"frac\": %.3f%s}\n",
         OPN[op], tag, minb, fa.numRegs, (size_t)fa.localSizeBytes, resident, blocks, best, rate / 1e9, rate / g_fq_rate, e == cudaSuccess ? "" : ", \"error\": true");
  fflush(stdout);
}

int main() {
  uint32_t* out; cudaMalloc(&out, (size_t)148 * 64 * 128 * 4);
  MillerLine* lines; cudaMalloc(&lines, sizeof(MillerLine) * MILLER_LINES); cudaMemset(lines, 0x11, sizeof(MillerLine) * MILLER_LINES);
  cudaDeviceSetLimit(cudaLimitStackSize, 32 * 1024);
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 16, iters = 500;
    p_fq<<<blocks, 128>>>(2, out, lines); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0); p_fq<<<blocks, 128>>>(iters, out, lines); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    g_fq_rate = (double)blocks * 128 * iters * 4 / (best * 1e-3);
    printf("{\"layout\": \"fq\", \"op\": \"mul\", \"gfpmul_s\": %.2f}\n", g_fq_rate / 1e9);
  }
  // "sat": one full wave at the kernel's occupancy; "b4096": the grid of one 4096-item AC17 decrypt (12288 terms = 192 blocks)
#define CO(OP, MB, IT) run(OP, MB, p_co<OP, MB>, 148 * MB, IT, out, lines, "sat"); run(OP, MB, p_co<OP, MB>, 192, IT, out, lines, "b4096")
  CO(0, 1, 300); CO(0, 2, 300); CO(0, 3, 300); CO(0, 4, 300);
  CO(1, 2, 300); CO(1, 4, 300);
  CO(2, 2, 300); CO(2, 4, 300);
  CO(3, 1, 300); CO(3, 2, 300); CO(3, 3, 300); CO(3, 4, 300);
  CO(4, 1, 2); CO(4, 2, 2); CO(4, 3, 2);
  // one 4096-item decrypt (24576 threads) in smaller blocks: finer distribution over the 148 SMs
  run(4, 1, p_co<4, 1>, 384, 2, out, lines, "b4096/64", 64); run(4, 1, p_co<4, 1>, 768, 2, out, lines, "b4096/32", 32);
  run(3, 1, p_co<3, 1>, 384, 300, out, lines, "b4096/64", 64); run(3, 1, p_co<3, 1>, 768, 300, out, lines, "b4096/32", 32);
  run(0, 1, p_co<0, 1>, 384, 300, out, lines, "b4096/64", 64); run(0, 1, p_co<0, 1>, 768, 300, out, lines, "b4096/32", 32);
  run(2, 2, p_co<2, 2>, 384, 300, out, lines, "b4096/64", 64); run(2, 2, p_co<2, 2>, 768, 300, out, lines, "b4096/32", 32);
  // the final exponentiation's grid: 8192 threads
  run(2, 2, p_co<2, 2>, 64, 300, out, lines, "fe4096/128", 128); run(2, 2, p_co<2, 2>, 128, 300, out, lines, "fe4096/64", 64); run(2, 2, p_co<2, 2>, 256, 300, out, lines, "fe4096/32", 32);
  run(0, 1, p_co<0, 1>, 64, 300, out, lines, "fe4096/128", 128); run(0, 1, p_co<0, 1>, 256, 300, out, lines, "fe4096/32", 32);
  return 0;
}
