// Development probe (not part of the product): issue rates of the instructions a 254-bit
// Montgomery product could be built from on sm_100a -- IMAD.WIDE.U32 (what fp.cuh uses),
// 32-bit IMAD lo / hi, DFMA (52-bit-limb floating-point products) and IADD3 -- each as 8
// independent dependent-chains per thread at full occupancy.  Prints G ops/s per kind.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void probe(int iters, uint64_t* out, uint32_t seed) {
  uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 12345u;
  uint64_t acc[8]; double d[8]; uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[j] = a + j; d[j] = 1.0 + j * 1e-9; w[j] = a ^ j; }
  double da = 1.0000001, db = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (KIND == 0) acc[j] = (uint64_t)a * (uint32_t)(acc[j] >> 7) + acc[j];                 // IMAD.WIDE.U32
        if (KIND == 1) w[j] = w[j] * a + b;                                                      // IMAD (lo)
        if (KIND == 2) w[j] = __umulhi(w[j], a) + b;                                             // IMAD.HI
        if (KIND == 3) d[j] = fma(d[j], da, db);                                                 // DFMA
        if (KIND == 4) w[j] = w[j] + a + b;                                                      // IADD3
      }
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += acc[j] + (uint64_t)w[j] + (uint64_t)__double_as_longlong(d[j]);
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND> void run(const char* name, uint64_t* out) {
  const int blocks = 148 * 8, threads = 256, iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<KIND><<<blocks, threads>>>(10, out, 1);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); probe<KIND><<<blocks, threads>>>(iters, out, rep); cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double ops = (double)blocks * threads * iters * 64.0;
  printf("{\"probe\": \"%s\", \"ms\": %.3f, \"gops\": %.1f, \"per_sm_per_clk_at_1965MHz\": %.2f}\n", name, best, ops / best / 1e6, ops / (best * 1e-3) / 148 / 1.965e9);
}

int main() {
  uint64_t* out; cudaMalloc(&out, (size_t)148 * 8 * 256 * 8);
  run<0>("imad_wide_u32", out); run<1>("imad_lo", out); run<2>("imad_hi", out); run<3>("dfma", out); run<4>("iadd3", out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
