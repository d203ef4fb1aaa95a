#include <cstdio>
#include <cuda_runtime.h>
#include "../../rabe_b200/csrc/pairing.cuh"
using namespace rb;
#define DT_FN __device__ __forceinline__
#include "devtest_body.h"
__global__ void k(uint8_t* out) { run_tests(out); }
int main() {
  uint8_t* d; cudaMalloc(&d, 384 * N_SLOTS); cudaMemset(d, 0, 384 * N_SLOTS);
  static uint8_t out[384 * N_SLOTS];
  cudaDeviceSetLimit(cudaLimitStackSize, 64 * 1024);
  k<<<1, 1>>>(d);
  cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) printf("ERR %s\n", cudaGetErrorString(e));
  cudaMemcpy(out, d, sizeof out, cudaMemcpyDeviceToHost);
  for (int s = 0; s < N_SLOTS; ++s) { printf("%2d ", s); for (int i = 0; i < 16; ++i) printf("%02x", out[384 * s + i]); printf("..."); for (int i = 368; i < 384; ++i) printf("%02x", out[384 * s + i]); printf("\n"); }
}
