// development probe: G2 3*Q through several code shapes
#include <cstdio>
#include <cuda_runtime.h>
#include "../../rabe_b200/csrc/kernels.cuh"
using namespace rb;

__global__ void k_var(int nbits, uint32_t kval, uint8_t* out) {
  G2Affine b; b.x = G2_GEN_X; b.y = G2_GEN_Y;
  uint32_t k[8] = {kval, 0, 0, 0, 0, 0, 0, 0};
  G2Xyzz acc; xyzz_mul_affine(acc, b, k, nbits);
  g2_store_be(out, xyzz_normalize(acc));
}
__global__ void k_var254(uint32_t kval, uint8_t* out) {
  G2Affine b; b.x = G2_GEN_X; b.y = G2_GEN_Y;
  uint32_t k[8] = {kval, 0, 0, 0, 0, 0, 0, 0};
  G2Xyzz acc; xyzz_mul_affine(acc, b, k, 254);
  g2_store_be(out, xyzz_normalize(acc));
}
__global__ void k_manual(uint8_t* out) {
  G2Affine b; b.x = G2_GEN_X; b.y = G2_GEN_Y;
  G2Xyzz acc; xyzz_dbl_affine(acc, b); xyzz_add_affine(acc, b);
  g2_store_be(out, xyzz_normalize(acc));
}
int main() {
  uint8_t* d; cudaMalloc(&d, 128 * 8); uint8_t h[128 * 8];
  int* err; cudaMalloc(&err, 4); cudaMemset(err, 0, 4);
  cudaDeviceSetLimit(cudaLimitStackSize, 32 * 1024);
  k_var<<<1, 1>>>(8, 3, d); k_var<<<1, 1>>>(254, 3, d + 128); k_var254<<<1, 1>>>(3, d + 256); k_manual<<<1, 1>>>(d + 384);
  // through the product kernel
  uint8_t gen[128], kk[32] = {0}; kk[31] = 3;
  cudaMemcpy(gen, d, 0, cudaMemcpyDeviceToHost);
  uint8_t *dp, *dk; cudaMalloc(&dp, 128); cudaMalloc(&dk, 32);
  k_var<<<1, 1>>>(8, 1, dp); cudaMemcpy(dk, kk, 32, cudaMemcpyHostToDevice);
  k_g2_mul_var<<<1, 128>>>(dp, dk, 1, d + 512, err);
  cudaError_t e = cudaDeviceSynchronize(); printf("sync: %s\n", cudaGetErrorString(e));
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  const char* names[] = {"var(8)", "var(254 rt)", "var254 const", "manual", "k_g2_mul_var"};
  for (int v = 0; v < 5; ++v) { printf("%-14s ", names[v]); for (int i = 0; i < 16; ++i) printf("%02x", h[128 * v + i]); printf("\n"); }
  return 0;
}
