#include <cstdio>
#include <cuda_runtime.h>
#include "../../rabe_b200/csrc/kernels.cuh"
using namespace rb;

__global__ void k_gen(uint8_t* out) { G2Affine b; b.x = G2_GEN_X; b.y = G2_GEN_Y; g2_store_be(out, b); }

// V1: exact product body, no launch bounds / restrict
__global__ void v1(const uint8_t* p, const uint8_t* k, size_t n, uint8_t* out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  G2Affine b = load_g2_checked(p + 128 * i, err);
  Fr s = load_scalar(k + 32 * i, err);
  G2Xyzz acc; xyzz_mul_affine(acc, b, s.v, 254);
  g2_store_be(out + 128 * i, xyzz_normalize(acc));
}
// V2: unchecked load
__global__ void v2(const uint8_t* p, const uint8_t* k, size_t n, uint8_t* out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  G2Affine b = g2_load_be(p + 128 * i);
  Fr s = load_scalar(k + 32 * i, err);
  G2Xyzz acc; xyzz_mul_affine(acc, b, s.v, 254);
  g2_store_be(out + 128 * i, xyzz_normalize(acc));
}
// V3: checked load, constant scalar
__global__ void v3(const uint8_t* p, const uint8_t* k, size_t n, uint8_t* out, int* err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  G2Affine b = load_g2_checked(p + 128 * i, err);
  uint32_t kk[8] = {3, 0, 0, 0, 0, 0, 0, 0};
  G2Xyzz acc; xyzz_mul_affine(acc, b, kk, 254);
  g2_store_be(out + 128 * i, xyzz_normalize(acc));
}
// V4: checked load, manual dbl+add, dump intermediate
__global__ void v4(const uint8_t* p, uint8_t* out, int* err) {
  G2Affine b = load_g2_checked(p, err);
  G2Xyzz acc; xyzz_dbl_affine(acc, b); 
  g2_store_be(out + 128, xyzz_normalize(acc));
  xyzz_add_affine(acc, b);
  g2_store_be(out, xyzz_normalize(acc));
  g2_store_be(out + 256, b);
}
// V5: load, then compare limbs with constants
__global__ void v5(const uint8_t* p, uint32_t* out, int* err) {
  G2Affine b = load_g2_checked(p, err);
  Fp2 gx = G2_GEN_X, gy = G2_GEN_Y;
  out[0] = fp2_eq(b.x, gx) ? 1 : 0; out[1] = fp2_eq(b.y, gy) ? 1 : 0;
  for (int i = 0; i < 8; ++i) { out[2 + i] = b.x.a.v[i]; out[10 + i] = gx.a.v[i]; }
}
int main() {
  uint8_t* d; cudaMalloc(&d, 128 * 16); cudaMemset(d, 0, 128 * 16); uint8_t h[128 * 16];
  int* err; cudaMalloc(&err, 4); cudaMemset(err, 0, 4);
  uint8_t *dp, *dk; cudaMalloc(&dp, 128); cudaMalloc(&dk, 32);
  uint8_t kk[32] = {0}; kk[31] = 3; cudaMemcpy(dk, kk, 32, cudaMemcpyHostToDevice);
  k_gen<<<1, 1>>>(dp);
  v1<<<1, 128>>>(dp, dk, 1, d, err);
  v2<<<1, 128>>>(dp, dk, 1, d + 128, err);
  v3<<<1, 128>>>(dp, dk, 1, d + 256, err);
  v4<<<1, 1>>>(dp, d + 384, err);
  k_g2_mul_var<<<1, 128>>>(dp, dk, 1, d + 768, err);
  uint32_t* du; cudaMalloc(&du, 4 * 32); v5<<<1, 1>>>(dp, du, err);
  cudaError_t e = cudaDeviceSynchronize(); printf("sync: %s\n", cudaGetErrorString(e));
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  const char* names[] = {"v1 plain", "v2 unchecked", "v3 const k", "v4 3q", "v4 2q", "v4 b", "product"};
  for (int v = 0; v < 7; ++v) { printf("%-14s ", names[v]); for (int i = 0; i < 16; ++i) printf("%02x", h[128 * v + i]); printf("\n"); }
  uint32_t hu[32]; cudaMemcpy(hu, du, sizeof hu, cudaMemcpyDeviceToHost);
  printf("eq x %u y %u\n", hu[0], hu[1]);
  for (int i = 0; i < 8; ++i) printf("%08x %08x\n", hu[2 + i], hu[10 + i]);
  int herr; cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost); printf("err flag %d\n", herr);
  return 0;
}
