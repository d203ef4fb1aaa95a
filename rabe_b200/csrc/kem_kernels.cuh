// KEM tail on the device: rabe's `encrypt_symmetric` / `decrypt_symmetric`
// (/root/reference/src/utils/aes/mod.rs:10-55) for a batch of items -- key = SHA3-256 of the canonical 384-byte Gt
// encoding (`kdf`, :47-55), AES-256-GCM with a 12-byte nonce and no associated data, wire form nonce | ciphertext | tag.
// One thread per item: Keccak-f[1600] in registers (kernels.cuh), AES with the S-box staged in shared memory, GHASH by
// shift-and-add in GF(2^128).  Byte-stream work: bound by the item's own length, not by the integer pipe; it exists so
// that a bulk payload never has to leave the device between the pairing and the cipher.  Bit-exact against
// `cryptography`'s AESGCM in tests/test_gpu_kem.py.
#pragma once
#include "kernels.cuh"

namespace rb {

__constant__ uint8_t AES_SBOX[256] = {
    0x63, 0x7c, 0x77, 0x7b, 0xf2, 0x6b, 0x6f, 0xc5, 0x30, 0x01, 0x67, 0x2b, 0xfe, 0xd7, 0xab, 0x76, 0xca, 0x82, 0xc9, 0x7d, 0xfa, 0x59, 0x47, 0xf0, 0xad, 0xd4, 0xa2, 0xaf, 0x9c, 0xa4, 0x72, 0xc0,
    0xb7, 0xfd, 0x93, 0x26, 0x36, 0x3f, 0xf7, 0xcc, 0x34, 0xa5, 0xe5, 0xf1, 0x71, 0xd8, 0x31, 0x15, 0x04, 0xc7, 0x23, 0xc3, 0x18, 0x96, 0x05, 0x9a, 0x07, 0x12, 0x80, 0xe2, 0xeb, 0x27, 0xb2, 0x75,
    0x09, 0x83, 0x2c, 0x1a, 0x1b, 0x6e, 0x5a, 0xa0, 0x52, 0x3b, 0xd6, 0xb3, 0x29, 0xe3, 0x2f, 0x84, 0x53, 0xd1, 0x00, 0xed, 0x20, 0xfc, 0xb1, 0x5b, 0x6a, 0xcb, 0xbe, 0x39, 0x4a, 0x4c, 0x58, 0xcf,
    0xd0, 0xef, 0xaa, 0xfb, 0x43, 0x4d, 0x33, 0x85, 0x45, 0xf9, 0x02, 0x7f, 0x50, 0x3c, 0x9f, 0xa8, 0x51, 0xa3, 0x40, 0x8f, 0x92, 0x9d, 0x38, 0xf5, 0xbc, 0xb6, 0xda, 0x21, 0x10, 0xff, 0xf3, 0xd2,
    0xcd, 0x0c, 0x13, 0xec, 0x5f, 0x97, 0x44, 0x17, 0xc4, 0xa7, 0x7e, 0x3d, 0x64, 0x5d, 0x19, 0x73, 0x60, 0x81, 0x4f, 0xdc, 0x22, 0x2a, 0x90, 0x88, 0x46, 0xee, 0xb8, 0x14, 0xde, 0x5e, 0x0b, 0xdb,
    0xe0, 0x32, 0x3a, 0x0a, 0x49, 0x06, 0x24, 0x5c, 0xc2, 0xd3, 0xac, 0x62, 0x91, 0x95, 0xe4, 0x79, 0xe7, 0xc8, 0x37, 0x6d, 0x8d, 0xd5, 0x4e, 0xa9, 0x6c, 0x56, 0xf4, 0xea, 0x65, 0x7a, 0xae, 0x08,
    0xba, 0x78, 0x25, 0x2e, 0x1c, 0xa6, 0xb4, 0xc6, 0xe8, 0xdd, 0x74, 0x1f, 0x4b, 0xbd, 0x8b, 0x8a, 0x70, 0x3e, 0xb5, 0x66, 0x48, 0x03, 0xf6, 0x0e, 0x61, 0x35, 0x57, 0xb9, 0x86, 0xc1, 0x1d, 0x9e,
    0xe1, 0xf8, 0x98, 0x11, 0x69, 0xd9, 0x8e, 0x94, 0x9b, 0x1e, 0x87, 0xe9, 0xce, 0x55, 0x28, 0xdf, 0x8c, 0xa1, 0x89, 0x0d, 0xbf, 0xe6, 0x42, 0x68, 0x41, 0x99, 0x2d, 0x0f, 0xb0, 0x54, 0xbb, 0x16};

// SHA3-256 of n bytes (n arbitrary), digest as 32 bytes
__device__ __forceinline__ void sha3_256_bytes(const uint8_t* m, uint32_t len, uint8_t* digest) {
  uint64_t st[25];
#pragma unroll
  for (int k = 0; k < 25; ++k) st[k] = 0;
  const uint32_t RATE = 136;
  uint32_t pos = 0;
#pragma unroll 1
  for (uint32_t k = 0; k < len; ++k) {
    st[pos >> 3] ^= (uint64_t)m[k] << (8 * (pos & 7));
    if (++pos == RATE) { keccak_f1600(st); pos = 0; }
  }
  st[pos >> 3] ^= (uint64_t)0x06 << (8 * (pos & 7));
  st[(RATE - 1) >> 3] ^= (uint64_t)0x80 << (8 * ((RATE - 1) & 7));
  keccak_f1600(st);
#pragma unroll
  for (int k = 0; k < 32; ++k) digest[k] = (uint8_t)(st[k >> 3] >> (8 * (k & 7)));
}

struct Aes256 { uint32_t rk[60]; };      // round keys, big-endian words (FIPS 197)

__device__ __forceinline__ uint32_t aes_subword(const uint8_t* sb, uint32_t w) {
  return ((uint32_t)sb[w >> 24] << 24) | ((uint32_t)sb[(w >> 16) & 255] << 16) | ((uint32_t)sb[(w >> 8) & 255] << 8) | sb[w & 255];
}
__device__ __forceinline__ void aes256_expand(Aes256& a, const uint8_t* key, const uint8_t* sb) {
#pragma unroll
  for (int i = 0; i < 8; ++i) a.rk[i] = ((uint32_t)key[4 * i] << 24) | ((uint32_t)key[4 * i + 1] << 16) | ((uint32_t)key[4 * i + 2] << 8) | key[4 * i + 3];
  uint32_t rcon = 1;
#pragma unroll 1
  for (int i = 8; i < 60; ++i) {
    uint32_t t = a.rk[i - 1];
    if ((i & 7) == 0) { t = aes_subword(sb, (t << 8) | (t >> 24)) ^ (rcon << 24); rcon = (rcon << 1) ^ ((rcon & 0x80) ? 0x11b : 0); }
    else if ((i & 7) == 4) t = aes_subword(sb, t);
    a.rk[i] = a.rk[i - 8] ^ t;
  }
}
__device__ __forceinline__ uint32_t aes_xtime4(uint32_t x) {     // GF(2^8) doubling of four packed bytes
  return ((x & 0x7f7f7f7fu) << 1) ^ (((x >> 7) & 0x01010101u) * 0x1bu);
}
// one block: in/out as four big-endian column words
__device__ __forceinline__ void aes256_encrypt(const Aes256& a, const uint8_t* sb, uint32_t* s) {
#pragma unroll
  for (int c = 0; c < 4; ++c) s[c] ^= a.rk[c];
#pragma unroll 1
  for (int r = 1; r <= 14; ++r) {
    uint32_t t[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)      // SubBytes + ShiftRows: row i of column c comes from column c + i
      t[c] = ((uint32_t)sb[s[c] >> 24] << 24) | ((uint32_t)sb[(s[(c + 1) & 3] >> 16) & 255] << 16) | ((uint32_t)sb[(s[(c + 2) & 3] >> 8) & 255] << 8) | sb[s[(c + 3) & 3] & 255];
    if (r < 14) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {  // MixColumns on the packed column (b0 b1 b2 b3 from the top byte down)
        const uint32_t x = t[c], r1 = (x << 8) | (x >> 24), r2 = (x << 16) | (x >> 16), r3 = (x << 24) | (x >> 8);
        t[c] = aes_xtime4(x ^ r1) ^ r1 ^ r2 ^ r3;
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) s[c] = t[c] ^ a.rk[4 * r + c];
  }
}
// y = y * h in GF(2^128), GCM bit order (NIST SP 800-38D, algorithm 1); words big-endian, word 0 = leftmost
__device__ __forceinline__ void ghash_mul(uint32_t* y, const uint32_t* h) {
  uint32_t z[4] = {0, 0, 0, 0}, v[4] = {h[0], h[1], h[2], h[3]};
#pragma unroll 1
  for (int i = 0; i < 128; ++i) {
    const uint32_t bit = (y[i >> 5] >> (31 - (i & 31))) & 1u, m = 0u - bit;
    z[0] ^= v[0] & m; z[1] ^= v[1] & m; z[2] ^= v[2] & m; z[3] ^= v[3] & m;
    const uint32_t lsb = v[3] & 1u;
    v[3] = (v[3] >> 1) | (v[2] << 31); v[2] = (v[2] >> 1) | (v[1] << 31); v[1] = (v[1] >> 1) | (v[0] << 31); v[0] >>= 1;
    v[0] ^= (0u - lsb) & 0xe1000000u;
  }
  y[0] = z[0]; y[1] = z[1]; y[2] = z[2]; y[3] = z[3];
}
__device__ __forceinline__ uint32_t load_be32_partial(const uint8_t* p, uint32_t n) {      // n <= 4 bytes, zero padded
  uint32_t w = 0;
  for (uint32_t i = 0; i < n; ++i) w |= (uint32_t)p[i] << (24 - 8 * i);
  return w;
}

// item b: key = SHA3-256(gt[b]); decrypt == 0: out = nonce | AES-GCM(data) | tag ; decrypt != 0: in = nonce | ct | tag, out = plaintext,
// ok[b] = 1 iff the tag verifies (the plaintext of a forged item is zeroed); plaintext b is written at out + offs[b].  offs: [B+1] byte offsets of the INPUT blobs.
__global__ void __launch_bounds__(128) k_kem_aes256gcm(const uint8_t* __restrict__ gt, const uint8_t* __restrict__ nonce, const uint8_t* __restrict__ in,
                                                      const uint32_t* __restrict__ offs, size_t B, int decrypt, uint8_t* __restrict__ out, int* __restrict__ ok) {
  __shared__ uint8_t sb[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sb[i] = AES_SBOX[i];
  __syncthreads();
  const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  uint8_t key[32];
  sha3_256_bytes(gt + 384 * b, 384, key);
  Aes256 a; aes256_expand(a, key, sb);
  const uint8_t* src = in + offs[b];
  uint32_t len = offs[b + 1] - offs[b];
  const uint8_t* iv; uint8_t* dst;
  if (decrypt) {
    if (len < 28) { ok[b] = 0; return; }
    iv = src; src += 12; len -= 28; dst = out + offs[b];          // plaintext b at the offset of its blob: items stay independent
  } else {
    iv = nonce + 12 * b; dst = out + (offs[b] + 28 * b);
    for (int i = 0; i < 12; ++i) dst[i] = iv[i];
    dst += 12;
  }
  uint32_t h[4] = {0, 0, 0, 0};
  aes256_encrypt(a, sb, h);                                   // H = E_K(0)
  uint32_t j0[4] = {load_be32_partial(iv, 4), load_be32_partial(iv + 4, 4), load_be32_partial(iv + 8, 4), 1u};
  uint32_t y[4] = {0, 0, 0, 0};
  const uint32_t nblk = (len + 15) / 16;
#pragma unroll 1
  for (uint32_t blk = 0; blk < nblk; ++blk) {
    uint32_t ks[4] = {j0[0], j0[1], j0[2], j0[3] + 1 + blk};   // inc32 of the counter
    aes256_encrypt(a, sb, ks);
    const uint32_t n = (len - 16 * blk < 16) ? len - 16 * blk : 16;
    uint32_t c[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint32_t have = (n > 4u * w) ? ((n - 4u * w < 4u) ? n - 4u * w : 4u) : 0u;
      const uint32_t inw = load_be32_partial(src + 16 * blk + 4 * w, have);
      const uint32_t mask = have == 4 ? 0xffffffffu : (have == 0 ? 0u : ~(0xffffffffu >> (8 * have)));
      const uint32_t outw = (inw ^ ks[w]) & mask;
      c[w] = decrypt ? inw : outw;                             // GHASH runs over the CIPHERTEXT
      for (uint32_t i = 0; i < have; ++i) dst[16 * blk + 4 * w + i] = (uint8_t)(outw >> (24 - 8 * i));
    }
    y[0] ^= c[0]; y[1] ^= c[1]; y[2] ^= c[2]; y[3] ^= c[3];
    ghash_mul(y, h);
  }
  // length block: 64-bit bit lengths of the (empty) associated data and of the ciphertext
  y[3] ^= len << 3; y[2] ^= len >> 29;
  ghash_mul(y, h);
  aes256_encrypt(a, sb, j0);
  uint32_t tag[4] = {y[0] ^ j0[0], y[1] ^ j0[1], y[2] ^ j0[2], y[3] ^ j0[3]};
  if (decrypt) {
    uint32_t diff = 0;
    for (int w = 0; w < 4; ++w) diff |= tag[w] ^ load_be32_partial(src + len + 4 * w, 4);
    ok[b] = diff == 0;
    if (diff) for (uint32_t i = 0; i < len; ++i) dst[i] = 0;
  } else {
    for (int w = 0; w < 4; ++w) for (int i = 0; i < 4; ++i) dst[len + 4 * w + i] = (uint8_t)(tag[w] >> (24 - 8 * i));
  }
}

}  // namespace rb
