/* lzb_sha256.cuh -- SHA-256 compression function (FIPS 180-4), host and device.
 *
 * Replaces the reference's Sha256.c (Sha256_Init / Sha256_Update / Sha256_Final) where the xz path uses
 * it: the block check with id 10 (Xz.c:45-88) and the index-record digest of XzUnpacker_Code
 * (XzDec.c:646-652,694-703,805-812).  Payload bytes are hashed on the GPU (one lane per unit, fused
 * after the decode like the CRCs); the few framing bytes (index records, final padding block) use the
 * same routine on the host.
 */
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define LZB_HD __host__ __device__ inline
#else
#define LZB_HD inline
#endif

namespace lzb {

#define LZB_SHA256_K_VALUES                                                                       \
  0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u,     \
  0xab1c5ed5u, 0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu,     \
  0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu,     \
  0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u,     \
  0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu,     \
  0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, 0xa2bfe8a1u, 0xa81a664bu,     \
  0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u,     \
  0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,     \
  0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u,     \
  0xc67178f2u

#if defined(__CUDACC__)
static __device__ __constant__ uint32_t c_sha256_k[64] = { LZB_SHA256_K_VALUES };
#endif
static const uint32_t h_sha256_k[64] = { LZB_SHA256_K_VALUES };

LZB_HD uint32_t sha_rotr(uint32_t x, unsigned n) { return (x >> n) | (x << (32 - n)); }

LZB_HD void sha256_init(uint32_t s[8])
{
  s[0] = 0x6a09e667u; s[1] = 0xbb67ae85u; s[2] = 0x3c6ef372u; s[3] = 0xa54ff53au;
  s[4] = 0x510e527fu; s[5] = 0x9b05688cu; s[6] = 0x1f83d9abu; s[7] = 0x5be0cd19u;
}

/* one 64-byte block; a 16-word rolling message schedule */
LZB_HD void sha256_block(uint32_t s[8], const uint8_t *p)
{
#if defined(__CUDA_ARCH__)
  const uint32_t *K = c_sha256_k;
#else
  const uint32_t *K = h_sha256_k;
#endif
  uint32_t w[16];
  for (int i = 0; i < 16; i++)
    w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
  uint32_t a = s[0], b = s[1], c = s[2], d = s[3], e = s[4], f = s[5], g = s[6], h = s[7];
  for (int i = 0; i < 64; i++)
  {
    uint32_t wi;
    if (i < 16) wi = w[i];
    else
    {
      const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      const uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
      const uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
      wi = w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
    }
    const uint32_t t1 = h + (sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + wi;
    const uint32_t t2 = (sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  s[0] += a; s[1] += b; s[2] += c; s[3] += d; s[4] += e; s[5] += f; s[6] += g; s[7] += h;
}

/* digest of `n` bytes that continue a message of `before` bytes hashed into `s` so far (before % 64 == 0):
 * whole blocks, then the padded tail.  out: 32 bytes, big-endian words. */
LZB_HD void sha256_finish(uint32_t s[8], const uint8_t *p, uint64_t n, uint64_t before, uint8_t out[32])
{
  uint64_t i = 0;
  for (; i + 64 <= n; i += 64) sha256_block(s, p + i);
  uint8_t tail[128];
  const unsigned rem = (unsigned)(n - i);
  for (unsigned k = 0; k < rem; k++) tail[k] = p[i + k];
  tail[rem] = 0x80;
  const unsigned total = rem + 1 + 8 <= 64 ? 64 : 128;
  for (unsigned k = rem + 1; k < total - 8; k++) tail[k] = 0;
  const uint64_t bits = (before + n) << 3;
  for (int k = 0; k < 8; k++) tail[total - 1 - k] = (uint8_t)(bits >> (8 * k));
  sha256_block(s, tail);
  if (total == 128) sha256_block(s, tail + 64);
  for (int k = 0; k < 8; k++)
  {
    out[4 * k] = (uint8_t)(s[k] >> 24); out[4 * k + 1] = (uint8_t)(s[k] >> 16);
    out[4 * k + 2] = (uint8_t)(s[k] >> 8); out[4 * k + 3] = (uint8_t)s[k];
  }
}

}  // namespace lzb
