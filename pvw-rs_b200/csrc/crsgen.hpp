// crsgen.hpp -- host-side seeded CRS generation: PvwCrs::new_deterministic / new_from_tag (src/params/crs.rs:45-90).
//
// The reference expands a 32-byte master seed with ChaCha8 into one 32-byte seed per matrix element and hands each to
// fhe-math's Poly::random_from_seed.  fhe-math 0.1.0-beta.7 and rand 0.8.5 are not in the build image; their behaviour
// is restated here from memory -- PARITY UNPINNED, like psi:
//   * `rng.gen::<[u8; 32]>()` samples each byte as `next_u32() as u8` (rand 0.8 `Standard` for arrays / u8);
//   * Poly::random_from_seed(seed): prng = ChaCha8Rng::from_seed(SHA-256(seed)); every RNS row is filled with
//     Uniform(0..q_j) samples (rand 0.8.5 UniformInt<u64>::sample: widening multiply, zone = MAX - (MAX - q + 1) % q);
//   * new_from_tag: Rust's DefaultHasher (SipHash-1-3, zero keys) over the bytes of tag + "CRS" followed by 0xff, the
//     u64 repeated four times (little endian) as the master seed.
// Generation is host work in the reference too; the matrix is then uploaded like any other CRS.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "hostparams.hpp"

namespace pvw {

// ---- SHA-256 (FIPS 180-4) ----------------------------------------------------------------------------------------
inline void sha256(const uint8_t* msg, size_t len, uint8_t out[32]) {
  static const uint32_t K[64] = {
      0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
      0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
      0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
      0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
      0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
      0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  std::vector<uint8_t> m(msg, msg + len);
  m.push_back(0x80);
  while (m.size() % 64 != 56) m.push_back(0);
  const uint64_t bits = (uint64_t)len * 8;
  for (int i = 7; i >= 0; i--) m.push_back((uint8_t)(bits >> (8 * i)));
  auto rotr = [](uint32_t x, int n) { return (x >> n) | (x << (32 - n)); };
  for (size_t off = 0; off < m.size(); off += 64) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
      w[i] = (uint32_t)m[off + 4 * i] << 24 | (uint32_t)m[off + 4 * i + 1] << 16 | (uint32_t)m[off + 4 * i + 2] << 8 | m[off + 4 * i + 3];
    for (int i = 16; i < 64; i++) {
      const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
      const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g), t1 = hh + S1 + ch + K[i] + w[i];
      const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c), t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  for (int i = 0; i < 8; i++) {
    out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16); out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i];
  }
}

// ---- SipHash-1-3 with zero keys = Rust std::collections::hash_map::DefaultHasher::new() ------------------------------
inline uint64_t siphash13_zero_key(const uint8_t* msg, size_t len) {
  auto rotl = [](uint64_t x, int b) { return (x << b) | (x >> (64 - b)); };
  uint64_t v0 = 0x736f6d6570736575ull, v1 = 0x646f72616e646f6dull, v2 = 0x6c7967656e657261ull, v3 = 0x7465646279746573ull;
  auto round = [&] {
    v0 += v1; v1 = rotl(v1, 13); v1 ^= v0; v0 = rotl(v0, 32);
    v2 += v3; v3 = rotl(v3, 16); v3 ^= v2;
    v0 += v3; v3 = rotl(v3, 21); v3 ^= v0;
    v2 += v1; v1 = rotl(v1, 17); v1 ^= v2; v2 = rotl(v2, 32);
  };
  size_t i = 0;
  for (; i + 8 <= len; i += 8) {
    uint64_t m = 0;
    for (int b = 0; b < 8; b++) m |= (uint64_t)msg[i + b] << (8 * b);
    v3 ^= m; round(); v0 ^= m;
  }
  uint64_t last = (uint64_t)(len & 0xff) << 56;
  for (int b = 0; i + b < len; b++) last |= (uint64_t)msg[i + b] << (8 * b);
  v3 ^= last; round(); v0 ^= last;
  v2 ^= 0xff;
  round(); round(); round();
  return v0 ^ v1 ^ v2 ^ v3;
}

// ChaCha8Rng::from_seed(seed32): key = the seed, block counter 0, stream 0
inline ChaCha8 chacha8_from_seed(const uint8_t seed[32]) {
  ChaCha8 r(0);
  for (int i = 0; i < 8; i++)
    r.key[i] = (uint32_t)seed[4 * i] | (uint32_t)seed[4 * i + 1] << 8 | (uint32_t)seed[4 * i + 2] << 16 | (uint32_t)seed[4 * i + 3] << 24;
  r.counter = 0;
  r.pos = 16;
  return r;
}

// rand 0.8.5 Uniform::<u64>::from(0..q).sample(rng)
inline uint64_t uniform_u64_sample(ChaCha8& rng, uint64_t q) {
  const uint64_t ints_to_reject = (UINT64_MAX - q + 1) % q, zone = UINT64_MAX - ints_to_reject;
  for (;;) {
    const u128 m = (u128)rng.next_u64() * q;
    if ((uint64_t)m <= zone) return (uint64_t)(m >> 64);
  }
}

// fhe-math Poly::random_from_seed(ctx, Ntt, seed) -> out u64[L][ell]
inline void poly_random_from_seed(const uint8_t seed[32], const uint64_t* moduli, uint32_t L, uint32_t ell, uint64_t* out) {
  uint8_t digest[32];
  sha256(seed, 32, digest);
  ChaCha8 prng = chacha8_from_seed(digest);
  for (uint32_t j = 0; j < L; j++)
    for (uint32_t c = 0; c < ell; c++) out[(size_t)j * ell + c] = uniform_u64_sample(prng, moduli[j]);
}

// PvwCrs::new_deterministic (crs.rs:45-67): A u64[k][k][L][ell], elements in row-major (ndarray iter_mut) order
inline void crs_new_deterministic(const uint8_t seed[32], uint32_t k, const uint64_t* moduli, uint32_t L, uint32_t ell, uint64_t* A) {
  ChaCha8 master = chacha8_from_seed(seed);
  const size_t poly = (size_t)L * ell;
  for (size_t e = 0; e < (size_t)k * k; e++) {
    uint8_t element_seed[32];
    for (int b = 0; b < 32; b++) element_seed[b] = (uint8_t)master.next_u32();  // gen::<[u8; 32]>(): one u32 per byte
    poly_random_from_seed(element_seed, moduli, L, ell, A + e * poly);
  }
}

// new_from_tag (crs.rs:74-90): DefaultHasher over (tag + "CRS") as a str (bytes, then 0xff), u64 cycled to 32 bytes
inline void crs_tag_to_seed(const std::string& tag, uint8_t seed[32]) {
  std::string s = tag + "CRS";
  std::vector<uint8_t> bytes(s.begin(), s.end());
  bytes.push_back(0xff);
  const uint64_t h = siphash13_zero_key(bytes.data(), bytes.size());
  for (int i = 0; i < 32; i++) seed[i] = (uint8_t)(h >> (8 * (i % 8)));
}

}  // namespace pvw
