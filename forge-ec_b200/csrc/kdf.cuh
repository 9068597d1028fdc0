// kdf.cuh -- what follows k*P in a batched ECDH: SHA-256, HMAC-SHA-256 and the reference's key derivations, per lane.
//
//   KeyExchange::derive_shared_secret  = x-coordinate of k*P, 32 bytes big-endian, identity is an error
//                                        (forge-ec-curves/src/secp256k1.rs:1885-1905, p256.rs:2281-2302)
//   Secp256k1::derive_key              = HKDF-SHA-256 (RFC 5869) with a salt of 32 zero bytes (secp256k1.rs:1846-1883;
//                                        the reference uses the `hmac`/`sha2` crates, restated here from FIPS 180-4 / RFC 2104)
//   P256::derive_key                   = a placeholder that XORs secret and info into a zero buffer (p256.rs:2314-2344)
//
// Host-compilable (tests/hostsim).  All loops have public trip counts (lengths are call parameters, not secrets).
#pragma once
#include "u256.cuh"

namespace fec {

enum { KDF_NONE = 0, KDF_HKDF_SHA256 = 1, KDF_XOR_INFO = 2 };
constexpr int KDF_MAX_INFO = 256;
constexpr int KDF_MAX_OUT = 64;

FEC_DEV u32 rotr32(u32 x, int n) { return (x >> n) | (x << (32 - n)); }
FEC_DEV u32 sha256_k(int i) {
  const u32 K[64] = {
      0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u,
      0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu,
      0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u,
      0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
      0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u,
      0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u,
      0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
  return K[i];
}
FEC_DEV void sha256_init(u32 h[8]) {
  h[0] = 0x6a09e667u; h[1] = 0xbb67ae85u; h[2] = 0x3c6ef372u; h[3] = 0xa54ff53au;
  h[4] = 0x510e527fu; h[5] = 0x9b05688cu; h[6] = 0x1f83d9abu; h[7] = 0x5be0cd19u;
}
// one FIPS 180-4 compression: w = the 16 big-endian words of the block (destroyed)
FEC_DEV void sha256_compress(u32 h[8], u32 w[16]) {
  u32 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll 1
  for (int r = 0; r < 64; r += 16) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (r > 0) {
        const u32 w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
        const u32 s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
        const u32 s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
        w[i] = w[i] + s0 + w[(i + 9) & 15] + s1;
      }
      const u32 S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
      const u32 ch = (e & f) ^ (~e & g);
      const u32 t1 = hh + S1 + ch + sha256_k(r + i) + w[i];
      const u32 S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
      const u32 mj = (a & b) ^ (a & c) ^ (b & c);
      const u32 t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
// streaming SHA-256 over bytes (messages here are at most 64 + 32 + 256 + 1 bytes)
struct Sha256 {
  u32 h[8];
  u32 w[16];
  u32 fill;   // bytes in the current block
  u32 total;  // bytes so far
};
FEC_DEV void sha256_begin(Sha256& s) {
  sha256_init(s.h);
#pragma unroll
  for (int i = 0; i < 16; i++) s.w[i] = 0u;
  s.fill = 0u;
  s.total = 0u;
}
FEC_DEV void sha256_byte(Sha256& s, u32 v) {
  s.w[s.fill >> 2] |= (v & 0xFFu) << (24 - 8 * (s.fill & 3u));
  s.fill++;
  s.total++;
  if (s.fill == 64u) {
    sha256_compress(s.h, s.w);
#pragma unroll
    for (int i = 0; i < 16; i++) s.w[i] = 0u;
    s.fill = 0u;
  }
}
FEC_DEV void sha256_words(Sha256& s, const u32* be_words, int nwords) {  // only on a word boundary
  for (int i = 0; i < nwords; i++) {
    s.w[s.fill >> 2] = be_words[i];
    s.fill += 4u;
    s.total += 4u;
    if (s.fill == 64u) {
      sha256_compress(s.h, s.w);
#pragma unroll
      for (int j = 0; j < 16; j++) s.w[j] = 0u;
      s.fill = 0u;
    }
  }
}
FEC_DEV void sha256_end(Sha256& s, u32 digest[8]) {
  const u32 bits = s.total * 8u;
  sha256_byte(s, 0x80u);
  if (s.fill > 56u) {
    sha256_compress(s.h, s.w);
#pragma unroll
    for (int i = 0; i < 16; i++) s.w[i] = 0u;
  }
  s.w[15] = bits;  // messages are far below 2^32 bits
  sha256_compress(s.h, s.w);
#pragma unroll
  for (int i = 0; i < 8; i++) digest[i] = s.h[i];
}
// HMAC-SHA-256 with a 32-byte key (8 big-endian words): begin -> feed message -> end
FEC_DEV void hmac256_begin(Sha256& s, const u32 key[8]) {
  sha256_begin(s);
  u32 pad[16];
#pragma unroll
  for (int i = 0; i < 16; i++) pad[i] = (i < 8 ? key[i] : 0u) ^ 0x36363636u;
  sha256_words(s, pad, 16);
}
FEC_DEV void hmac256_end(Sha256& s, const u32 key[8], u32 mac[8]) {
  u32 inner[8], pad[16];
  sha256_end(s, inner);
  sha256_begin(s);
#pragma unroll
  for (int i = 0; i < 16; i++) pad[i] = (i < 8 ? key[i] : 0u) ^ 0x5c5c5c5cu;
  sha256_words(s, pad, 16);
  sha256_words(s, inner, 8);
  sha256_end(s, mac);
}
// secp256k1.rs:1846-1883: PRK = HMAC(0^32, secret); T(i) = HMAC(PRK, T(i-1) || info || i); OKM = T(1) || T(2) ...
// secret: 8 big-endian words (the x-coordinate as FieldElement::to_bytes writes it); okm: out_len <= KDF_MAX_OUT bytes
FEC_DEV void hkdf_sha256_x(uint8_t* okm, int out_len, const u32 secret_be[8], const uint8_t* info, int info_len) {
  Sha256 s;
  u32 zero[8] = {0, 0, 0, 0, 0, 0, 0, 0}, prk[8], t[8];
  hmac256_begin(s, zero);
  sha256_words(s, secret_be, 8);
  hmac256_end(s, zero, prk);
  int done = 0;
  for (u32 counter = 1; done < out_len; counter++) {
    hmac256_begin(s, prk);
    if (counter > 1) sha256_words(s, t, 8);
    for (int i = 0; i < info_len; i++) sha256_byte(s, info[i]);
    sha256_byte(s, counter);
    hmac256_end(s, prk, t);
    for (int i = 0; i < 32 && done < out_len; i++, done++) okm[done] = (uint8_t)(t[i >> 2] >> (24 - 8 * (i & 3)));
  }
}
// p256.rs:2314-2344: okm = zeros; okm[i] ^= secret[i]; okm[i] ^= info[i]   (each for i < out_len)
FEC_DEV void kdf_xor_info(uint8_t* okm, int out_len, const u32 secret_be[8], const uint8_t* info, int info_len) {
  for (int i = 0; i < out_len; i++) {
    u32 v = 0u;
    if (i < 32) v ^= (secret_be[i >> 2] >> (24 - 8 * (i & 3))) & 0xFFu;
    if (i < info_len) v ^= info[i];
    okm[i] = (uint8_t)v;
  }
}

}  // namespace fec
