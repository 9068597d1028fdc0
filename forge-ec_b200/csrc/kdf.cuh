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
// one FIPS 180-4 compression: w = the 16 big-endian words of the block (destroyed).  Fully unrolled: every index is a
// compile-time constant, so the message schedule and the state live in registers and the round constants are immediates.
FEC_DEV void sha256_compress(u32 h[8], u32 w[16]) {
  u32 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
  for (int i = 0; i < 64; i++) {
    if (i >= 16) {
      const u32 w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      const u32 s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
      const u32 s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
      w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
    }
    const u32 S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    const u32 ch = (e & f) ^ (~e & g);
    const u32 t1 = hh + S1 + ch + sha256_k(i) + w[i & 15];
    const u32 S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    const u32 mj = (a & b) ^ (a & c) ^ (b & c);
    const u32 t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// ---- HKDF-SHA-256 with the all-zero salt (secp256k1.rs:1846-1883):
//        PRK = HMAC(0^32, secret);  T(i) = HMAC(PRK, T(i-1) || info || i);  OKM = T(1) || T(2) ...
// Everything that does not depend on the lane is prepared once per call on the host (KdfPlan, passed as a kernel
// parameter): the padded message blocks `info || counter || 0x80 0.. || length` of T(1) and T(2) (out_len <= 64, so
// there are at most two), with eight placeholder words for T(1) at the head of T(2)'s message.  The states after the
// ipad / opad blocks of the zero key are constants.  A lane then runs 6 compressions for 32 bytes of output and a short
// info (2 for PRK, 2 for the key pads of PRK, 1 + 1 for T(1)), all in registers.
constexpr int KDF_MAX_BLOCKS = 5;  // 32 + 256 + 1 + 9 bytes padded to 64
struct KdfPlan {
  u32 msg1[KDF_MAX_BLOCKS * 16];  // T(1): info || 0x01, padded
  u32 msg2[KDF_MAX_BLOCKS * 16];  // T(2): T(1) (words 0..7, filled in per lane) || info || 0x02, padded
  int blocks1, blocks2;
  int out_len;
};
// host side (also tests/hostsim): big-endian words of prefix_zeros || info || counter, SHA-256 padding for a message that
// follows one 64-byte key-pad block
inline int kdf_plan_message(u32* words, int prefix_bytes, const uint8_t* info, int info_len, int counter) {
  const int len = prefix_bytes + info_len + 1;
  const int blocks = (len + 9 + 63) / 64;
  for (int i = 0; i < blocks * 16; i++) words[i] = 0u;
  auto put = [&](int pos, u32 v) { words[pos >> 2] |= (v & 0xFFu) << (24 - 8 * (pos & 3)); };
  for (int i = 0; i < info_len; i++) put(prefix_bytes + i, info[i]);
  put(prefix_bytes + info_len, (u32)counter);
  put(len, 0x80u);
  words[blocks * 16 - 1] = (u32)(64 + len) * 8u;
  return blocks;
}
inline void kdf_plan_build(KdfPlan& p, int kdf, const uint8_t* info, int info_len, int out_len) {
  p.out_len = out_len;
  if (kdf == KDF_HKDF_SHA256) {
    p.blocks1 = kdf_plan_message(p.msg1, 0, info, info_len, 1);
    p.blocks2 = kdf_plan_message(p.msg2, 32, info, info_len, 2);
  } else {  // KDF_XOR_INFO: the first 64 info bytes as big-endian words, zero-padded
    p.blocks1 = p.blocks2 = 0;
    for (int i = 0; i < 16; i++) p.msg1[i] = 0u;
    for (int i = 0; i < info_len && i < 64; i++) p.msg1[i >> 2] |= (u32)info[i] << (24 - 8 * (i & 3));
  }
}
// HMAC tail shared by every use: outer = compress(outer_state, inner_digest || 0x80 0.. || (64 + 32) * 8)
FEC_DEV void hmac256_finish(u32 mac[8], const u32 outer_state[8], const u32 inner_digest[8]) {
  u32 w[16];
#pragma unroll
  for (int i = 0; i < 8; i++) { w[i] = inner_digest[i]; mac[i] = outer_state[i]; }
  w[8] = 0x80000000u;
#pragma unroll
  for (int i = 9; i < 15; i++) w[i] = 0u;
  w[15] = 768u;
  sha256_compress(mac, w);
}
// okm: out_len bytes as big-endian words okm_be[0 .. (out_len + 3) / 4)
FEC_DEV void hkdf_sha256_plan(u32 okm_be[16], const u32 secret_be[8], const KdfPlan& plan) {
  const u32 IPAD0[8] = {0xf454deadu, 0x9725214fu, 0x90daf2a0u, 0xdf1228eau, 0x64e5750fu, 0xa3924181u, 0x824a932bu, 0xf8e04e32u};
  const u32 OPAD0[8] = {0xd385480fu, 0x7abb6477u, 0x37c9c538u, 0x5dd82467u, 0x8e043a72u, 0x753434b0u, 0xdeb82818u, 0x361d45a6u};
  u32 w[16], st[8], prk[8], kin[8], kout[8];
  // PRK = HMAC(0^32, secret)
#pragma unroll
  for (int i = 0; i < 8; i++) { w[i] = secret_be[i]; st[i] = IPAD0[i]; }
  w[8] = 0x80000000u;
#pragma unroll
  for (int i = 9; i < 15; i++) w[i] = 0u;
  w[15] = 768u;
  sha256_compress(st, w);
  hmac256_finish(prk, OPAD0, st);
  // the two key-pad states of PRK
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = (i < 8 ? prk[i] : 0u) ^ 0x36363636u;
  sha256_init(kin);
  sha256_compress(kin, w);
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = (i < 8 ? prk[i] : 0u) ^ 0x5c5c5c5cu;
  sha256_init(kout);
  sha256_compress(kout, w);
  // T(1)
#pragma unroll
  for (int i = 0; i < 8; i++) st[i] = kin[i];
#pragma unroll 1
  for (int b = 0; b < plan.blocks1; b++) {
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = plan.msg1[16 * b + i];
    sha256_compress(st, w);
  }
  hmac256_finish(okm_be, kout, st);
  if (plan.out_len <= 32) return;  // uniform
  // T(2)
#pragma unroll
  for (int i = 0; i < 8; i++) st[i] = kin[i];
#pragma unroll 1
  for (int b = 0; b < plan.blocks2; b++) {
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = plan.msg2[16 * b + i];
    if (b == 0) {
#pragma unroll
      for (int i = 0; i < 8; i++) w[i] = okm_be[i];
    }
    sha256_compress(st, w);
  }
  hmac256_finish(okm_be + 8, kout, st);
}
// p256.rs:2314-2344: okm = zeros; okm[i] ^= secret[i]; okm[i] ^= info[i]   (each for i < out_len).  info_be: the info
// bytes as big-endian words, zero-padded to 64 bytes (only the first out_len <= 64 bytes of info take part)
FEC_DEV void kdf_xor_plan(u32 okm_be[16], const u32 secret_be[8], const u32 info_be[16]) {
#pragma unroll
  for (int i = 0; i < 16; i++) okm_be[i] = (i < 8 ? secret_be[i] : 0u) ^ info_be[i];
}

// one lane: the derived key as big-endian words (the first plan.out_len bytes are the output)
FEC_DEV void kdf_lane(u32 okm_be[16], int kdf, const u32 secret_be[8], const KdfPlan& plan) {
  if (kdf == KDF_HKDF_SHA256) {
    hkdf_sha256_plan(okm_be, secret_be, plan);
  } else {
    kdf_xor_plan(okm_be, secret_be, plan.msg1);
  }
}

}  // namespace fec
