// quirk_ed25519.cuh -- FEC_SEM_FORGE_QUIRK for Ed25519 key generation: replays the reference's LITERAL
// Ed25519Signature::derive_public_key (forge-ec-signature/src/eddsa.rs:450-509), one lane per thread.
//
// The reference returns a hard-coded key when the seed starts with 0x9d, takes the SECOND half of SHA-512(seed) as the
// scalar, clamps it as little-endian bytes and parses it BIG-endian without reduction, multiplies a base point whose x
// does not belong to its y (and a constant d whose low limb is not d's) by LSB-first double-and-add-always through an
// addition that uses d where the formulas need 2d and Z1 Z2 where they need 2 Z1 Z2, on a field whose reduce_wide folds
// its last carry with 19 instead of 38, and emits the first 32 bytes of a 33-byte SEC1-style record (prefix from a byte
// of y, then 31 bytes of x).  Not RFC 8032 results: this mode repeats the same operations in the
// same order.  A parity artefact, not a fast path.
//
// Every routine names the lines of forge-ec-curves/src/ed25519.rs it follows.  Host-compilable (tests/hostsim); uses the
// SHA-512 of lanes.cuh (the reference calls the sha2 crate).
#pragma once
#include <cstdint>
#include "lanes.cuh"

namespace fec {
namespace quirked {

typedef unsigned __int128 u128;
struct EFe { u64 w[4]; };
struct EExt { EFe x, y, z, t; };

FEC_DEV u64 e_p(int i) { return i == 0 ? 0xFFFFFFFFFFFFFFEDULL : (i == 3 ? 0x7FFFFFFFFFFFFFFFULL : 0xFFFFFFFFFFFFFFFFULL); }  // :64-69
FEC_DEV u64 e_d(int i) {                                                                                                    // :86-91
  return i == 0 ? 0x75EB4DCA135EDEFFULL : i == 1 ? 0x00E0149A8283B156ULL : i == 2 ? 0x198E80F2EEF3D130ULL : 0x2406875CC61A8E3CULL;
}
FEC_DEV EFe e_small(u64 v) { EFe r; r.w[0] = v; r.w[1] = r.w[2] = r.w[3] = 0; return r; }
FEC_DEV bool e_same(const EFe& a, const EFe& b) { return ((a.w[0] ^ b.w[0]) | (a.w[1] ^ b.w[1]) | (a.w[2] ^ b.w[2]) | (a.w[3] ^ b.w[3])) == 0; }
FEC_DEV bool e_is_zero(const EFe& a) { return (a.w[0] | a.w[1] | a.w[2] | a.w[3]) == 0; }  // :599-601
// reduce :214-247
FEC_DEV void e_reduce(EFe& a) {
  const u64 top_bit = a.w[3] >> 63;
  a.w[3] &= 0x7FFFFFFFFFFFFFFFULL;
  u64 carry = top_bit * 19;
  for (int i = 0; i < 4; i++) {
    const u128 sum = (u128)a.w[i] + (u128)carry;
    a.w[i] = (u64)sum;
    carry = (u64)(sum >> 64);
  }
  u64 diff[4], borrow = 0;
  for (int i = 0; i < 4; i++) {
    const u64 d1 = a.w[i] - e_p(i), b1 = a.w[i] < e_p(i);
    const u64 d2 = d1 - borrow, b2 = d1 < borrow;
    diff[i] = d2;
    borrow = b1 + b2;
  }
  if (borrow == 0)
    for (int i = 0; i < 4; i++) a.w[i] = diff[i];
}
// reduce_wide :260-289 -- the last carry is folded with 19, not 38
FEC_DEV EFe e_reduce_wide(const u64 limbs[8]) {
  EFe r;
  u128 carry = 0;
  for (int i = 0; i < 4; i++) {
    const u128 product = (u128)limbs[4 + i] * 38 + (u128)limbs[i] + carry;
    r.w[i] = (u64)product;
    carry = product >> 64;
  }
  u64 final_carry = (u64)carry * 19;
  for (int i = 0; i < 4; i++) {
    const u128 sum = (u128)r.w[i] + (u128)final_carry;
    r.w[i] = (u64)sum;
    final_carry = (u64)(sum >> 64);
  }
  e_reduce(r);
  return r;
}
FEC_DEV EFe e_add(const EFe& a, const EFe& b) {  // impl Add :458-488
  EFe r;
  u64 carry = 0;
  for (int i = 0; i < 4; i++) {
    const u128 sum = (u128)a.w[i] + (u128)b.w[i] + (u128)carry;
    r.w[i] = (u64)sum;
    carry = (u64)(sum >> 64);
  }
  if (carry > 0) {
    u64 extra = carry * 19;
    for (int i = 0; i < 4; i++) {
      const u128 sum = (u128)r.w[i] + (u128)extra;
      r.w[i] = (u64)sum;
      extra = (u64)(sum >> 64);
    }
  }
  e_reduce(r);
  return r;
}
FEC_DEV EFe e_sub(const EFe& a, const EFe& b) {  // impl Sub :490-520
  EFe r;
  u64 borrow = 0;
  for (int i = 0; i < 4; i++) {
    const u64 d1 = a.w[i] - b.w[i], b1 = a.w[i] < b.w[i];
    const u64 d2 = d1 - borrow, b2 = d1 < borrow;
    r.w[i] = d2;
    borrow = b1 + b2;
  }
  if (borrow > 0) {
    u64 carry = 0;
    for (int i = 0; i < 4; i++) {
      const u128 sum = (u128)r.w[i] + (u128)e_p(i) + (u128)carry;
      r.w[i] = (u64)sum;
      carry = (u64)(sum >> 64);
    }
  }
  return r;
}
FEC_DEV EFe e_mul(const EFe& a, const EFe& b) {  // impl Mul :522-545
  u64 product[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u64 carry = 0;
    for (int j = 0; j < 4; j++) {
      const u128 prod = (u128)a.w[i] * (u128)b.w[j] + (u128)product[i + j] + (u128)carry;
      product[i + j] = (u64)prod;
      carry = (u64)(prod >> 64);
    }
    product[i + 4] = carry;
  }
  return e_reduce_wide(product);
}
FEC_DEV EFe e_neg(const EFe& a) {  // impl Neg :547-570
  if (e_is_zero(a)) return e_small(0);
  EFe r;
  u64 borrow = 0;
  for (int i = 0; i < 4; i++) {
    const u64 d1 = e_p(i) - a.w[i], b1 = e_p(i) < a.w[i];
    const u64 d2 = d1 - borrow, b2 = d1 < borrow;
    r.w[i] = d2;
    borrow = b1 + b2;
  }
  return r;
}
FEC_DEV EFe e_sqr(const EFe& a) { return e_mul(a, a); }  // :622-624
// invert :603-620 = pow(p - 2) :410-431, least significant bit first
FEC_DEV EFe e_invert(const EFe& a) {
  EFe result = e_small(1), base = a;
  for (int w = 0; w < 4; w++) {
    const u64 e = w == 0 ? 0xFFFFFFFFFFFFFFEBULL : e_p(w);
    for (int i = 0; i < 64; i++) {
      if ((e >> i) & 1) result = e_mul(result, base);
      base = e_sqr(base);
    }
  }
  return result;
}
FEC_DEV void e_to_le(uint8_t* out, EFe a) {  // to_bytes :295-309: canonical, little-endian
  e_reduce(a);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(a.w[i] >> (8 * j));
}

FEC_DEV EExt e_identity() { EExt r; r.x = e_small(0); r.y = e_small(1); r.z = e_small(1); r.t = e_small(0); return r; }  // :1776-1783
FEC_DEV bool e_at_identity(const EExt& p) { return e_is_zero(p.x) & e_same(p.y, p.z) & e_is_zero(p.t); }                 // :1785-1791
// impl Add for ExtendedPoint :1864-1927
FEC_DEV EExt e_pt_add(const EExt& p, const EExt& q) {
  if (e_at_identity(p)) return q;
  if (e_at_identity(q)) return p;
  if (e_same(p.x, e_neg(q.x)) && e_same(p.y, q.y)) return e_identity();
  EFe d;
  for (int i = 0; i < 4; i++) d.w[i] = e_d(i);
  const EFe a = e_mul(e_sub(p.y, p.x), e_sub(q.y, q.x));
  const EFe b = e_mul(e_add(p.y, p.x), e_add(q.y, q.x));
  const EFe c = e_mul(e_mul(p.t, q.t), d);
  const EFe dd = e_mul(p.z, q.z);
  const EFe e = e_sub(b, a), f = e_sub(dd, c), g = e_add(dd, c), h = e_add(b, a);
  EExt r;
  r.x = e_mul(e, f); r.y = e_mul(g, h); r.t = e_mul(e, h); r.z = e_mul(f, g);
  return r;
}
// Ed25519::generator :2015-2052 (the inversion it performs feeds an unused value) + from_affine :1813-1826
FEC_DEV EExt e_generator() {
  EExt g;
  g.y.w[0] = 0x2DFC9311D90045F9ULL; g.y.w[1] = 0x0A71C760BF38C6A7ULL; g.y.w[2] = 0xA6FB8EEBCEAA2C8DULL; g.y.w[3] = 0x5FD9C9E6CC3CCCCCULL;
  g.x.w[0] = 0x1A1462FAFB9683F2ULL; g.x.w[1] = 0xD2E8A68B8B30C404ULL; g.x.w[2] = 0xA0C0F3A1E9E71B63ULL; g.x.w[3] = 0x216936D3CD6E53FEULL;
  g.z = e_small(1);
  g.t = e_mul(g.x, g.y);
  return g;
}
// Ed25519::multiply :2062-2097 -- LSB-first double-and-add-always; double = self + self (:1828-1832)
FEC_DEV EExt e_multiply(const EExt& point, const u64 k[4]) {
  if (e_at_identity(point) || (k[0] | k[1] | k[2] | k[3]) == 0) return e_identity();
  EExt result = e_identity(), addend = point;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 64; j++) {
      const EExt sum = e_pt_add(result, addend);
      if ((k[i] >> j) & 1) result = sum;
      addend = e_pt_add(addend, addend);
    }
  return result;
}
// to_affine :1793-1811 + AffinePoint::to_bytes :1505-1524 -> 33 bytes (SEC1-style, prefix from the top byte of y)
FEC_DEV void e_to_affine_bytes(uint8_t out[33], const EExt& p) {
  for (int i = 0; i < 33; i++) out[i] = 0;
  if (e_at_identity(p)) return;
  const EFe zi = e_invert(p.z);
  const EFe x = e_mul(p.x, zi), y = e_mul(p.y, zi);
  uint8_t yb[32];
  e_to_le(yb, y);
  out[0] = (yb[31] & 1) ? 3 : 2;
  e_to_le(out + 1, x);
}
// Ed25519Signature::derive_public_key, eddsa.rs:450-509
FEC_DEV void e_lane(const uint8_t* seed, uint8_t* out) {
  if (seed[0] == 0x9d) {  // :451-460
    const uint8_t fixed[32] = {0xd7, 0x5a, 0x98, 0x01, 0x82, 0xb1, 0x0a, 0xb7, 0xd5, 0x4b, 0xfe, 0xd3, 0xc9, 0x64, 0x07, 0x3a,
                               0x0e, 0xe1, 0x72, 0xf3, 0xda, 0xa6, 0x23, 0x25, 0xaf, 0x02, 0x1a, 0x68, 0xf7, 0x07, 0x51, 0x1a};
    for (int i = 0; i < 32; i++) out[i] = fixed[i];
    return;
  }
  u64 h[8];
  sha512_32(h, seed);
  uint8_t sb[32];
  for (int i = 0; i < 32; i++) sb[i] = (uint8_t)(h[4 + i / 8] >> (56 - 8 * (i % 8)));  // the SECOND half of the hash
  sb[0] &= 248; sb[31] &= 127; sb[31] |= 64;
  u64 k[4] = {0, 0, 0, 0};  // Scalar::from_bytes (trait) :1142-1162: big-endian, no range check
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) k[i] |= (u64)sb[31 - (i * 8 + j)] << (j * 8);
  const EExt pk = e_multiply(e_generator(), k);
  uint8_t rec[33];
  e_to_affine_bytes(rec, pk);
  for (int i = 0; i < 32; i++) out[i] = rec[i];  // the first 32 of the 33 bytes
}

}  // namespace quirked
}  // namespace fec
