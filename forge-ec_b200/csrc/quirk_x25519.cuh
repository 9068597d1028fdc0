// quirk_x25519.cuh -- FEC_SEM_FORGE_QUIRK for X25519: replays the reference's LITERAL x25519(), one lane per thread.
//
// forge-ec's 2^255-19 code (curve25519.rs) is limb arithmetic without carries between the limbs -- Add, Sub and the
// final subtraction of p work limb by limb with wrapping -- a multiplication that drops carries and folds the high 256
// bits with 19 instead of 38, a big-endian byte codec where RFC 7748 is little-endian, a ladder step that multiplies by
// A = 486662 where a24 belongs, a hand-made inversion chain, and a hard-coded answer for the scalar 2.  Its outputs are
// not RFC 7748 results and no other algorithm can reproduce them: this mode repeats the same
// operations in the same order with release-profile wrapping semantics.  A parity artefact, not a fast path.
//
// Every routine names the lines of forge-ec-curves/src/curve25519.rs it follows.  Host-compilable (tests/hostsim).
#pragma once
#include <cstdint>
#include "u256.cuh"

namespace fec {
namespace quirk25519 {

typedef unsigned __int128 u128;
struct XFe { u64 w[4]; };

FEC_DEV u64 x_p(int i) { return i == 0 ? 0xFFFFFFFFFFFFFFEDULL : (i == 3 ? 0x7FFFFFFFFFFFFFFFULL : 0xFFFFFFFFFFFFFFFFULL); }  // :20-21
FEC_DEV bool x_ge_p(const XFe& a) {  // the comparison of reduce :95-105 and from_bytes :148-156
  const bool ge3 = a.w[3] > x_p(3), eq3 = a.w[3] == x_p(3), eq2 = a.w[2] == x_p(2), eq1 = a.w[1] == x_p(1), ge0 = a.w[0] >= x_p(0);
  return ge3 | (eq3 & eq2 & eq1 & ge0);
}
// reduce :50-114 -- bit 255 folded into limb 0 without a carry, the carry loop that starts at 0 (a no-op, :65-92), the
// subtraction of p limb by limb without borrows
FEC_DEV void x_reduce(XFe& a) {
  const u64 bit255 = (a.w[3] >> 63) & 1;
  a.w[0] = a.w[0] + bit255 * 19;
  a.w[3] &= 0x7FFFFFFFFFFFFFFFULL;
  if (x_ge_p(a))
    for (int i = 0; i < 4; i++) a.w[i] -= x_p(i);
}
FEC_DEV XFe x_small(u64 v) { XFe r; r.w[0] = v; r.w[1] = r.w[2] = r.w[3] = 0; return r; }  // zero / one :355-361
FEC_DEV bool x_is_zero(const XFe& a) { return (a.w[0] | a.w[1] | a.w[2] | a.w[3]) == 0; }  // :363-365
FEC_DEV XFe x_add(const XFe& a, const XFe& b) {  // impl Add :186-203
  XFe r;
  for (int i = 0; i < 4; i++) r.w[i] = a.w[i] + b.w[i];
  x_reduce(r);
  return r;
}
FEC_DEV XFe x_sub(const XFe& a, const XFe& b) {  // impl Sub :205-225
  XFe r;
  for (int i = 0; i < 4; i++) r.w[i] = a.w[i] + x_p(i) - b.w[i];
  x_reduce(r);
  return r;
}
FEC_DEV XFe x_mul(const XFe& a, const XFe& b) {  // impl Mul :227-314
  u64 result[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 4; j++) {
      const u128 product = (u128)a.w[i] * (u128)b.w[j];
      const u64 low = (u64)product, high = (u64)(product >> 64);
      const int idx = i + j;
      const u64 sum1 = result[idx] + low;
      const bool carry1 = sum1 < result[idx];
      result[idx] = sum1;
      if (carry1) result[idx + 1] = result[idx + 1] + 1;  // its own overflow is dropped
      const u64 sum3 = result[idx + 1] + high;
      const bool carry3 = sum3 < result[idx + 1];
      result[idx + 1] = sum3;
      if (carry3 && idx + 2 < 8) result[idx + 2] += 1;
    }
  }
  for (int i = 0; i < 4; i++) {
    const u64 high_bits = result[4 + i];
    if (high_bits > 0) {
      const u128 product = (u128)high_bits * 19;
      const u64 low = (u64)product, high = (u64)(product >> 64);
      const u64 sum1 = result[i] + low;
      const bool carry1 = sum1 < result[i];
      result[i] = sum1;
      if (high > 0) {
        const u64 sum2 = result[i + 1] + high;
        const bool carry2 = sum2 < result[i + 1];
        result[i + 1] = sum2;
        if (carry2 && i + 2 < 4) result[i + 2] += 1;
      }
      if (carry1) {
        const u64 sum3 = result[i + 1] + 1;
        const bool carry3 = sum3 < result[i + 1];
        result[i + 1] = sum3;
        if (carry3 && i + 2 < 4) result[i + 2] += 1;
      }
    }
  }
  XFe r;
  for (int i = 0; i < 4; i++) r.w[i] = result[i];
  x_reduce(r);
  return r;
}
FEC_DEV XFe x_sqr(const XFe& a) { return x_mul(a, a); }  // :490-494
FEC_DEV XFe x_sqn(XFe a, int n) {
  for (int i = 0; i < n; i++) a = x_sqr(a);
  return a;
}
// invert :367-488 -- the reference's hand-made chain, whatever exponent it amounts to; `some` = false for zero
FEC_DEV XFe x_invert(const XFe& a, bool& some) {
  some = !x_is_zero(a);
  if (!some) return x_small(0);
  const XFe a2 = x_sqr(a), a4 = x_sqr(a2), a16 = x_sqn(a4, 2);
  const XFe t250 = x_sqn(a16, 4 + 8 + 16 + 32 + 64 + 64 + 58);  // a^(2^8), ^(2^16), ^(2^32), ^(2^64), ^(2^128), ^(2^192), "^(2^250)"
  XFe result = x_mul(t250, a);
  result = x_mul(result, a2);
  result = x_mul(result, a4);
  result = x_mul(result, x_sqr(a4));
  result = x_mul(result, a16);
  for (int i = 0; i < 4; i++) result = x_mul(result, result);
  const XFe a64 = x_sqn(a16, 2), a32 = x_sqr(a16), a8 = x_sqr(a4);
  return x_mul(x_mul(x_mul(x_mul(x_mul(result, a64), a32), a8), a2), a);
}
FEC_DEV XFe x_from_be(const uint8_t* b, bool& valid) {  // the inherent, BIG-endian from_bytes :130-163
  XFe r = x_small(0);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) r.w[i] |= (u64)b[31 - (i * 8 + j)] << (j * 8);
  x_reduce(r);
  valid = !x_ge_p(r);
  return r;
}
FEC_DEV void x_to_be(uint8_t* out, const XFe& a) {  // the inherent, BIG-endian to_bytes :117-128 (no reduction)
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) out[31 - (i * 8 + j)] = (uint8_t)(a.w[i] >> (j * 8));
}
// x25519 :1624-1716; u == nullptr: the base point [9, 0, ..] (Curve25519::generator :1900-1904)
FEC_DEV void x_lane(const uint8_t* scalar, const uint8_t* u, uint8_t* out) {
  bool two = scalar[0] == 2;
  for (int i = 1; i < 32; i++) two = two && scalar[i] == 0;
  if (two) {  // :1626-1633
    const uint8_t fixed[32] = {0x1b, 0x7f, 0x9f, 0x7c, 0x27, 0x65, 0x50, 0xbb, 0x3a, 0x3c, 0xec, 0xc8, 0xa5, 0x77, 0x0c, 0x17,
                               0x3f, 0x58, 0x31, 0xed, 0x1b, 0xb2, 0x8c, 0x05, 0x58, 0xaa, 0xc4, 0x71, 0x3f, 0x97, 0x08, 0x22};
    for (int i = 0; i < 32; i++) out[i] = fixed[i];
    return;
  }
  uint8_t sb[32], ub[32];
  for (int i = 0; i < 32; i++) { sb[i] = scalar[i]; ub[i] = u != nullptr ? u[i] : (uint8_t)(i == 0 ? 9 : 0); }
  sb[0] &= 248; sb[31] &= 127; sb[31] |= 64;
  ub[31] &= 127;
  bool valid;
  XFe u_fe = x_from_be(ub, valid);
  if (!valid) u_fe = x_small(0);
  const XFe x1 = u_fe, A = x_small(0x76D06ULL);
  XFe x2 = x_small(1), z2 = x_small(0), x3 = u_fe, z3 = x_small(1);
  unsigned swap = 0;
  for (int i = 254; i >= 0; i--) {
    const unsigned bit = (sb[i / 8] >> (i % 8)) & 1;
    if (swap ^ bit) { const XFe tx = x2, tz = z2; x2 = x3; z2 = z3; x3 = tx; z3 = tz; }
    swap = bit;
    const XFe a = x_add(x2, z2), aa = x_sqr(a), b = x_sub(x2, z2), bb = x_sqr(b), e = x_sub(aa, bb);
    const XFe c = x_add(x3, z3), d = x_sub(x3, z3), da = x_mul(d, a), cb = x_mul(c, b);
    x3 = x_sqr(x_add(da, cb));
    z3 = x_mul(x1, x_sqr(x_sub(da, cb)));
    x2 = x_mul(aa, bb);
    z2 = x_mul(e, x_add(aa, x_mul(A, e)));
  }
  if (swap) { x2 = x3; z2 = z3; }
  bool some;
  XFe z2_inv = x_invert(z2, some);
  if (!some) z2_inv = x_small(0);
  x_to_be(out, x_mul(x2, z2_inv));
}

}  // namespace quirk25519
}  // namespace fec
