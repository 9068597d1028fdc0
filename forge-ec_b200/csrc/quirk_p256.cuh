// quirk_p256.cuh -- FEC_SEM_FORGE_QUIRK for P-256: replays the reference's LITERAL code path, one lane per thread.
//
// forge-ec's P-256 field is nearly the field (Add, Mul with the FIPS 186-4 Solinas sums and the LSB-first pow are right
// up to corner cases of probability ~2^-30 per operation, where the closing +-p loops of reduce_wide_p256 wrap modulo
// 2^256), but `Sub` adds p through `+=`, which reduces it away again, so a - b with a < b comes out as a - b + 2^256,
// and `double` is the a = 0 formula on a curve with a = -3.  Its outputs are therefore not
// FIPS 186-4 results and no re-associated algorithm can reproduce them: this mode repeats the same operations in the
// same order -- 4 x u64 limbs, 128-bit products and sums, MSB-first double-and-add with its data-dependent branches, a
// projective equality test in front of every addition, one LSB-first Fermat inversion per point.  A parity artefact,
// not a fast path.
//
// Every routine names the lines of forge-ec-curves/src/p256.rs it follows.  Host-compilable (tests/hostsim).
#pragma once
#include <cstdint>
#include "u256.cuh"

namespace fec {
namespace quirk256 {

typedef unsigned __int128 u128;
typedef __int128 i128;
struct PFe { u64 w[4]; };
struct PPt { PFe x, y, z; };

FEC_DEV u64 p_p(int i) { return i == 0 ? 0xFFFFFFFFFFFFFFFFULL : i == 1 ? 0x00000000FFFFFFFFULL : i == 2 ? 0ULL : 0xFFFFFFFF00000001ULL; }  // :18-19
FEC_DEV u64 p_n(int i) {                                                                                                                  // :23-24
  return i == 0 ? 0xF3B9CAC2FC632551ULL : i == 1 ? 0xBCE6FAADA7179E84ULL : i == 2 ? 0xFFFFFFFFFFFFFFFFULL : 0xFFFFFFFF00000000ULL;
}
FEC_DEV PFe p_small(u64 v) { PFe r; r.w[0] = v; r.w[1] = r.w[2] = r.w[3] = 0; return r; }  // From<u64> :869-873, zero/one :748-761
FEC_DEV PFe p_modulus() { PFe r; for (int i = 0; i < 4; i++) r.w[i] = p_p(i); return r; }
// compare :70-80
FEC_DEV int p_cmp(const PFe& a, const PFe& b) {
  for (int i = 3; i >= 0; i--) {
    if (a.w[i] < b.w[i]) return -1;
    if (a.w[i] > b.w[i]) return 1;
  }
  return 0;
}
// r = a + b (op 0) or a - b (op 1) on 256 bits, wrapping; returns the carry / borrow count of the limb loop
FEC_DEV u64 p_addsub(PFe& r, const PFe& a, const PFe& b, int op) {
  u64 c = 0;
  for (int i = 0; i < 4; i++) {
    if (op == 0) {
      const u64 s1 = a.w[i] + b.w[i], o1 = s1 < a.w[i];
      const u64 s2 = s1 + c, o2 = s2 < s1;
      r.w[i] = s2;
      c = o1 + o2;
    } else {
      const u64 d1 = a.w[i] - b.w[i], b1 = a.w[i] < b.w[i];
      const u64 d2 = d1 - c, b2 = d1 < c;
      r.w[i] = d2;
      c = b1 + b2;
    }
  }
  return c;
}
// reduce :88-100 -- subtract p while >= p
FEC_DEV void p_reduce(PFe& a) {
  const PFe m = p_modulus();
  while (p_cmp(a, m) >= 0) p_addsub(a, a, m, 1);
}
FEC_DEV bool p_is_zero(const PFe& a) { return (a.w[0] | a.w[1] | a.w[2] | a.w[3]) == 0; }  // :763-765
FEC_DEV bool p_same(const PFe& a, const PFe& b) {                                           // ct_eq :407-414
  return ((a.w[0] ^ b.w[0]) | (a.w[1] ^ b.w[1]) | (a.w[2] ^ b.w[2]) | (a.w[3] ^ b.w[3])) == 0;
}
FEC_DEV bool p_is_one(const PFe& a) { return p_same(a, p_small(1)); }  // :1342-1344
// impl Add :416-468
FEC_DEV PFe p_add(const PFe& a, const PFe& b) {
  PFe r;
  u64 carry = p_addsub(r, a, b, 0);
  PFe red;  // 2^256 - p
  red.w[0] = 1ULL; red.w[1] = 0xFFFFFFFF00000000ULL; red.w[2] = 0xFFFFFFFFFFFFFFFFULL; red.w[3] = 0x00000000FFFFFFFEULL;
  while (carry > 0) {
    const u64 add_carry = p_addsub(r, r, red, 0);
    carry = carry - 1 + add_carry;
  }
  p_reduce(r);
  return r;
}
// impl Sub :470-496 -- `result += P` goes through Add, which reduces the sum again; the subtraction then wraps
FEC_DEV PFe p_sub(const PFe& a, const PFe& b) {
  PFe r = a, d;
  if (p_cmp(a, b) < 0) r = p_add(r, p_modulus());
  p_addsub(d, r, b, 1);
  return d;
}
// impl Neg :707-729
FEC_DEV PFe p_neg(const PFe& a) {
  if (p_is_zero(a)) return a;
  PFe r;
  p_addsub(r, p_modulus(), a, 1);
  return r;
}
// reduce_wide_p256 :544-704
FEC_DEV PFe p_reduce_wide(const u64 wide[8]) {
  i128 c[16], acc[8];
  for (int i = 0; i < 8; i++) { c[2 * i] = (i128)(wide[i] & 0xFFFFFFFFULL); c[2 * i + 1] = (i128)(wide[i] >> 32); }
  for (int i = 0; i < 8; i++) acc[i] = c[i];
  acc[3] += 2 * c[11]; acc[4] += 2 * c[12]; acc[5] += 2 * c[13]; acc[6] += 2 * c[14]; acc[7] += 2 * c[15];
  acc[3] += 2 * c[12]; acc[4] += 2 * c[13]; acc[5] += 2 * c[14]; acc[6] += 2 * c[15];
  acc[0] += c[8]; acc[1] += c[9]; acc[2] += c[10]; acc[6] += c[14]; acc[7] += c[15];
  acc[0] += c[9]; acc[1] += c[10]; acc[2] += c[11]; acc[3] += c[13]; acc[4] += c[14]; acc[5] += c[15]; acc[6] += c[13]; acc[7] += c[8];
  acc[0] -= c[11]; acc[1] -= c[12]; acc[2] -= c[13]; acc[6] -= c[8]; acc[7] -= c[10];
  acc[0] -= c[12]; acc[1] -= c[13]; acc[2] -= c[14]; acc[3] -= c[15]; acc[6] -= c[9]; acc[7] -= c[11];
  acc[0] -= c[13]; acc[1] -= c[14]; acc[2] -= c[15]; acc[3] -= c[8]; acc[4] -= c[9]; acc[5] -= c[10]; acc[7] -= c[12];
  acc[0] -= c[14]; acc[1] -= c[15]; acc[3] -= c[9]; acc[4] -= c[10]; acc[5] -= c[11]; acc[7] -= c[13];
  for (int i = 0; i < 7; i++) {
    const i128 carry = acc[i] >> 32;  // arithmetic shift, as Rust's i128 >>
    acc[i] &= 0xFFFFFFFF;
    acc[i + 1] += carry;
  }
  i128 carry = acc[7] >> 32;
  acc[7] &= 0xFFFFFFFF;
  PFe r;
  for (int i = 0; i < 4; i++) r.w[i] = (u64)acc[2 * i] | ((u64)acc[2 * i + 1] << 32);
  const PFe m = p_modulus();
  while (carry > 0) { p_addsub(r, r, m, 1); carry -= 1; }  // wrapping, :672-683
  while (carry < 0) { p_addsub(r, r, m, 0); carry += 1; }  // wrapping, :686-696
  p_reduce(r);
  return r;
}
// impl Mul :498-534
FEC_DEV PFe p_mul(const PFe& a, const PFe& b) {
  u64 wide[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 carry = 0;
    for (int j = 0; j < 4; j++) {
      const u128 product = (u128)a.w[i] * (u128)b.w[j] + (u128)wide[i + j] + carry;
      wide[i + j] = (u64)product;
      carry = product >> 64;
    }
    const u128 sum = (u128)wide[i + 4] + carry;
    wide[i + 4] = (u64)sum;
    if ((sum >> 64) != 0) {
      for (int k = i + 5; k < 8; k++) {
        wide[k] += 1;
        if (wide[k] != 0) break;
      }
    }
  }
  return p_reduce_wide(wide);
}
FEC_DEV PFe p_sqr(const PFe& a) { return p_mul(a, a); }  // :772-776
// invert :343-370 = pow(P - 2) :376-393, least significant bit first
FEC_DEV PFe p_invert(const PFe& a) {
  if (p_is_zero(a)) return p_small(0);
  PFe result = p_small(1), base = a;
  for (int w = 0; w < 4; w++) {
    u64 e = w == 0 ? p_p(0) - 2 : p_p(w);
    for (int i = 0; i < 64; i++) {
      if (e & 1) result = p_mul(result, base);
      base = p_sqr(base);
      e >>= 1;
    }
  }
  return result;
}
FEC_DEV PFe p_from_be(const uint8_t* b) {  // from_bytes :303-317 (limbs only)
  PFe r = p_small(0);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) r.w[i] |= (u64)b[31 - (i * 8 + j)] << (j * 8);
  return r;
}
FEC_DEV void p_to_be(uint8_t* out, const PFe& a) {  // to_bytes :285-298
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) out[31 - (i * 8 + j)] = (uint8_t)(a.w[i] >> (j * 8));
}

FEC_DEV PPt p_identity() { PPt r; r.x = p_small(0); r.y = p_small(1); r.z = p_small(0); return r; }  // :1826-1828
FEC_DEV bool p_at_inf(const PPt& p) { return p_is_zero(p.z); }                                       // :1830-1832
// double :1869-1912 (the a = 0 formula, 3 * xx and 8 * yyyy by field multiplications)
FEC_DEV PPt p_double(const PPt& p) {
  if (p_at_inf(p)) return p_identity();
  const PFe xx = p_sqr(p.x), yy = p_sqr(p.y), yyyy = p_sqr(yy);
  const PFe xy2 = p_sqr(p_add(p.x, yy));
  const PFe t = p_sub(p_sub(xy2, xx), yyyy);
  const PFe d = p_add(t, t);
  const PFe e = p_mul(p_small(3), xx);
  const PFe ee = p_sqr(e);
  PPt r;
  r.x = p_sub(p_sub(ee, d), d);
  r.y = p_sub(p_mul(e, p_sub(d, r.x)), p_mul(p_small(8), yyyy));
  r.z = p_add(p.y, p.y);
  if (!p_is_one(p.z)) r.z = p_mul(r.z, p.z);
  return r;
}
// ConstantTimeEq for ProjectivePoint :2035-2067
FEC_DEV bool p_pt_same(const PPt& a, const PPt& b) {
  if (p_at_inf(a) && p_at_inf(b)) return true;
  if (p_at_inf(a) || p_at_inf(b)) return false;
  const PFe z1z1 = p_sqr(a.z), z2z2 = p_sqr(b.z);
  const PFe u1 = p_mul(a.x, z2z2), u2 = p_mul(b.x, z1z1);
  const PFe s1 = p_mul(p_mul(a.y, b.z), z2z2), s2 = p_mul(p_mul(b.y, a.z), z1z1);
  return p_same(u1, u2) & p_same(s1, s2);
}
// impl Add for ProjectivePoint :1941-2006
FEC_DEV PPt p_pt_add(const PPt& a, const PPt& b) {
  if (p_at_inf(a)) return b;
  if (p_at_inf(b)) return a;
  if (p_pt_same(a, b)) return p_double(a);
  const PFe z1z1 = p_sqr(a.z), z2z2 = p_sqr(b.z);
  const PFe u1 = p_mul(a.x, z2z2), u2 = p_mul(b.x, z1z1);
  const PFe s1 = p_mul(p_mul(a.y, b.z), z2z2), s2 = p_mul(p_mul(b.y, a.z), z1z1);
  if (p_same(u1, u2) && p_same(s1, p_neg(s2))) return p_identity();
  const PFe h = p_sub(u2, u1);
  const PFe i = p_sqr(p_add(h, h));
  const PFe j = p_mul(h, i);
  const PFe r = p_add(p_sub(s2, s1), p_sub(s2, s1));
  const PFe v = p_mul(u1, i);
  PPt o;
  o.x = p_sub(p_sub(p_sub(p_sqr(r), j), v), v);
  o.y = p_sub(p_mul(r, p_sub(v, o.x)), p_mul(p_add(s1, s1), j));
  o.z = p_mul(p_sub(p_sub(p_sqr(p_add(a.z, b.z)), z1z1), z2z2), h);
  return o;
}
// P256::generator :2092-2110
FEC_DEV PPt p_generator() {
  PPt g;
  g.x.w[0] = 0xF4A13945D898C296ULL; g.x.w[1] = 0x77037D812DEB33A0ULL; g.x.w[2] = 0xF8BCE6E563A440F2ULL; g.x.w[3] = 0x6B17D1F2E12C4247ULL;
  g.y.w[0] = 0xCBB6406837BF51F5ULL; g.y.w[1] = 0x2BCE33576B315ECEULL; g.y.w[2] = 0x8EE7EB4A7C0F9E16ULL; g.y.w[3] = 0x4FE342E2FE1A7F9BULL;
  g.z = p_small(1);
  return g;
}
// P256::multiply :2120-2156 -- MSB-first double-and-add over the big-endian bytes of the scalar
FEC_DEV PPt p_multiply(const PPt& point, const uint8_t* k_be) {
  bool zero = true;
  for (int i = 0; i < 32; i++) zero = zero && k_be[i] == 0;
  if (p_at_inf(point) || zero) return p_identity();
  PPt result = p_identity();
  for (int i = 0; i < 256; i++) {
    const int bit = (k_be[i / 8] >> (7 - (i % 8))) & 1;
    result = p_double(result);
    if (bit) result = p_pt_add(result, point);
  }
  return result;
}
// one lane: Scalar::from_bytes (:1041-1055) -> generator() | from_affine (:1859-1867) -> multiply -> to_affine
// (:1835-1857) -> encodings (AffinePoint::to_bytes :1558-1577 / point.rs:39-67, 186-211).  Returns the lane status.
FEC_DEV int p_lane(const uint8_t* k_be, const uint8_t* xy_be, int out_fmt, uint8_t* out) {
  const int osz = out_fmt == 0 ? 33 : (out_fmt == 1 ? 65 : 64);
  for (int i = 0; i < osz; i++) out[i] = 0;
  const PFe k = p_from_be(k_be);
  PFe nn;
  for (int i = 0; i < 4; i++) nn.w[i] = p_n(i);
  if (p_cmp(k, nn) >= 0) return 1;  // CtOption::none: the caller never gets a scalar
  PPt p;
  if (xy_be == nullptr) {
    p = p_generator();
  } else {
    p.x = p_from_be(xy_be);
    p.y = p_from_be(xy_be + 32);
    p.z = p_small(1);
  }
  const PPt r = p_multiply(p, k_be);
  if (p_at_inf(r)) return 3;
  const PFe zi = p_invert(r.z), zi2 = p_sqr(zi), zi3 = p_mul(zi2, zi);
  const PFe x = p_mul(r.x, zi2), y = p_mul(r.y, zi3);
  uint8_t yb[32];
  p_to_be(yb, y);
  if (out_fmt == 0) {
    out[0] = (yb[31] & 1) ? 3 : 2;
    p_to_be(out + 1, x);
  } else if (out_fmt == 1) {
    out[0] = 4;
    p_to_be(out + 1, x);
    for (int i = 0; i < 32; i++) out[33 + i] = yb[i];
  } else {
    p_to_be(out, x);
    for (int i = 0; i < 32; i++) out[32 + i] = yb[i];
  }
  return 0;
}

}  // namespace quirk256
}  // namespace fec
