// quirk_k1.cuh -- FEC_SEM_FORGE_QUIRK for secp256k1: replays the reference's LITERAL limb code, one lane per thread.
//
// forge-ec's secp256k1 field code does not implement the field (the Montgomery product drops
// the carry out of t[7], `square` is a different truncating fold, R^2 is wrong, the ladder reads a byte-reversed
// scalar), so its outputs are NOT SEC2 results and no re-associated algorithm (windows, GLV, shared inversion) can
// reproduce them.  This mode exists to give one literally bit-exact artefact for BASELINE.json configs[0]
// (k*G over 2^16 scalars): the same operations in the same order, 4 x u64 limbs with 128-bit products, release-mode
// wrapping arithmetic.  It is a parity artefact, not a fast path (data-dependent branches, 256 ladder steps of
// one addition and two doublings, a Fermat inversion per point).
//
// Every routine names the lines of forge-ec-curves/src/secp256k1.rs it follows.  Host-compilable (tests/hostsim).
#pragma once
#include <cstdint>
#include "u256.cuh"

namespace fec {
namespace quirk {

typedef unsigned __int128 u128;
struct QFe { u64 w[4]; };
struct QPt { QFe x, y, z; };

FEC_DEV u64 q_p(int i) { return i == 0 ? 0xFFFFFFFEFFFFFC2FULL : 0xFFFFFFFFFFFFFFFFULL; }  // :22-23
// the reference's N constant as written: its two top limbs are swapped with respect to the group order (:27-28)
FEC_DEV u64 q_n(int i) {
  return i == 0 ? 0xBFD25E8CD0364141ULL : i == 1 ? 0xBAAEDCE6AF48A03BULL : i == 2 ? 0xFFFFFFFFFFFFFFFFULL : 0xFFFFFFFFFFFFFFFEULL;
}

// compare_with_p :47-75 -- sign of (a - p), most significant limb decides
FEC_DEV int q_cmp_p(const QFe& a) {
  int res = 0;
  for (int i = 3; i >= 0; i--) {
    const int c = a.w[i] < q_p(i) ? -1 : (a.w[i] > q_p(i) ? 1 : 0);
    if (res == 0) res = c;
  }
  return res;
}
// a - p with wrap-around (the subtraction loop shared by reduce :78-102 and Add :373-388)
FEC_DEV QFe q_minus_p(const QFe& a) {
  QFe r;
  u64 bw = 0;
  for (int i = 0; i < 4; i++) {
    const u64 d = a.w[i] - q_p(i);
    const u64 b1 = a.w[i] < q_p(i);
    r.w[i] = d - bw;
    bw = b1 | (u64)(d < bw);
  }
  return r;
}
FEC_DEV QFe q_reduce(const QFe& a) { return q_cmp_p(a) >= 0 ? q_minus_p(a) : a; }  // :78-102
FEC_DEV bool q_is_zero(const QFe& a) { return (a.w[0] | a.w[1] | a.w[2] | a.w[3]) == 0; }  // :595-597
FEC_DEV bool q_same(const QFe& a, const QFe& b) {                                           // ct_eq :344-351
  return ((a.w[0] ^ b.w[0]) | (a.w[1] ^ b.w[1]) | (a.w[2] ^ b.w[2]) | (a.w[3] ^ b.w[3])) == 0;
}
// impl Add :353-393
FEC_DEV QFe q_add(const QFe& a, const QFe& b) {
  QFe s;
  u64 cy = 0;
  for (int i = 0; i < 4; i++) {
    const u64 t = a.w[i] + b.w[i];
    const u64 c1 = a.w[i] > ~b.w[i];
    const u64 c2 = t > (~(u64)0 - cy);
    s.w[i] = t + cy;
    cy = c1 | c2;
  }
  return (cy != 0 || q_cmp_p(s) >= 0) ? q_minus_p(s) : s;
}
// impl Sub :395-440 -- on borrow, p is added back modulo 2^256
FEC_DEV QFe q_sub(const QFe& a, const QFe& b) {
  QFe d;
  u64 bw = 0;
  for (int i = 0; i < 4; i++) {
    const u64 t = a.w[i] - b.w[i];
    const u64 b1 = a.w[i] < b.w[i];
    const u64 b2 = t < bw;
    d.w[i] = t - bw;
    bw = b1 | b2;
  }
  if (bw == 0) return d;
  QFe r;
  u64 cy = 0;
  for (int i = 0; i < 4; i++) {
    const u64 t = d.w[i] + q_p(i);
    const u64 c1 = d.w[i] > ~q_p(i);
    const u64 c2 = t > (~(u64)0 - cy);
    r.w[i] = t + cy;
    cy = c1 | c2;
  }
  return r;
}
// the reduction rounds shared by Mul :470-494 and mont_reduce :270-294, 314-331.  `cy` is carried from one round
// into the next, and what is left of it after limb 7 is dropped -- exactly as in the reference.
FEC_DEV QFe q_mont_rounds(u64 t[8]) {
  const u64 n0 = 0xD838091DD2253531ULL;
  u64 cy = 0;
  for (int i = 0; i < 4; i++) {
    const u64 m = t[i] * n0;
    u128 acc = (u128)t[i] + (u128)m * q_p(0) + cy;
    cy = (u64)(acc >> 64);
    for (int j = 1; j < 4; j++) {
      acc = (u128)t[i + j] + (u128)m * q_p(j) + cy;
      t[i + j] = (u64)acc;
      cy = (u64)(acc >> 64);
    }
    for (int j = i + 4; j < 8 && cy > 0; j++) {
      const u128 s = (u128)t[j] + cy;
      t[j] = (u64)s;
      cy = (u64)(s >> 64);
    }
  }
  QFe r = {{t[4], t[5], t[6], t[7]}};
  if (q_cmp_p(r) >= 0) r = q_reduce(r);
  return r;
}
// impl Mul :442-507
FEC_DEV QFe q_mul(const QFe& a, const QFe& b) {
  u64 t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u64 cy = 0;
    for (int j = 0; j < 4; j++) {
      const u128 pr = (u128)a.w[i] * b.w[j] + t[i + j] + cy;
      t[i + j] = (u64)pr;
      cy = (u64)(pr >> 64);
    }
    t[i + 4] = cy;
  }
  return q_mont_rounds(t);
}
FEC_DEV QFe q_from_mont(const QFe& a) {  // mont_reduce :314-331
  u64 t[8] = {a.w[0], a.w[1], a.w[2], a.w[3], 0, 0, 0, 0};
  return q_mont_rounds(t);
}
FEC_DEV QFe q_to_mont(const QFe& a) {  // to_montgomery :219-235 with the reference's R_SQUARED constant (:225-230)
  const QFe r2 = {{0x000E9F61ULL, 0x07A20000ULL, 0x00000100ULL, 0ULL}};
  return q_mul(a, r2);
}
// square :634-713 -- diagonal terms, cross terms doubled in 128 bits (top bit lost), carries rippled word by
// word, then the fold multiplies each high word by 0x1000003D1 in 64 bits (high half lost)
FEC_DEV QFe q_sqr(const QFe& a) {
  u64 pr[8];
  for (int i = 0; i < 4; i++) {
    const u128 s = (u128)a.w[i] * a.w[i];
    pr[2 * i] = (u64)s;
    pr[2 * i + 1] = (u64)(s >> 64);
  }
  for (int i = 0; i < 4; i++)
    for (int j = i + 1; j < 4; j++) {
      const u128 cr = ((u128)a.w[i] * a.w[j]) << 1;
      const u64 lo = (u64)cr, hi = (u64)(cr >> 64);
      pr[i + j] += lo;
      const bool c1 = pr[i + j] < lo;
      pr[i + j + 1] += hi;
      const bool c2 = pr[i + j + 1] < hi;
      if (c1 || c2)
        for (int k = i + j + 2; k < 8; k++)
          if (++pr[k] != 0) break;
    }
  u64 r[4] = {pr[0], pr[1], pr[2], pr[3]};
  u64 cy = 0;
  for (int i = 4; i < 8; i++) {
    const u64 wv = pr[i] * 0x1000003D1ULL;
    u64 t = r[0] + wv;
    t += cy;
    r[0] = t;
    cy = (u64)(t < wv) | ((u64)(t < cy) & (u64)(wv != 0));
    for (int j = 1; j < 4; j++) {
      const u64 tt = r[j] + cy;
      r[j] = tt;
      cy = tt < cy;
    }
  }
  const QFe out = {{r[0], r[1], r[2], r[3]}};
  return q_reduce(out);
}
FEC_DEV QFe q_small(u64 v) { const QFe r = {{v, 0, 0, 0}}; return r; }  // one() :585-593 is the plain integer 1
// invert :599-632 -- exponent p - 2 scanned limb 0 first, each limb from bit 63 down
FEC_DEV QFe q_invert(const QFe& a) {
  QFe acc = q_small(1);
  for (int i = 0; i < 4; i++) {
    const u64 e = i == 0 ? 0xFFFFFFFEFFFFFC2DULL : 0xFFFFFFFFFFFFFFFFULL;
    for (int j = 63; j >= 0; j--) {
      acc = q_sqr(acc);
      if ((e >> j) & 1) acc = q_mul(acc, a);
    }
  }
  return acc;
}
FEC_DEV QFe q_from_be(const uint8_t* b) {  // from_bytes :182-213
  QFe r = q_small(0);
  for (int i = 0; i < 32; i++) r.w[3 - i / 8] |= (u64)b[i] << (56 - 8 * (i % 8));
  return q_to_mont(r);
}
FEC_DEV void q_to_be(uint8_t* o, const QFe& a) {  // to_bytes :138-179
  const QFe t = q_from_mont(a);
  for (int i = 0; i < 32; i++) o[i] = (uint8_t)(t.w[3 - i / 8] >> (56 - 8 * (i % 8)));
}

FEC_DEV QPt q_identity() { const QPt r = {q_small(0), q_small(1), q_small(0)}; return r; }  // :1322-1324
FEC_DEV bool q_at_inf(const QPt& p) { return q_is_zero(p.z); }                                // :1326-1340
// inherent double :1502-1540 (the one `multiply` resolves to)
FEC_DEV QPt q_double(const QPt& p) {
  if (q_at_inf(p)) return q_identity();
  const QFe a = q_sqr(p.x), b = q_sqr(p.y), c = q_sqr(b);
  QFe d = q_sub(q_sub(q_sqr(q_add(p.x, b)), a), c);
  d = q_add(d, d);
  const QFe e = q_mul(a, q_small(3)), f = q_sqr(e);
  QPt r;
  r.x = q_sub(f, q_add(d, d));
  r.y = q_sub(q_mul(e, q_sub(d, r.x)), q_mul(c, q_small(8)));
  const QFe yz = q_mul(p.y, p.z);
  r.z = q_add(yz, yz);
  return r;
}
// impl Add :1447-1497
FEC_DEV QPt q_padd(const QPt& p, const QPt& q) {
  if (q_at_inf(p)) return q;
  if (q_at_inf(q)) return p;
  const QFe z1s = q_sqr(p.z), z2s = q_sqr(q.z);
  const QFe u1 = q_mul(p.x, z2s), u2 = q_mul(q.x, z1s);
  const QFe z1c = q_mul(z1s, p.z), z2c = q_mul(z2s, q.z);
  const QFe s1 = q_mul(p.y, z2c), s2 = q_mul(q.y, z1c);
  if (q_same(u1, u2)) return q_same(s1, s2) ? q_double(p) : q_identity();
  const QFe h = q_sub(u2, u1), r = q_sub(s2, s1);
  const QFe h2 = q_sqr(h), h3 = q_mul(h2, h), v = q_mul(u1, h2);
  QPt o;
  o.x = q_sub(q_sub(q_sub(q_sqr(r), h3), v), v);
  o.y = q_sub(q_mul(r, q_sub(v, o.x)), q_mul(s1, h3));
  o.z = q_mul(q_mul(h, p.z), q.z);
  return o;
}
FEC_DEV QPt q_generator() {  // Secp256k1::generator :2608-2625
  const QFe gx = {{0x59F2815B16F81798ULL, 0x029BFCDB2DCE28D9ULL, 0x55A06295CE870B07ULL, 0x79BE667EF9DCBBACULL}};
  const QFe gy = {{0x9C47D08FFB10D4B8ULL, 0xFD17B448A6855419ULL, 0x5DA4FBFC0E1108A8ULL, 0x483ADA7726A3C465ULL}};
  const QPt g = {q_to_mont(gx), q_to_mont(gy), q_small(1)};
  return g;
}
// Secp256k1::multiply :2635-2692.  k: limbs as Scalar::from_bytes leaves them (:2271-2297); `scalar.to_bytes()`
// at :2644 is the inherent little-endian one (:1924-1933) and :2655-2659 walks those bytes "most significant first".
FEC_DEV QPt q_multiply(const QPt& pt, const u64 k[4]) {
  if (q_at_inf(pt) || (k[0] | k[1] | k[2] | k[3]) == 0) return q_identity();
  QPt r0 = q_identity(), r1 = pt;
  for (int i = 0; i < 256; i++) {
    const int byte = i >> 3;
    const u64 bit = (k[byte >> 3] >> (8 * (byte & 7) + (7 - (i & 7)))) & 1u;
    const QPt sum = q_padd(r0, r1), d0 = q_double(r0), d1 = q_double(r1);
    r0 = bit ? sum : d0;  // conditional_select(a, b, c) yields b iff c == 1
    r1 = bit ? d1 : sum;
  }
  return r0;
}

// one lane: Scalar::from_bytes -> generator() | from_affine -> multiply -> to_affine (:1342-1363) -> encodings
// (AffinePoint::to_bytes :875-894 / point.rs:39-67, 186-211).  Returns the lane status.
FEC_DEV int q_lane(const uint8_t* k_be, const uint8_t* xy_be, int out_fmt, uint8_t* out) {
  const int osz = out_fmt == 0 ? 33 : (out_fmt == 1 ? 65 : 64);
  for (int i = 0; i < osz; i++) out[i] = 0;
  u64 k[4] = {0, 0, 0, 0};
  for (int i = 0; i < 32; i++) k[3 - i / 8] |= (u64)k_be[i] << (56 - 8 * (i % 8));
  bool lt = false, decided = false;  // k < n
  for (int i = 3; i >= 0 && !decided; i--)
    if (k[i] != q_n(i)) { lt = k[i] < q_n(i); decided = true; }
  if (!lt) return 1;  // CtOption::none: the caller never gets a scalar
  QPt p;
  if (xy_be == nullptr) {
    p = q_generator();
  } else {
    p.x = q_from_be(xy_be);
    p.y = q_from_be(xy_be + 32);
    p.z = q_small(1);  // from_affine :1365-1373
  }
  const QPt r = q_multiply(p, k);
  if (q_at_inf(r)) return 3;
  const QFe zi = q_invert(r.z), zi2 = q_sqr(zi), zi3 = q_mul(zi2, zi);
  const QFe x = q_mul(r.x, zi2), y = q_mul(r.y, zi3);
  uint8_t yb[32];
  q_to_be(yb, y);
  if (out_fmt == 0) {
    out[0] = (yb[31] & 1) ? 3 : 2;
    q_to_be(out + 1, x);
  } else if (out_fmt == 1) {
    out[0] = 4;
    q_to_be(out + 1, x);
    for (int i = 0; i < 32; i++) out[33 + i] = yb[i];
  } else {
    q_to_be(out, x);
    for (int i = 0; i < 32; i++) out[32 + i] = yb[i];
  }
  return 0;
}

}  // namespace quirk
}  // namespace fec
