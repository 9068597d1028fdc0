// fields.cuh -- the three prime fields of the hot path on 8 x u32 register limbs.
//
//   FeK1    p = 2^256 - 2^32 - 977   (secp256k1; replaces forge-ec-curves/src/secp256k1.rs:353-539, 599-736)
//   FeP256  p = 2^256 - 2^224 + 2^192 + 2^96 - 1  (replaces p256.rs:343-393, 416-534, 544-729, 772-776)
//   Fe25519 p = 2^255 - 19           (replaces curve25519.rs:50-114, 186-314, 369-494 and
//                                     ed25519.rs:214-289, 410-431, 458-625)
//
// Representation invariants
//   FeK1 / Fe25519: "weak" form -- any value in [0, 2^256) stands for its residue; add/sub/mul
//     fold the overflow with 2^256 = 2^32+977 (resp. 38) and never need a comparison with p.
//     canon() produces the unique representative in [0,p) for output / equality.
//   FeP256: Montgomery domain (x held as x 2^256 mod p); weak since r2 -- any representative below 2^256, which only the residues
//     below 2^256 - p have two of: what the form buys is the end of the Montgomery reduction (a masked subtraction instead of
//     compare-and-select), at the price of a second correction in add / sub that sits under a warp-uniform branch (2^-32 per
//     operation).  FEC_P256_CANONICAL restores [0,p).  from_plain / to_plain convert at the byte-record boundary.
// All functions are branch-free (masks only): control flow never depends on data.
//
// IMAD.WIDE count per op (the canonical IMUL32 figures of SURVEY.md section 8d):
//   FeK1 mul 74 / sqr 46, FeP256 mul 64 / sqr 36, Fe25519 mul 73 / sqr 45.
#pragma once
#include "u256.cuh"

namespace fec {

// Field multiplication / squaring entry points.  With FEC_NOINLINE_FIELD defined (device code of the scalar
// multiplication kernels) they are calls to ONE out-of-line subroutine per field and kernel image, operands and
// result in registers: the hot loops shrink from ~70 KB to ~15 KB of SASS and stay resident in the instruction
// cache (measured: instruction-fetch stalls were the top stall reason with everything inlined).  Otherwise
// (microbenchmark, host simulation) they are inlined.
template <class F> FEC_DEV void field_mul(u32 o[8], const u32 a[8], const u32 b[8]);
template <class F> FEC_DEV void field_sqr(u32 o[8], const u32 a[8]);

// =====================================================================================
// secp256k1 base field
// =====================================================================================
struct FeK1 {
  static constexpr int PRIME_ID = 0;
  static constexpr u32 C977 = 977u;
  FEC_DEV static void P(u32 p[8]) {
    p[0] = 0xFFFFFC2Fu; p[1] = 0xFFFFFFFEu;
#pragma unroll
    for (int i = 2; i < 8; i++) p[i] = 0xFFFFFFFFu;
  }

  // 512 -> 256 bits:  lo + hi * (2^32 + 977), twice, plus a 1-bit wrap fix.
  FEC_DEV static void reduce512(u32 o[8], const u32 r[16]) {
    const u32* h = r + 8;
    u32 acc[8];
    copy8(acc, r);
    u32 ca = mad_row<4, 0>(acc, h, C977);  // += {h0,h2,h4,h6}*977 at limb pairs (0,1)..(6,7)
    // s = {h1,h3,h5,h7}*977 + hi, both at weight 2^32 (the "<< 32" half of the fold): hi enters as the 64-bit ADDEND of the four
    // products -- (h[2k], h[2k+1]) is the pair the product of h[2k+1] lands on -- so the sum costs no additions of its own
    // (ptxas finds the same fusion when the sum is written as products + an 8-word addition)
    u32 s[8];
    s[0] = mad_lo_cc(h[1], C977, h[0]);
    s[1] = madc_hi_cc(h[1], C977, h[1]);
#pragma unroll
    for (int k = 1; k < 4; k++) {
      s[2 * k] = madc_lo_cc(h[2 * k + 1], C977, h[2 * k]);
      s[2 * k + 1] = madc_hi_cc(h[2 * k + 1], C977, h[2 * k + 1]);
    }
    u32 s8 = addc(0u, 0u);
    acc[1] = add_cc(acc[1], s[0]);
#pragma unroll
    for (int i = 2; i < 8; i++) acc[i] = addc_cc(acc[i], s[i - 1]);
    u32 v_lo = addc_cc(s[7], ca);  // overflow above 2^256:  v = v_lo + v_hi*2^32  (< 2^32 + 979)
    u32 v_hi = addc(s8, 0u);
    // second fold: acc += v*977 + (v << 32)   (three limbs)
    u64 m = (u64)v_lo * C977 + ((u64)(v_hi * C977) << 32);
    u32 w0 = (u32)m;
    u64 t = (m >> 32) + v_lo;
    u32 w1 = (u32)t;
    u32 w2 = (u32)(t >> 32) + v_hi;
    acc[0] = add_cc(acc[0], w0);
    acc[1] = addc_cc(acc[1], w1);
    acc[2] = addc_cc(acc[2], w2);
#pragma unroll
    for (int i = 3; i < 8; i++) acc[i] = addc_cc(acc[i], 0u);
    u32 c2 = addc(0u, 0u);
    // wrap fix: if that carried, the wrapped value is < 2^67, so adding 2^32+977 touches 3 limbs
    o[0] = add_cc(acc[0], (0u - c2) & C977);
    o[1] = addc_cc(acc[1], c2);
    o[2] = addc(acc[2], 0u);
#pragma unroll
    for (int i = 3; i < 8; i++) o[i] = acc[i];
  }
  FEC_DEV static void mul_inl(u32 o[8], const u32 a[8], const u32 b[8]) {
    u32 r[16];
    mul_wide(r, a, b);
    reduce512(o, r);
  }
  FEC_DEV static void sqr_inl(u32 o[8], const u32 a[8]) {
    u32 r[16];
    sqr_wide(r, a);
    reduce512(o, r);
  }
  // mul / sqr: inlined, or (FEC_NOINLINE_FIELD, the big kernels) one shared subroutine per kernel image
  FEC_DEV static void mul(u32 o[8], const u32 a[8], const u32 b[8]) { field_mul<FeK1>(o, a, b); }
  FEC_DEV static void sqr(u32 o[8], const u32 a[8]) { field_sqr<FeK1>(o, a); }
  FEC_DEV static void add(u32 o[8], const u32 a[8], const u32 b[8]) {
    u32 s[8];
    u32 c = add8(s, a, b);
    // + c*(2^32+977); may wrap once more only if the sum was within 2^33 of 2^256
    s[0] = add_cc(s[0], (0u - c) & C977);
    s[1] = addc_cc(s[1], c);
#pragma unroll
    for (int i = 2; i < 8; i++) s[i] = addc_cc(s[i], 0u);
    u32 c2 = addc(0u, 0u);
    o[0] = add_cc(s[0], (0u - c2) & C977);
    o[1] = addc(s[1], c2);
#pragma unroll
    for (int i = 2; i < 8; i++) o[i] = s[i];
  }
  FEC_DEV static void sub(u32 o[8], const u32 a[8], const u32 b[8]) {
    u32 d[8];
    u32 bw = sub8(d, a, b);
    // - bw*(2^32+977)  (== + p mod 2^256); may borrow once more only if d < 2^33
    d[0] = sub_cc(d[0], (0u - bw) & C977);
    d[1] = subc_cc(d[1], bw);
#pragma unroll
    for (int i = 2; i < 8; i++) d[i] = subc_cc(d[i], 0u);
    u32 bw2 = subc(0u, 0u) & 1u;
    o[0] = sub_cc(d[0], (0u - bw2) & C977);
    o[1] = subc(d[1], bw2);
#pragma unroll
    for (int i = 2; i < 8; i++) o[i] = d[i];
  }
  FEC_DEV static void neg(u32 o[8], const u32 a[8]) {
    u32 z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    sub(o, z, a);
  }
  // ---- chains of linear operations with ONE fold (r2).  A weak-form addition or subtraction is 8 instructions of arithmetic
  // and ~16 of folding the carry back in; a chain a - b - c ... keeps the 256-bit value wrapped and counts the wraps in a small
  // signed v instead (value = t + v 2^256), and folds once: o = t + v (2^32 + 977).
  // extra (0 or 1, and 0 when v == 0 ... or any v <= 0) rides in the low limb: o = t + v (2^32 + 977) + extra
  FEC_DEV static void fold_signed(u32 o[8], const u32 t[8], u32 v, u32 extra = 0u) {  // |v| < 2^20 (as a signed number)
    const u32 s = (u32)((int)v >> 31);
    u32 x[8];
    x[0] = add_cc(t[0], v * C977 + extra);  // v (2^32 + 977) as a 256-bit two's-complement number: (v 977, v + s, s, s, s, s, s, s)
    x[1] = addc_cc(t[1], v + s);
#pragma unroll
    for (int i = 2; i < 8; i++) x[i] = addc_cc(t[i], s);
    const u32 corr = addc(s, 0u);   // carry + s: +1 wrapped above 2^256, -1 wrapped below 0 (the value is then within 2^37 of the wrap)
    const u32 cs = (u32)((int)corr >> 31);
    o[0] = add_cc(x[0], corr * C977);
    o[1] = addc_cc(x[1], corr + cs);
    o[2] = addc(x[2], cs);
#pragma unroll
    for (int i = 3; i < 8; i++) o[i] = x[i];
  }
  // t -= b, wraps counted in v (a borrow is -1)
  FEC_DEV static void sub_acc(u32 t[8], u32& v, const u32 b[8]) {
    t[0] = sub_cc(t[0], b[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = subc_cc(t[i], b[i]);
    v += subc(0u, 0u);  // 0 or 0xFFFFFFFF
  }
  FEC_DEV static void sub2_dbl(u32 o[8], const u32 t[8], const u32 a[8], const u32 b[8]) {  // 2 (t - a - b)
    u32 d[8], v = 0u;
    copy8(d, t);
    sub_acc(d, v, a);
    sub_acc(d, v, b);
    v = (v << 1) + (d[7] >> 31);
#pragma unroll
    for (int i = 7; i >= 1; i--) d[i] = funnel_l(d[i - 1], d[i], 1);
    d[0] <<= 1;
    fold_signed(o, d, v);
  }
  // the same with the fold left to the consumers: the value is d + v 2^256 (v in -4 .. 1), d goes to a slot and v stays in a
  // register -- for a value that only enters further chains (the D of a doubling) this saves its fold
  FEC_DEV static void sub2_dbl_raw(u32 d[8], u32& v, const u32 t[8], const u32 a[8], const u32 b[8]) {
    v = 0u;
    copy8(d, t);
    sub_acc(d, v, a);
    sub_acc(d, v, b);
    v = (v << 1) + (d[7] >> 31);
#pragma unroll
    for (int i = 7; i >= 1; i--) d[i] = funnel_l(d[i - 1], d[i], 1);
    d[0] <<= 1;
  }
  FEC_DEV static void sub_twice_raw(u32 o[8], const u32 t[8], const u32 a[8], u32 va) {  // t - 2 (a + va 2^256)
    u32 d[8], v = 0u - (va << 1);
    copy8(d, t);
    sub_acc(d, v, a);
    sub_acc(d, v, a);
    fold_signed(o, d, v);
  }
  FEC_DEV static void rsub_raw(u32 o[8], const u32 a[8], u32 va, const u32 x[8]) {  // (a + va 2^256) - x
    u32 d[8], v = va;
    copy8(d, a);
    sub_acc(d, v, x);
    fold_signed(o, d, v);
  }
  FEC_DEV static void sub_twice(u32 o[8], const u32 t[8], const u32 a[8]) {  // t - 2a
    u32 d[8], v = 0u;
    copy8(d, t);
    sub_acc(d, v, a);
    sub_acc(d, v, a);
    fold_signed(o, d, v);
  }
  FEC_DEV static void sub3(u32 o[8], const u32 t[8], const u32 a[8], const u32 b[8]) {  // t - a - 2b
    u32 d[8], v = 0u;
    copy8(d, t);
    sub_acc(d, v, a);
    sub_acc(d, v, b);
    sub_acc(d, v, b);
    fold_signed(o, d, v);
  }
  FEC_DEV static void sub_mul8(u32 o[8], const u32 t[8], const u32 a[8]) {  // t - 8a
    u32 e[8], d[8], v = 0u - (a[7] >> 29);
#pragma unroll
    for (int i = 7; i >= 1; i--) e[i] = funnel_l(a[i - 1], a[i], 3);
    e[0] = a[0] << 3;
    copy8(d, t);
    sub_acc(d, v, e);
    fold_signed(o, d, v);
  }
  // (m ? -a : a) - b for a full mask m.  -a = ~a + 1 - 2^256 == ~a + 1 - (2^32 + 977): the complement is one XOR per limb, and
  // its "- (2^32 + 977)" joins the wrap count of the subtraction (v starts at -1), its "+ 1" the low limb of the fold.
  FEC_DEV static void cneg_sub(u32 o[8], const u32 a[8], u32 m, const u32 b[8]) {
    u32 d[8], v = m;
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = a[i] ^ m;
    sub_acc(d, v, b);
    fold_signed(o, d, v, m & 1u);
  }
  static constexpr bool HAS_FUSED_LINEAR = true;
  // o = a * c for a small constant c (< 2^16): one 9-limb product + fold
  FEC_DEV static void mul_small(u32 o[8], const u32 a[8], u32 c) {
    u32 t[8];
    u64 carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      u64 m = (u64)a[i] * c + carry;
      t[i] = (u32)m;
      carry = m >> 32;
    }
    // fold carry (< 2^16): carry*(2^32+977)
    u32 v = (u32)carry;
    u64 m = (u64)v * C977;
    u32 w0 = (u32)m;
    u32 w1 = (u32)(m >> 32) + v;  // < 2^17, no overflow
    t[0] = add_cc(t[0], w0);
    t[1] = addc_cc(t[1], w1);
#pragma unroll
    for (int i = 2; i < 8; i++) t[i] = addc_cc(t[i], 0u);
    u32 c2 = addc(0u, 0u);
    o[0] = add_cc(t[0], (0u - c2) & C977);
    o[1] = addc(t[1], c2);
#pragma unroll
    for (int i = 2; i < 8; i++) o[i] = t[i];
  }
  FEC_DEV static void dbl(u32 o[8], const u32 a[8]) { add(o, a, a); }
  // unique representative in [0,p)
  FEC_DEV static void canon(u32 o[8], const u32 a[8]) {
    u32 t[8];
    t[0] = add_cc(a[0], C977);
    t[1] = addc_cc(a[1], 1u);
#pragma unroll
    for (int i = 2; i < 8; i++) t[i] = addc_cc(a[i], 0u);
    u32 c = addc(0u, 0u);  // a + (2^256 - p) overflowed  <=>  a >= p
    select8(o, 0u - c, t, a);
  }
  FEC_DEV static u32 is_zero(const u32 a[8]) {  // mask; accepts both representatives 0 and p
    const u32 o = a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7];
    const u32 n = (a[0] ^ 0xFFFFFC2Fu) | (a[1] ^ 0xFFFFFFFEu) | ~(a[2] & a[3] & a[4] & a[5] & a[6] & a[7]);
    return ((o == 0u) | (n == 0u)) ? 0xFFFFFFFFu : 0u;
  }
  FEC_DEV static u32 in_range(const u32 a[8]) {  // mask: a < p (input validation)
    u32 p[8];
    P(p);
    return ~geq8(a, p);
  }
  // a cheap NECESSARY condition for a == 0 (mod p): the representatives of zero below 2^256 are 0 and p, top word all zeros or all
  // ones.  Non-zero where the exact test must be asked (2^-31 of all values).
  FEC_DEV static u32 maybe_zero(const u32 a[8]) { return (a[7] + 1u <= 1u) ? 1u : 0u; }
  // byte records <-> the internal representation (here the integer itself; FeP256 works in the Montgomery domain)
  FEC_DEV static void one(u32 o[8]) { o[0] = 1u; o[1] = o[2] = o[3] = o[4] = o[5] = o[6] = o[7] = 0u; }
  FEC_DEV static void from_plain(u32 o[8], const u32 a[8]) { copy8(o, a); }
  FEC_DEV static void to_plain(u32 o[8], const u32 a[8]) { canon(o, a); }
  // o = a * 2^S (S = 1 .. 4): one funnel-shift pass, the S bits shifted out are folded with 2^32 + 977
  template <int S>
  FEC_DEV static void mul_pow2(u32 o[8], const u32 a[8]) {
    const u32 v = a[7] >> (32 - S);
    u32 t[8];
#pragma unroll
    for (int i = 7; i >= 1; i--) t[i] = funnel_l(a[i - 1], a[i], S);
    t[0] = a[0] << S;
    t[0] = add_cc(t[0], v * C977);
    t[1] = addc_cc(t[1], v);
#pragma unroll
    for (int i = 2; i < 8; i++) t[i] = addc_cc(t[i], 0u);
    const u32 c2 = addc(0u, 0u);
    o[0] = add_cc(t[0], (0u - c2) & C977);
    o[1] = addc_cc(t[1], c2);
    o[2] = addc(t[2], 0u);
#pragma unroll
    for (int i = 3; i < 8; i++) o[i] = t[i];
  }
  FEC_DEV static void mul2(u32 o[8], const u32 a[8]) { mul_pow2<1>(o, a); }
  FEC_DEV static void mul4(u32 o[8], const u32 a[8]) { mul_pow2<2>(o, a); }
  FEC_DEV static void mul8(u32 o[8], const u32 a[8]) { mul_pow2<3>(o, a); }
  FEC_DEV static void mul3(u32 o[8], const u32 a[8]) {  // 2a + a in one carry pass, overflow (0..2) folded once
    u32 t[8];
    t[0] = add_cc(a[0] << 1, a[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = addc_cc(funnel_l(a[i - 1], a[i], 1), a[i]);
    const u32 v = addc(a[7] >> 31, 0u);
    t[0] = add_cc(t[0], v * C977);
    t[1] = addc_cc(t[1], v);
#pragma unroll
    for (int i = 2; i < 8; i++) t[i] = addc_cc(t[i], 0u);
    const u32 c2 = addc(0u, 0u);
    o[0] = add_cc(t[0], (0u - c2) & C977);
    o[1] = addc_cc(t[1], c2);
    o[2] = addc(t[2], 0u);
#pragma unroll
    for (int i = 3; i < 8; i++) o[i] = t[i];
  }
  FEC_DEV static void sqr_n(u32 x[8], int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) sqr(x, x);
  }
  // a^(p-2): 255 S + 15 M (addition chain on the run structure of p-2 = 1^223 0 1^22 0000 1 0 11 0 1)
  FEC_DEV static void inv(u32 o[8], const u32 a[8]) {
    u32 x2[8], x3[8], x22[8], x44[8], t[8], u[8];
    sqr(t, a); mul(x2, t, a);
    sqr(t, x2); mul(x3, t, a);
    copy8(t, x3); sqr_n(t, 3); mul(t, t, x3);        // x6
    sqr_n(t, 3); mul(t, t, x3);                      // x9
    sqr_n(t, 2); mul(t, t, x2);                      // x11
    copy8(u, t); sqr_n(t, 11); mul(x22, t, u);       // x22
    copy8(t, x22); sqr_n(t, 22); mul(x44, t, x22);   // x44
    copy8(t, x44); sqr_n(t, 44); mul(t, t, x44);     // x88
    copy8(u, t); sqr_n(t, 88); mul(t, t, u);         // x176
    sqr_n(t, 44); mul(t, t, x44);                    // x220
    sqr_n(t, 3); mul(t, t, x3);                      // x223
    sqr_n(t, 23); mul(t, t, x22);
    sqr_n(t, 5); mul(t, t, a);
    sqr_n(t, 3); mul(t, t, x2);
    sqr_n(t, 2); mul(o, t, a);
  }
  // a^((p+1)/4) (candidate square root; p = 3 mod 4).  (p+1)/4 = 1^223 0 1^22 0000 11 00
  FEC_DEV static void sqrt_cand(u32 o[8], const u32 a[8]) {
    u32 x2[8], x3[8], x22[8], x44[8], t[8], u[8];
    sqr(t, a); mul(x2, t, a);
    sqr(t, x2); mul(x3, t, a);
    copy8(t, x3); sqr_n(t, 3); mul(t, t, x3);
    sqr_n(t, 3); mul(t, t, x3);
    sqr_n(t, 2); mul(t, t, x2);
    copy8(u, t); sqr_n(t, 11); mul(x22, t, u);
    copy8(t, x22); sqr_n(t, 22); mul(x44, t, x22);
    copy8(t, x44); sqr_n(t, 44); mul(t, t, x44);
    copy8(u, t); sqr_n(t, 88); mul(t, t, u);
    sqr_n(t, 44); mul(t, t, x44);
    sqr_n(t, 3); mul(t, t, x3);
    sqr_n(t, 23); mul(t, t, x22);
    sqr_n(t, 6); mul(t, t, x2);
    sqr_n(t, 2);
    copy8(o, t);
  }
};

// fields that provide chains of linear operations under one fold (FeK1: HAS_FUSED_LINEAR)
template <class F, class = void> struct has_fused_linear { static constexpr bool value = false; };
template <class F> struct has_fused_linear<F, decltype((void)F::HAS_FUSED_LINEAR)> { static constexpr bool value = F::HAS_FUSED_LINEAR; };

// =====================================================================================
// Curve25519 / Ed25519 base field
// =====================================================================================
struct Fe25519 {
  static constexpr int PRIME_ID = 2;
  FEC_DEV static void P(u32 p[8]) {
    p[0] = 0xFFFFFFEDu;
#pragma unroll
    for (int i = 1; i < 7; i++) p[i] = 0xFFFFFFFFu;
    p[7] = 0x7FFFFFFFu;
  }
  // The last step of both reductions: acc (< 2^256) + v 2^256 with v < 2^7.
  // Default: fold v at 2^256 (x 38) and the rare wrap once more: any value below 2^256 comes out.
  // FEC_25519_FOLD255 (the X25519 ladder, k_x25519.cu): fold at 2^255 instead (bit 255 and v, x 19): the result is below
  // 2^255 + 2^11, so the SUM of two results cannot leave 2^256 except when both are within 2^12 of 2^255 -- add_r1 / sub_r1 below
  // use that (an addition without a fold, a subtraction with one fold and no second level).
  FEC_DEV static void fold_tail(u32 o[8], u32 acc[8], u32 v) {
#ifdef FEC_25519_FOLD255
    const u32 t = (v << 1) | (acc[7] >> 31);  // < 2^8
    o[0] = add_cc(acc[0], t * 19u);
#pragma unroll
    for (int i = 1; i < 7; i++) o[i] = addc_cc(acc[i], 0u);
    o[7] = addc(acc[7] & 0x7FFFFFFFu, 0u);
#else
    acc[0] = add_cc(acc[0], v * 38u);
#pragma unroll
    for (int i = 1; i < 8; i++) acc[i] = addc_cc(acc[i], 0u);
    u32 c2 = addc(0u, 0u);
    o[0] = acc[0] + ((0u - c2) & 38u);  // wrapped value < 2^13: no further carry
#pragma unroll
    for (int i = 1; i < 8; i++) o[i] = acc[i];
#endif
  }
  FEC_DEV static bool any_lane(u32 m) {
#if defined(__CUDA_ARCH__)
    return __any_sync(__activemask(), m != 0u) != 0;  // among the lanes that are here together: the callers may be in divergent code
#else
    return m != 0u;
#endif
  }
  // a + b for a, b < 2^255 + 2^11 (results of mul / sqr under FEC_25519_FOLD255): eight additions.  The sum leaves 2^256 only when
  // both operands are within 2^12 of 2^255 -- no ladder intermediate ever is, and the warp-uniform branch keeps that case exact.
  FEC_DEV static void add_r1(u32 o[8], const u32 a[8], const u32 b[8]) {
    const u32 c = add8(o, a, b);
    if (FEC_UNLIKELY(any_lane(c))) {
      o[0] = add_cc(o[0], (0u - c) & 38u);  // the wrapped sum is below 2^13: no further carry out of the top
#pragma unroll
      for (int i = 1; i < 8; i++) o[i] = addc_cc(o[i], 0u);
    }
  }
  // a - b for b in that range (a anywhere below 2^256): a borrow means the difference is above 2^256 - 2^255 - 2^11, so "- 38"
  // cannot borrow again
  FEC_DEV static void sub_r1(u32 o[8], const u32 a[8], const u32 b[8]) {
    u32 d[8];
    const u32 bw = sub8(d, a, b);
    o[0] = sub_cc(d[0], (0u - bw) & 38u);
#pragma unroll
    for (int i = 1; i < 7; i++) o[i] = subc_cc(d[i], 0u);
    o[7] = subc(d[7], 0u);
  }
  FEC_DEV static void reduce512(u32 o[8], const u32 r[16]) {
    const u32* h = r + 8;
    u32 acc[8];
    copy8(acc, r);
    u32 ca = mad_row<4, 0>(acc, h, 38u);
    u32 od[8];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      od[2 * k] = mul_lo(h[2 * k + 1], 38u);
      od[2 * k + 1] = mul_hi(h[2 * k + 1], 38u);
    }
    acc[1] = add_cc(acc[1], od[0]);
#pragma unroll
    for (int i = 2; i < 8; i++) acc[i] = addc_cc(acc[i], od[i - 1]);
    u32 v = addc(od[7], ca);  // < 40
    fold_tail(o, acc, v);
  }
  // r2: the product is reduced from its UNMERGED column-parity accumulators (u256.cuh: mul_wide_evod).  With
  // a b = EV + OD 2^32 and h = the limbs of weight >= 2^256 (ev[8..16) + od[7..14), one 8-word chain), the even-indexed words
  // of h times 38 land on the pairs of EV's low half and the odd-indexed ones on the pairs of OD's low half (which sit one limb
  // higher): both folds are IMAD.WIDE chains with the accumulators as their own addends, and ONE merge of the two low halves
  // follows -- 16 additions instead of 23, and no moves to re-align pairs (88 MOV per X25519 ladder step were such moves).
  FEC_DEV static void mul_inl(u32 o[8], const u32 a[8], const u32 b[8]) {
#ifndef FEC_25519_EVOD_REDUCE  // set by k_x25519.cu: the ladder gains 0.8 %, the Ed25519 comb loses 1.6 % (profiles/r2z_25519_evod_ab.log)
    u32 r[16];
    mul_wide(r, a, b);
    reduce512(o, r);
#else
    u32 ev[16], od[14], h[8];
    mul_wide_evod(ev, od, a, b);
    h[0] = add_cc(ev[8], od[7]);
#pragma unroll
    for (int i = 1; i < 7; i++) h[i] = addc_cc(ev[8 + i], od[7 + i]);
    h[7] = addc(ev[15], 0u);                         // a b < 2^512: no carry out
    const u32 ca = mad_row<4, 0>(ev, h, 38u);        // ev[0..8) += {h0,h2,h4,h6} * 38
    mad_row<4, 1>(od, h + 1, 38u);                   // (od0,od1) .. (od6, 0) += {h1,h3,h5,h7} * 38: limbs 1 .. 8, cannot carry out
    u32 acc[8];
    acc[0] = ev[0];
    acc[1] = add_cc(ev[1], od[0]);
#pragma unroll
    for (int i = 2; i < 8; i++) acc[i] = addc_cc(ev[i], od[i - 1]);
    const u32 v = addc(od[7], ca);                   // limb 8: < 2^7
    fold_tail(o, acc, v);
#endif
  }
  FEC_DEV static void sqr_inl(u32 o[8], const u32 a[8]) {
    u32 r[16];
    sqr_wide(r, a);
    reduce512(o, r);
  }
  // mul / sqr: inlined, or (FEC_NOINLINE_FIELD, the big kernels) one shared subroutine per kernel image
  FEC_DEV static void mul(u32 o[8], const u32 a[8], const u32 b[8]) { field_mul<Fe25519>(o, a, b); }
  FEC_DEV static void sqr(u32 o[8], const u32 a[8]) { field_sqr<Fe25519>(o, a); }
  FEC_DEV static void add(u32 o[8], const u32 a[8], const u32 b[8]) {
    u32 s[8];
    u32 c = add8(s, a, b);
    s[0] = add_cc(s[0], (0u - c) & 38u);
#pragma unroll
    for (int i = 1; i < 8; i++) s[i] = addc_cc(s[i], 0u);
    u32 c2 = addc(0u, 0u);
    o[0] = s[0] + ((0u - c2) & 38u);
#pragma unroll
    for (int i = 1; i < 8; i++) o[i] = s[i];
  }
  FEC_DEV static void sub(u32 o[8], const u32 a[8], const u32 b[8]) {
    u32 d[8];
    u32 bw = sub8(d, a, b);
    d[0] = sub_cc(d[0], (0u - bw) & 38u);
#pragma unroll
    for (int i = 1; i < 8; i++) d[i] = subc_cc(d[i], 0u);
    u32 bw2 = subc(0u, 0u) & 1u;
    o[0] = d[0] - ((0u - bw2) & 38u);  // d >= 2^256 - 38 here: no further borrow
#pragma unroll
    for (int i = 1; i < 8; i++) o[i] = d[i];
  }
  FEC_DEV static void neg(u32 o[8], const u32 a[8]) {
    u32 z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    sub(o, z, a);
  }
  FEC_DEV static void dbl(u32 o[8], const u32 a[8]) { add(o, a, a); }
  FEC_DEV static void mul_small(u32 o[8], const u32 a[8], u32 c) {  // c < 2^26
    u32 t[8];
    u64 carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      u64 m = (u64)a[i] * c + carry;
      t[i] = (u32)m;
      carry = m >> 32;
    }
    t[0] = add_cc(t[0], (u32)carry * 38u);  // carry < 2^26 -> < 2^32
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = addc_cc(t[i], 0u);
    u32 c2 = addc(0u, 0u);
    o[0] = t[0] + ((0u - c2) & 38u);
#pragma unroll
    for (int i = 1; i < 8; i++) o[i] = t[i];
  }
  // o = a * c + b (c < 2^26) in one pass: b rides as the addend of the eight products, so the sum has one fold instead of the
  // two of mul_small + add (r2: the "AA + a24 E" of the X25519 step)
  FEC_DEV static void mul_small_add(u32 o[8], const u32 a[8], u32 c, const u32 b[8]) {
    u32 t[8];
    u64 carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      u64 m = (u64)a[i] * c + b[i];  // < 2^58 + 2^32
      m += carry;
      t[i] = (u32)m;
      carry = m >> 32;
    }
    t[0] = add_cc(t[0], (u32)carry * 38u);  // carry <= 2^26 -> < 2^32
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = addc_cc(t[i], 0u);
    u32 c2 = addc(0u, 0u);
    o[0] = t[0] + ((0u - c2) & 38u);
#pragma unroll
    for (int i = 1; i < 8; i++) o[i] = t[i];
  }
  FEC_DEV static void canon(u32 o[8], const u32 a[8]) {
    // fold bit 255 (x19), then one conditional subtract of p
    u32 t[8];
    u32 top = a[7] >> 31;
    t[0] = add_cc(a[0], top * 19u);
#pragma unroll
    for (int i = 1; i < 7; i++) t[i] = addc_cc(a[i], 0u);
    t[7] = addc(a[7] & 0x7FFFFFFFu, 0u);  // t in [0, 2^255 + 19)
    u32 u[8];
    u[0] = add_cc(t[0], 19u);
#pragma unroll
    for (int i = 1; i < 7; i++) u[i] = addc_cc(t[i], 0u);
    u[7] = addc(t[7], 0u);
    u32 ge = 0u - (u[7] >> 31);  // t + 19 >= 2^255  <=>  t >= p
    u[7] &= 0x7FFFFFFFu;
    select8(o, ge, u, t);
  }
  FEC_DEV static u32 is_zero(const u32 a[8]) {  // the three representatives below 2^256: 0, p, 2p
    const u32 o = a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7];
    const u32 mid = ~(a[1] & a[2] & a[3] & a[4] & a[5] & a[6]);
    const u32 p1 = (a[0] ^ 0xFFFFFFEDu) | mid | (a[7] ^ 0x7FFFFFFFu);
    const u32 p2 = (a[0] ^ 0xFFFFFFDAu) | mid | ~a[7];
    return ((o == 0u) | (p1 == 0u) | (p2 == 0u)) ? 0xFFFFFFFFu : 0u;
  }
  FEC_DEV static void mul2(u32 o[8], const u32 a[8]) { add(o, a, a); }
  FEC_DEV static void mul3(u32 o[8], const u32 a[8]) { u32 t[8]; add(t, a, a); add(o, t, a); }
  FEC_DEV static void mul4(u32 o[8], const u32 a[8]) { u32 t[8]; add(t, a, a); add(o, t, t); }
  FEC_DEV static void mul8(u32 o[8], const u32 a[8]) { u32 t[8]; add(t, a, a); add(t, t, t); add(o, t, t); }
  FEC_DEV static void sqr_n(u32 x[8], int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) sqr(x, x);
  }
  // byte records <-> the internal representation (see FeK1)
  FEC_DEV static void one(u32 o[8]) { o[0] = 1u; o[1] = o[2] = o[3] = o[4] = o[5] = o[6] = o[7] = 0u; }
  FEC_DEV static void from_plain(u32 o[8], const u32 a[8]) { copy8(o, a); }
  FEC_DEV static void to_plain(u32 o[8], const u32 a[8]) { canon(o, a); }
  // z^(p-2) = z^(2^255-21): 254 S + 11 M
  FEC_DEV static void inv(u32 o[8], const u32 z[8]) {
    u32 z2[8], z9[8], z11[8], z5[8], z10[8], z50[8], t[8], u[8];
    sqr(z2, z);
    copy8(t, z2); sqr_n(t, 2);
    mul(z9, z, t);
    mul(z11, z2, z9);
    sqr(t, z11);
    mul(z5, z9, t);  // 2^5 - 1
    copy8(t, z5); sqr_n(t, 5); mul(z10, t, z5);
    copy8(t, z10); sqr_n(t, 10); mul(t, t, z10);      // z20
    copy8(u, t); sqr_n(t, 20); mul(t, t, u);          // z40
    sqr_n(t, 10); mul(z50, t, z10);                   // z50
    copy8(t, z50); sqr_n(t, 50); mul(t, t, z50);      // z100
    copy8(u, t); sqr_n(t, 100); mul(t, t, u);         // z200
    sqr_n(t, 50); mul(t, t, z50);                     // z250
    sqr_n(t, 5); mul(o, t, z11);
  }
};

// =====================================================================================
// NIST P-256 base field, Montgomery domain: an element x is held as x * 2^256 mod p, canonical in [0,p)
//
// p = 2^256 - 2^224 + 2^192 + 2^96 - 1 is -1 modulo 2^96, so the Montgomery quotient digit of a 64-bit round is the
// low 64 bits themselves and q * p is made of shifted copies of q: the reduction has NO multiplications and about half
// the additions of the FIPS 186-4 Solinas sums on 32-bit words (which made P-256 ALU-bound).  Byte records enter with
// from_plain (one multiplication by 2^512 mod p) and leave with to_plain (one reduction); everything between --
// formulas, tables, inversion, square root, comparisons -- is domain-agnostic.
// =====================================================================================
struct FeP256 {
  static constexpr int PRIME_ID = 1;
  FEC_DEV static void P(u32 p[8]) {
    p[0] = 0xFFFFFFFFu; p[1] = 0xFFFFFFFFu; p[2] = 0xFFFFFFFFu; p[3] = 0u;
    p[4] = 0u; p[5] = 0u; p[6] = 1u; p[7] = 0xFFFFFFFFu;
  }
  FEC_DEV static void one(u32 o[8]) {  // 2^256 mod p
    o[0] = 1u; o[1] = 0u; o[2] = 0u; o[3] = 0xFFFFFFFFu; o[4] = 0xFFFFFFFFu; o[5] = 0xFFFFFFFFu; o[6] = 0xFFFFFFFEu; o[7] = 0u;
  }
  FEC_DEV static void RR(u32 o[8]) {  // 2^512 mod p
    o[0] = 3u; o[1] = 0u; o[2] = 0xFFFFFFFFu; o[3] = 0xFFFFFFFBu; o[4] = 0xFFFFFFFEu; o[5] = 0xFFFFFFFFu; o[6] = 0xFFFFFFFDu; o[7] = 4u;
  }
  // o = c / 2^256 mod p for c < p * 2^256, canonical.  Four rounds on 64 bits: with q = the two lowest live words,
  // c += q p = q (2^96 - 1) + (q (2^64 - 2^32 + 1)) 2^192 clears them; the carry out of a round's chain (at word
  // o + 10) is folded into the next round's m (its word 2), so every chain is 7 words long.
  // `dep` is a value that is always zero but that ptxas cannot prove to be: it is or-ed into the first quotient digit to
  // keep the reduction behind whatever `dep` was computed from.  Without it ptxas starts the rounds while the
  // multiplication's carry chains are still open, runs out of predicate registers and moves carries through register
  // bits (55 of the 221 instructions of the out-of-line multiplication were such LOP3s).
  FEC_DEV static void reduce512(u32 o[8], const u32 c[16], u32 dep = 0u) {
    u32 t[16];
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = c[i];
    t[0] |= dep;
    u32 cy = 0u;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int b = 2 * r;
      const u32 q0 = t[b], q1 = t[b + 1];
      // m = q + ((q + cy) << 64) - (q << 32)   (mod 2^128; the true value is below 2^128)
      u32 s2, s3;
      if (r == 0) { s2 = q0; s3 = q1; } else { s2 = add_cc(q0, cy); s3 = addc(q1, 0u); }
      const u32 m1 = sub_cc(q1, q0);
      const u32 m2 = subc_cc(s2, q1);
      const u32 m3 = subc(s3, 0u);
      t[b + 3] = add_cc(t[b + 3], q0);
      t[b + 4] = addc_cc(t[b + 4], q1);
      t[b + 5] = addc_cc(t[b + 5], 0u);
      t[b + 6] = addc_cc(t[b + 6], q0);
      t[b + 7] = addc_cc(t[b + 7], m1);
      t[b + 8] = addc_cc(t[b + 8], m2);
      t[b + 9] = addc_cc(t[b + 9], m3);
      cy = addc(0u, 0u);
    }
    red_final(o, t, cy);
  }
  // The last step.  Canonical form (FEC_P256_CANONICAL, and always for to_plain / canon): (cy, t[8..16)) < 2p for operands below
  // p, one compare-and-select.  WEAK form (r2, the default): operands are anywhere below 2^256, so (cy, t[8..16)) < 2^256 + p, and
  // subtracting p exactly when cy is set leaves a value below 2^256 -- a masked subtraction whose borrow cancels cy: 10
  // instructions instead of 18, on the pipe that bounds this field.  (The four rounds are exact for any c < 2^512.)
  FEC_DEV static void red_final_canon(u32 o[8], const u32 t[16], u32 cy) {
    u32 p[8], d[8];
    P(p);
    const u32 bw = sub8(d, t + 8, p);
    select8(o, (0u - cy) | (bw - 1u), d, t + 8);
  }
  FEC_DEV static void sub_p_masked(u32 o[8], const u32 a[8], u32 m) {  // a - (p & m) mod 2^256, p = {~0,~0,~0,0,0,0,1,~0}
    o[0] = sub_cc(a[0], m);
    o[1] = subc_cc(a[1], m);
    o[2] = subc_cc(a[2], m);
    o[3] = subc_cc(a[3], 0u);
    o[4] = subc_cc(a[4], 0u);
    o[5] = subc_cc(a[5], 0u);
    o[6] = subc_cc(a[6], m & 1u);
    o[7] = subc_cc(a[7], m);
  }
  FEC_DEV static void add_p_masked(u32 o[8], const u32 a[8], u32 m) {  // a + (p & m) mod 2^256
    o[0] = add_cc(a[0], m);
    o[1] = addc_cc(a[1], m);
    o[2] = addc_cc(a[2], m);
    o[3] = addc_cc(a[3], 0u);
    o[4] = addc_cc(a[4], 0u);
    o[5] = addc_cc(a[5], 0u);
    o[6] = addc_cc(a[6], m & 1u);
    o[7] = addc_cc(a[7], m);
  }
  FEC_DEV static bool any_lane(u32 m) {
#if defined(__CUDA_ARCH__)
    return __any_sync(__activemask(), m != 0u) != 0;  // among the lanes that are here together: the callers may be in divergent code
#else
    return m != 0u;
#endif
  }
  // reduce512 in stages (see mul2_inl): one 64-bit round, and the conditional subtraction
  template <int R>
  FEC_DEV static void red_round(u32 t[16], u32& cy) {
    const int b = 2 * R;
    const u32 q0 = t[b], q1 = t[b + 1];
    u32 s2, s3;
    if (R == 0) { s2 = q0; s3 = q1; } else { s2 = add_cc(q0, cy); s3 = addc(q1, 0u); }
    const u32 m1 = sub_cc(q1, q0);
    const u32 m2 = subc_cc(s2, q1);
    const u32 m3 = subc(s3, 0u);
    t[b + 3] = add_cc(t[b + 3], q0);
    t[b + 4] = addc_cc(t[b + 4], q1);
    t[b + 5] = addc_cc(t[b + 5], 0u);
    t[b + 6] = addc_cc(t[b + 6], q0);
    t[b + 7] = addc_cc(t[b + 7], m1);
    t[b + 8] = addc_cc(t[b + 8], m2);
    t[b + 9] = addc_cc(t[b + 9], m3);
    cy = addc(0u, 0u);
  }
  FEC_DEV static void red_final(u32 o[8], const u32 t[16], u32 cy) {
#ifdef FEC_P256_CANONICAL
    red_final_canon(o, t, cy);
#else
    sub_p_masked(o, t + 8, 0u - cy);
#endif
  }
  // Two independent multiplications, INTERLEAVED at carry-chain granularity (FEC_VM_DUAL): all rows of
  // the first product, then the rows of the second one between the merge, the four reduction rounds and the final
  // subtraction of the first (between two chains the carry flag is dead, so the order is legal in PTX), then the tail of the
  // second.  A multiplication is 64 IMAD.WIDE (multiply pipe) and ~88 dependent additions (ALU pipe): alone, a warp uses
  // the two pipes one after the other.  Without the tokens below ptxas regroups the stages (all 128 IMAD.WIDE first).
  FEC_DEV static void mul2_inl(u32 o0[8], const u32 a0[8], const u32 b0[8], u32 o1[8], const u32 a1[8], const u32 b1[8]) {
    ProdSt A, B;
    u32 t[16], tc, cy = 0u;
    const u32 z0 = opaque_zero_always(a1[0]);
    prod_row<0>(A, a0, b0); prod_row<1>(A, a0, b0); prod_row<2>(A, a0, b0); prod_row<3>(A, a0, b0);
    prod_row<4>(A, a0, b0); prod_row<5>(A, a0, b0); prod_row<6>(A, a0, b0); prod_row<7>(A, a0, b0);
    // row k + 2 of the second product waits for tail stage k of the first one: its multiplier word is or-ed with an
    // always-zero value computed from the stage's last result (the carry out of the merge; `carry & z0` afterwards)
    u32 bb[8];
    copy8(bb, b1);
    bb[0] |= A.ev[15] & z0;   // rows 0 and 1 behind the first product's last row: they run next to its merge
    prod_row<0>(B, a1, bb);
    prod_merge(t, tc, A);
    t[0] |= tc;
    bb[1] |= A.od[14] & z0;
    prod_row<1>(B, a1, bb);
    red_round<0>(t, cy);
    bb[2] |= tc;
    prod_row<2>(B, a1, bb);
    red_round<1>(t, cy);
    bb[3] |= cy & z0;
    prod_row<3>(B, a1, bb);
    red_round<2>(t, cy);
    bb[4] |= cy & z0;
    prod_row<4>(B, a1, bb);
    red_round<3>(t, cy);
    bb[5] |= cy & z0;
    prod_row<5>(B, a1, bb);
    red_final(o0, t, cy);
    bb[6] |= cy & z0;
    prod_row<6>(B, a1, bb);
    bb[7] |= o0[7] & z0;
    prod_row<7>(B, a1, bb);
    prod_merge(t, tc, B);
    reduce512(o1, t, tc);
  }
  // always-zero values for reduce512 (device code only).  top_is_ones: the top word of a product of two residues below p
  // is at most 0xFFFFFFFE (p^2 < 2^512 - 2^481).  top_overflows: the two top partial words of the schoolbook product
  // add up to the top word of a number below 2^512.
  FEC_DEV static u32 top_is_ones(u32 c15) {
#if defined(__CUDA_ARCH__)
    u32 dep;
    asm volatile("{ .reg .pred pz; setp.eq.u32 pz, %1, 0xffffffff; selp.u32 %0, 1, 0, pz; }" : "=r"(dep) : "r"(c15));
    return dep;
#else
    return 0u;
#endif
  }
  FEC_DEV static u32 top_overflows(u32 te, u32 to) {
#if defined(__CUDA_ARCH__)
    u32 dep;
    asm volatile("{ .reg .u32 s; add.cc.u32 s, %1, %2; addc.u32 %0, 0, 0; }" : "=r"(dep) : "r"(te), "r"(to));
    return dep;
#else
    return 0u;
#endif
  }
#ifndef FEC_P256_FENCE_MUL
#define FEC_P256_FENCE_MUL 3  // 0 none, 1 top word all ones, 2 behind the product rows (merge and reduction may overlap), 3 carry out of the merge
#endif
#ifndef FEC_P256_FENCE_SQR
#define FEC_P256_FENCE_SQR 3  // 0 none, 1 top word all ones, 3 carry out of the last chain
#endif
  // Variants 2 and 3 are zero for EVERY input; variant 1 relies on operands below p (all field routines return them).
  FEC_DEV static void mul_inl(u32 o[8], const u32 a[8], const u32 b[8]) {
    u32 r[16], te, to, tc;
#ifdef FEC_KARATSUBA
    mul_wide(r, a, b);
    te = to = tc = 0u;
#else
    mul_wide_school_tok<FEC_P256_FENCE_MUL == 3>(r, a, b, te, to, tc);
#endif
    reduce512(o, r, FEC_P256_FENCE_MUL == 1 ? top_is_ones(r[15]) : FEC_P256_FENCE_MUL == 2 ? top_overflows(te, to) : FEC_P256_FENCE_MUL == 3 ? tc : 0u);
  }
  FEC_DEV static void sqr_inl(u32 o[8], const u32 a[8]) {
    u32 r[16], tc;
    sqr_wide_tok<FEC_P256_FENCE_SQR == 3>(r, a, tc);
    reduce512(o, r, FEC_P256_FENCE_SQR == 1 ? top_is_ones(r[15]) : FEC_P256_FENCE_SQR == 3 ? tc : 0u);
  }
  // mul / sqr: inlined, or (FEC_NOINLINE_FIELD, the big kernels) one shared subroutine per kernel image
  FEC_DEV static void mul(u32 o[8], const u32 a[8], const u32 b[8]) { field_mul<FeP256>(o, a, b); }
  FEC_DEV static void sqr(u32 o[8], const u32 a[8]) { field_sqr<FeP256>(o, a); }
#ifdef FEC_P256_CANONICAL
  FEC_DEV static void add(u32 o[8], const u32 a[8], const u32 b[8]) {
    u32 s[8], d[8], p[8];
    P(p);
    u32 c = add8(s, a, b);
    u32 bw = sub8(d, s, p);
    u32 take = (0u - c) | (bw - 1u);  // carried out, or s >= p
    select8(o, take, d, s);
  }
  FEC_DEV static void sub(u32 o[8], const u32 a[8], const u32 b[8]) {
    u32 d[8];
    u32 bw = sub8(d, a, b);
    add_p_masked(o, d, 0u - bw);
  }
#else
  // Weak form: operands and results anywhere below 2^256.  A carry out of a + b takes one p away (the borrow of that subtraction
  // cancels the carry); the sum is still above 2^256 only when a + b >= 2^256 + p, i.e. both operands above p -- the top 2^-32 of the
  // range -- and that second correction sits under a warp-uniform branch.  The same for a borrow of a - b and b - a > p.
  FEC_DEV static void add(u32 o[8], const u32 a[8], const u32 b[8]) {
    u32 s[8];
    const u32 c = add8(s, a, b);
    sub_p_masked(o, s, 0u - c);
    const u32 bw = subc(0u, 0u) & 1u;
    const u32 still = c & (bw ^ 1u);
    if (FEC_UNLIKELY(any_lane(still))) {
      u32 t[8];
      copy8(t, o);
      sub_p_masked(o, t, 0u - still);
    }
  }
  FEC_DEV static void sub(u32 o[8], const u32 a[8], const u32 b[8]) {
    u32 d[8];
    const u32 bw = sub8(d, a, b);
    add_p_masked(o, d, 0u - bw);
    const u32 cy = addc(0u, 0u);
    const u32 still = bw & (cy ^ 1u);
    if (FEC_UNLIKELY(any_lane(still))) {
      u32 t[8];
      copy8(t, o);
      add_p_masked(o, t, 0u - still);
    }
  }
#endif
  FEC_DEV static void neg(u32 o[8], const u32 a[8]) {
    u32 z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    sub(o, z, a);
  }
  FEC_DEV static void dbl(u32 o[8], const u32 a[8]) { add(o, a, a); }
  FEC_DEV static void mul_small(u32 o[8], const u32 a[8], u32 c) {  // c in {2,3,4,8}: by additions
    u32 t[8];
    copy8(t, a);
    u32 acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (u32 bit = 0; bit < 5; bit++) {
      if ((c >> bit) & 1u) add(acc, acc, t);  // c is uniform: not data-dependent
      add(t, t, t);
    }
    copy8(o, acc);
  }
#ifdef FEC_P256_CANONICAL
  FEC_DEV static void canon(u32 o[8], const u32 a[8]) { copy8(o, a); }
  FEC_DEV static u32 is_zero(const u32 a[8]) { return is_zero8(a); }
  FEC_DEV static u32 maybe_zero(const u32 a[8]) { return a[7] == 0u ? 1u : 0u; }  // necessary condition (canonical form)
#else
  FEC_DEV static void canon(u32 o[8], const u32 a[8]) {  // the representative below p
    u32 p[8], d[8];
    P(p);
    const u32 bw = sub8(d, a, p);
    select8(o, bw - 1u, d, a);
  }
  FEC_DEV static u32 is_zero(const u32 a[8]) {  // mask; the representatives of zero below 2^256 are 0 and p
    const u32 z = a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7];
    const u32 q = ~(a[0] & a[1] & a[2] & a[7]) | a[3] | a[4] | a[5] | (a[6] ^ 1u);
    return ((z == 0u) | (q == 0u)) ? 0xFFFFFFFFu : 0u;
  }
  FEC_DEV static u32 maybe_zero(const u32 a[8]) { return (a[7] + 1u <= 1u) ? 1u : 0u; }  // necessary: top word of 0 or of p
#endif
  FEC_DEV static void from_plain(u32 o[8], const u32 a[8]) {  // a < p as an integer -> a * 2^256 mod p
    u32 rr[8];
    RR(rr);
    mul(o, a, rr);
  }
  FEC_DEV static void to_plain(u32 o[8], const u32 a[8]) {  // back to the integer in [0,p)
    u32 t[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { t[i] = a[i]; t[8 + i] = 0u; }
    u32 cy = 0u;
    red_round<0>(t, cy); red_round<1>(t, cy); red_round<2>(t, cy); red_round<3>(t, cy);
    red_final_canon(o, t, cy);  // a / 2^256 < p + 1: canonical after one compare-and-select, whatever form a came in
  }
  FEC_DEV static void mul2(u32 o[8], const u32 a[8]) { add(o, a, a); }
  FEC_DEV static void mul3(u32 o[8], const u32 a[8]) { u32 t[8]; add(t, a, a); add(o, t, a); }
  FEC_DEV static void mul4(u32 o[8], const u32 a[8]) { u32 t[8]; add(t, a, a); add(o, t, t); }
  FEC_DEV static void mul8(u32 o[8], const u32 a[8]) { u32 t[8]; add(t, a, a); add(t, t, t); add(o, t, t); }
  FEC_DEV static u32 in_range(const u32 a[8]) {
    u32 p[8];
    P(p);
    return ~geq8(a, p);
  }
  // a / 2: (a + (a odd ? p : 0)) >> 1; any representative in (a + p is even for either), below 2^256 out
  FEC_DEV static void half(u32 o[8], const u32 a[8]) {
    const u32 m = 0u - (a[0] & 1u);
    u32 t[8];
    t[0] = add_cc(a[0], m);
    t[1] = addc_cc(a[1], m);
    t[2] = addc_cc(a[2], m);
    t[3] = addc_cc(a[3], 0u);
    t[4] = addc_cc(a[4], 0u);
    t[5] = addc_cc(a[5], 0u);
    t[6] = addc_cc(a[6], m & 1u);
    t[7] = addc_cc(a[7], m);
    const u32 c = addc(0u, 0u);
#pragma unroll
    for (int i = 0; i < 7; i++) o[i] = funnel_l(t[i], t[i + 1], 31);
    o[7] = funnel_l(t[7], c, 31);
  }
  FEC_DEV static void sqr_n(u32 x[8], int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) sqr(x, x);
  }
  // a^(p-2), p-2 = 1^32 0^31 1 0^96 1^94 0 1 : 255 S + 12 M
  FEC_DEV static void inv(u32 o[8], const u32 a[8]) {
    u32 x2[8], x3[8], x30[8], x32[8], t[8], u[8];
    sqr(t, a); mul(x2, t, a);
    sqr(t, x2); mul(x3, t, a);
    copy8(t, x3); sqr_n(t, 3); mul(t, t, x3);      // x6
    copy8(u, t); sqr_n(t, 6); mul(t, t, u);        // x12
    sqr_n(t, 3); mul(t, t, x3);                    // x15
    copy8(u, t); sqr_n(t, 15); mul(x30, t, u);     // x30
    copy8(t, x30); sqr_n(t, 2); mul(x32, t, x2);   // x32
    copy8(t, x32); sqr_n(t, 32); mul(t, t, a);
    sqr_n(t, 128); mul(t, t, x32);
    sqr_n(t, 32); mul(t, t, x32);
    sqr_n(t, 30); mul(t, t, x30);
    sqr_n(t, 2); mul(o, t, a);
  }
  // a^((p+1)/4), (p+1)/4 = 2^254 - 2^222 + 2^190 + 2^94 = 1^32 0^31 1 0^31 1 0^94
  FEC_DEV static void sqrt_cand(u32 o[8], const u32 a[8]) {
    u32 x2[8], x4[8], x8[8], x16[8], t[8];
    sqr(t, a); mul(x2, t, a);
    copy8(t, x2); sqr_n(t, 2); mul(x4, t, x2);
    copy8(t, x4); sqr_n(t, 4); mul(x8, t, x4);
    copy8(t, x8); sqr_n(t, 8); mul(x16, t, x8);
    copy8(t, x16); sqr_n(t, 16); mul(t, t, x16);   // x32
    sqr_n(t, 32); mul(t, t, a);
    sqr_n(t, 96); mul(t, t, a);
    sqr_n(t, 94);
    copy8(o, t);
  }
};

#if defined(__CUDACC__) && defined(FEC_NOINLINE_FIELD)
struct fe8 { u32 v[8]; };
template <class F> __device__ __noinline__ fe8 fe_mul_ni(fe8 a, fe8 b) { fe8 r; F::mul_inl(r.v, a.v, b.v); return r; }
template <class F> __device__ __noinline__ fe8 fe_sqr_ni(fe8 a) { fe8 r; F::sqr_inl(r.v, a.v); return r; }
template <class F> FEC_DEV void field_mul(u32 o[8], const u32 a[8], const u32 b[8]) {
  fe8 x, y;
  copy8(x.v, a); copy8(y.v, b);
  const fe8 r = fe_mul_ni<F>(x, y);
  copy8(o, r.v);
}
template <class F> FEC_DEV void field_sqr(u32 o[8], const u32 a[8]) {
#ifdef FEC_INLINE_SQR
  F::sqr_inl(o, a);
#else
  fe8 x;
  copy8(x.v, a);
  const fe8 r = fe_sqr_ni<F>(x);
  copy8(o, r.v);
#endif
}
#else
template <class F> FEC_DEV void field_mul(u32 o[8], const u32 a[8], const u32 b[8]) { F::mul_inl(o, a, b); }
template <class F> FEC_DEV void field_sqr(u32 o[8], const u32 a[8]) { F::sqr_inl(o, a); }
#endif

}  // namespace fec
