// lanes.cuh -- what ONE lane (one thread) does on the hot path, independent of where its table
// and its batch inversion live.  Everything here is branch-free with respect to secret data:
// loops have fixed trip counts, selections are masks, table entries are fetched by index (the
// address depends on the digit, the control flow does not).
//
//   scalar_prepare / next_digit   regular signed odd-digit recoding of k (all 64 digits non-zero)
//   wmul_var_core                 variable-base k*P, short Weierstrass, 4-bit signed window
//                                 (batched replacement of Curve::multiply,
//                                  forge-ec-curves/src/secp256k1.rs:2635-2692, p256.rs:2120-2156)
//   wmul_fixed_core               fixed-base k*G from a 16 x 32768 table of odd multiples (16-bit signed windows)
//   x25519_core                   RFC 7748 ladder (curve25519.rs:1624-1716)
//   ed25519_expand_seed / ed_comb_core / ed_encode
//                                 RFC 8032 5.1.5 key generation (forge-ec-signature/src/eddsa.rs:450-509,
//                                 ed25519.rs:2062-2097, 1793-1811)
//   sec1_encode                   PointAffine::to_bytes / CompressedPoint::from_affine
//                                 (secp256k1.rs:875-894, forge-ec-encoding/src/point.rs:39-67, 186-211)
//
// Recoding.  For odd k' < 2^256 let K = 2^255 + (k'-1)/2 and u_i its i-th nibble.  Then
// d_i = 2 u_i - 15 is odd, |d_i| <= 15 and sum d_i 16^i = 2K - (2^256 - 1) = k'.  The top digit
// is positive.  An even scalar is replaced by n - k (odd, because n is odd) and the result is
// negated at the end; k = 0 and k >= n are replaced by 1 and reported through the status byte.
#pragma once
#include "curves.cuh"
#include "sha512_k.inc"

namespace fec {

enum { LANE_OK = 0, LANE_SCALAR_RANGE = 1, LANE_POINT_INVALID = 2, LANE_INFINITY = 3 };

#ifdef __CUDA_ARCH__
FEC_DEV bool warp_any(u32 m) { return __any_sync(0xFFFFFFFFu, m != 0u) != 0; }
#else
FEC_DEV bool warp_any(u32 m) { return m != 0u; }
#endif

FEC_DEV void set_small(u32 r[8], u32 v) {
  r[0] = v;
#pragma unroll
  for (int i = 1; i < 8; i++) r[i] = 0u;
}
FEC_DEV void shl4_256(u32 K[8]) {
#pragma unroll
  for (int i = 7; i >= 1; i--) K[i] = funnel_l(K[i - 1], K[i], 4);
  K[0] <<= 4;
}
// K = 2^255 | (k >> 1)
FEC_DEV void recode_reg(u32 K[8], const u32 k[8]) {
#pragma unroll
  for (int i = 0; i < 7; i++) K[i] = funnel_l(k[i], k[i + 1], 31);
  K[7] = (k[7] >> 1) | 0x80000000u;
}
// pops the most significant remaining digit: idx in 0..7 (|d| = 2 idx + 1), sgn = all-ones if d < 0
FEC_DEV void next_digit(u32 K[8], u32& idx, u32& sgn) {
  u32 u = K[7] >> 28;
  shl4_256(K);
  sgn = (u >> 3) - 1u;
  idx = (u ^ sgn) & 7u;
}

// Weierstrass scalar: k (limbs, from big-endian bytes) -> recoding register, negate mask, status
template <class C>
FEC_DEV void scalar_prepare(u32 K[8], u32& negm, u32& st, const u32 k[8]) {
  u32 n[8], one[8], kk[8], nk[8];
  C::N(n);
  set_small(one, 1u);
  u32 bad = geq8(k, n);
  u32 zero = is_zero8(k);
  select8(kk, bad | zero, one, k);
  sub8(nk, n, kk);
  u32 even = (kk[0] & 1u) - 1u;
  select8(kk, even, nk, kk);
  negm = even;
  recode_reg(K, kk);
  st = (bad & LANE_SCALAR_RANGE) | (~bad & zero & LANE_INFINITY);
}

// Q <- Q + (x2,y2), (x2,y2) affine, as jac_madd but also returns zr = Z3/Z1 = 2H (table building)
template <class C>
FEC_DEV void jac_madd_zr(Jac& q, const u32 x2[8], const u32 y2[8], u32 zr[8]) {
  typedef typename C::F F;
  u32 Z1Z1[8], U2[8], S2[8], H[8], HH[8], I[8], J[8], r[8], V[8], t[8];
  F::sqr(Z1Z1, q.Z);
  F::mul(U2, x2, Z1Z1);
  F::mul(S2, q.Z, Z1Z1);
  F::mul(S2, y2, S2);
  F::sub(H, U2, q.X);
  F::sqr(HH, H);
  F::mul4(I, HH);
  F::mul(J, H, I);
  F::sub(r, S2, q.Y);
  F::mul2(r, r);
  F::mul(V, q.X, I);
  F::mul2(zr, H);
  F::mul(q.Z, q.Z, zr);              // Z3 = 2 Z1 H
  F::sqr(q.X, r);
  F::sub(q.X, q.X, J);
  F::sub(q.X, q.X, V);
  F::sub(q.X, q.X, V);
  F::sub(t, V, q.X);
  F::mul(t, r, t);
  F::mul(J, q.Y, J);
  F::mul2(J, J);
  F::sub(q.Y, t, J);
}

// Co-Z addition with update (Meloni; Goundar-Joye-Miyaji ZADDU), for the tables of odd multiples: R = (rx, ry) and P = (px, py)
// share one (implicit) Z.  R <- R + P, P <- the same point P rescaled to the Z of the new R, zr = Z_new / Z_old = px - rx.
// 4M + 2S and 7 subtractions -- against 8M + 3S and 10 linear operations for the mixed addition that also tracks Z
// (jac_madd_zr above): the table builders only need the RATIOS of consecutive Z (r2: -24 % of the secp256k1 prepare kernel).
// No exceptional case for a point of prime order > 15: R = (2j-1) P and P = 2 P' never share an x.
template <class C>
FEC_DEV void coz_addu(u32 rx[8], u32 ry[8], u32 px[8], u32 py[8], u32 zr[8]) {
  typedef typename C::F F;
  u32 Cc[8], W1[8], W2[8], D[8], A1[8], t[8];
  F::sub(zr, px, rx);                // X1 - X2
  F::sqr(Cc, zr);
  F::mul(W1, px, Cc);
  F::mul(W2, rx, Cc);
  F::sub(t, py, ry);                 // Y1 - Y2
  F::sqr(D, t);
  F::sub(Cc, W1, W2);
  F::mul(A1, py, Cc);                // Y1 (W1 - W2)
  F::sub(D, D, W1);
  F::sub(rx, D, W2);                 // X3 = D - W1 - W2
  F::sub(Cc, W1, rx);
  F::mul(Cc, t, Cc);
  F::sub(ry, Cc, A1);                // Y3 = (Y1 - Y2)(W1 - X3) - A1
  copy8(px, W1);
  copy8(py, A1);
}

// mixed addition with the two exceptional cases repaired under a warp-uniform branch:
// Q == T -> 2T, Q == -T -> infinity.  (They need k = +-d (mod n) for a partial scalar, which the
// recoding only produces for a handful of scalars next to n; see DESIGN.md.)
template <class C>
FEC_DEV void jac_madd_safe(Jac& q, const u32 tx[8], const u32 ty[8]) {
  const u32 flags = jac_madd_lean<C>(q, tx, ty);
  if (FEC_UNLIKELY(warp_any(flags))) jac_madd_fix<typename C::Cold>(q, tx, ty, flags);
}

// the same for the addend (tx, neg_y ? -ty : ty), zero tests behind the filter (curves.cuh: jac_madd_lean_signed)
template <class C>
FEC_DEV void jac_madd_safe_signed(Jac& q, const u32 tx[8], const u32 ty[8], u32 neg_y) {
#ifndef FEC_MADD_EXACT_FLAGS
  const u32 flags = jac_madd_lean_signed<C>(q, tx, ty, neg_y, [](u32 m) { return warp_any(m); });
  if (FEC_UNLIKELY(warp_any(flags))) {
    u32 ny[8], sy[8];
    C::F::neg(ny, ty);
    select8(sy, neg_y, ny, ty);
    jac_madd_fix<typename C::Cold>(q, tx, sy, flags);
  }
#else
  u32 ny[8], sy[8];
  C::F::neg(ny, ty);
  select8(sy, neg_y, ny, ty);
  jac_madd_safe<C>(q, tx, sy);
#endif
}

// P <- P + Q (Jacobian + Jacobian) including P == Q, which is repaired under a warp-uniform branch
template <class C>
FEC_DEV void jac_add_safe(Jac& p, const Jac& q) {
  const u32 same = jac_add<C>(p, q);
  if (warp_any(same)) {
    Jac d = p;
    jac_dbl<C>(d);
    select8(p.X, same, d.X, p.X); select8(p.Y, same, d.Y, p.Y); select8(p.Z, same, d.Z, p.Z);
  }
}

// -------------------------------------------------------------------------------------
// variable base.  Tab: per-lane table of 8 affine points with load(idx,x,y) / store(idx,x,y).
// Inv: simultaneous inversion across the lanes that call it together: inv(zinv, z), z != 0.
// Cost: table 1 dbl + 7 madd + 7 x (4M+1S) + shared inversion; loop 252 dbl + 63 madd.
// -------------------------------------------------------------------------------------
// the table: affine odd multiples P, 3P, .., 15P (entry j = (2j+1) P), ONE shared inversion
// the table in two halves, so that a caller may put ONE inversion between the halves of several tables (k_p3_prep):
// build: entries 1..7 on the curve isomorphic by zt (entry 0 is P itself), zt = their common true Z, zr = the co-Z ratios
template <class C, class Tab>
FEC_DEV void wmul_var_table_build(Tab& tab, const u32 px[8], const u32 py[8], u32 zt[8], u32 zr[7][8]) {
  typedef typename C::F F;
  // odd multiples P, 3P, .., 15P on the curve isomorphic by C = Z(2P), where 2P is affine
  Jac D, A;
  u32 C2[8], C3[8];
  copy8(D.X, px); copy8(D.Y, py); F::one(D.Z);
  jac_dbl<C>(D);
  F::sqr(C2, D.Z);
  F::mul(C3, C2, D.Z);
  F::mul(A.X, px, C2);
  F::mul(A.Y, py, C3);
  tab.store(0, px, py);
  copy8(zt, D.Z);
#pragma unroll 1
  for (int j = 1; j < 8; j++) {  // A = (2j+1) P and D = 2P kept at one common Z (co-Z additions)
    coz_addu<C>(A.X, A.Y, D.X, D.Y, zr[j - 1]);
    tab.store(j, A.X, A.Y);
    F::mul(zt, zt, zr[j - 1]);   // true Z of the newest entry: Z(2P) times the ratios so far
  }
}
// finish: zi = 1 / zt; every entry back to the curve itself (affine)
template <class C, class Tab>
FEC_DEV void wmul_var_table_finish(Tab& tab, const u32 zi_in[8], const u32 zr[7][8]) {
  typedef typename C::F F;
  u32 zi[8];
  copy8(zi, zi_in);
#pragma unroll 1
  for (int j = 7; j >= 1; j--) {
    u32 x[8], y[8], z2[8];
    tab.load(j, x, y);
    F::sqr(z2, zi);
    F::mul(x, x, z2);
    F::mul(z2, z2, zi);
    F::mul(y, y, z2);
    tab.store(j, x, y);
    F::mul(zi, zi, zr[j - 1]);
  }
}
template <class C, class Tab, class Inv>
FEC_DEV void wmul_var_table(Tab& tab, const u32 px[8], const u32 py[8], Inv& inv) {
  u32 zt[8], zi[8];
  u32 zr[7][8];
  wmul_var_table_build<C>(tab, px, py, zt, zr);
  inv(zi, zt);
  wmul_var_table_finish<C>(tab, zi, zr);
}
template <class C, class Tab, class Inv>
FEC_DEV void wmul_var_core(Jac& Q, u32 K[8], const u32 px[8], const u32 py[8], Tab& tab, Inv& inv) {
  typedef typename C::F F;
  wmul_var_table<C>(tab, px, py, inv);
  u32 idx, sgn;
  next_digit(K, idx, sgn);  // top digit: always positive
  tab.load(idx, Q.X, Q.Y);
  F::one(Q.Z);
#pragma unroll 1
  for (int i = 0; i < 63; i++) {
#pragma unroll 1
    for (int d = 0; d < 4; d++) jac_dbl<C>(Q);
    u32 tx[8], ty[8], ny[8];
    next_digit(K, idx, sgn);
    tab.load(idx, tx, ty);
    F::neg(ny, ty);
    select8(ty, sgn, ny, ty);
    jac_madd_safe<C>(Q, tx, ty);
  }
}

// -------------------------------------------------------------------------------------
// fixed base.  The same regular recoding with W-bit digits: d_i = 2 u_i - (2^W - 1), u_i the i-th W-bit chunk of
// K = 2^255 + (k'-1)/2, so every digit is odd, |d_i| < 2^W, and the table holds the 2^(W-1) odd multiples of
// 2^(W i) G per window.  With W = 16 that is 16 windows x 32768 affine points = 32 MiB per curve: it lives in HBM
// and is served by the 126 MB L2, and a scalar multiplication is 15 mixed additions instead of the 63 of a 4-bit
// comb.  comb(i, idx, x, y) returns (2 idx + 1) 2^(W i) G.
// -------------------------------------------------------------------------------------
constexpr int COMB_W = 16;
constexpr int COMB_WINDOWS = 256 / COMB_W;
constexpr int COMB_ENTRIES = 1 << (COMB_W - 1);
template <int W>
FEC_DEV void next_digit_w(u32 K[8], u32& idx, u32& sgn) {
  const u32 u = K[7] >> (32 - W);
#pragma unroll
  for (int i = 7; i >= 1; i--) K[i] = (W == 32) ? K[i - 1] : funnel_l(K[i - 1], K[i], W & 31);
  K[0] = (W == 32) ? 0u : (K[0] << (W & 31));
  sgn = (u >> (W - 1)) - 1u;
  idx = (u ^ sgn) & ((1u << (W - 1)) - 1u);
}
template <class C, class Comb, int W = COMB_W>
FEC_DEV void wmul_fixed_core(Jac& Q, u32 K[8], const Comb& comb) {
  typedef typename C::F F;
  constexpr int WINDOWS = 256 / W;
  u32 idx, sgn, nx[8], ny2[8];
  next_digit_w<W>(K, idx, sgn);  // top digit: always positive
  comb(WINDOWS - 1, idx, Q.X, Q.Y);
  F::one(Q.Z);
  next_digit_w<W>(K, idx, sgn);
  comb(WINDOWS - 2, idx, nx, ny2);
#pragma unroll 1
  for (int i = WINDOWS - 2; i >= 0; i--) {
    u32 tx[8], ty[8];
    const u32 sg = sgn;
    copy8(tx, nx); copy8(ty, ny2);
    if (i > 0) {  // fetch the next addend while this addition runs
      next_digit_w<W>(K, idx, sgn);
      comb(i - 1, idx, nx, ny2);
    }
    jac_madd_safe_signed<C>(Q, tx, ty, sg);
  }
}
// The constant-time form of the same (FORGE_EC_CT_TABLE=1): 4-bit windows, 64 x 8 odd multiples = 32 KiB per curve, staged in
// SHARED memory, and an addend is read by scanning ALL 8 entries of its window under masks -- the addresses a lane touches do
// not depend on its scalar (the 16-bit-window table is indexed by secret digits: no divergence, but a data-dependent address
// stream into the L2).  63 additions instead of 15.
constexpr int CT_W = 4;
constexpr int CT_WINDOWS = 256 / CT_W;
constexpr int CT_ENTRIES = 1 << (CT_W - 1);
constexpr int CT_TABLE_WORDS = CT_WINDOWS * CT_ENTRIES * 16;
struct ScanComb {
  const u32* base;  // shared memory: entry (i, j) at word (i * 8 + j) * 16
  FEC_DEV void operator()(int i, u32 idx, u32 x[8], u32 y[8]) const {
#pragma unroll
    for (int w = 0; w < 8; w++) { x[w] = 0u; y[w] = 0u; }
#pragma unroll
    for (u32 j = 0; j < (u32)CT_ENTRIES; j++) {
      const u32* e = base + ((size_t)i * CT_ENTRIES + j) * 16;  // the same address in every lane: a broadcast read
      const u32 m = (j == idx) ? 0xFFFFFFFFu : 0u;
#pragma unroll
      for (int w = 0; w < 8; w++) { x[w] |= e[w] & m; y[w] |= e[8 + w] & m; }
    }
  }
};

// the Ed25519 / X25519-base-point counterpart: 64 x 8 Niels triples = 48 KiB in shared memory, every triple of a window scanned
constexpr int CT_NIELS_WORDS = CT_WINDOWS * CT_ENTRIES * 24;
struct ScanNiels {
  const u32* base;  // shared memory: entry (i, j) at word (i * 8 + j) * 24
  FEC_DEV void operator()(int i, u32 idx, u32 a[8], u32 b[8], u32 c[8]) const {
#pragma unroll
    for (int w = 0; w < 8; w++) { a[w] = 0u; b[w] = 0u; c[w] = 0u; }
#pragma unroll
    for (u32 j = 0; j < (u32)CT_ENTRIES; j++) {
      const u32* e = base + ((size_t)i * CT_ENTRIES + j) * 24;  // the same address in every lane: a broadcast read
      const u32 m = (j == idx) ? 0xFFFFFFFFu : 0u;
#pragma unroll
      for (int w = 0; w < 8; w++) { a[w] |= e[w] & m; b[w] |= e[8 + w] & m; c[w] |= e[16 + w] & m; }
    }
  }
};

// -------------------------------------------------------------------------------------
// secp256k1 only: GLV endomorphism phi(x,y) = (beta x, y) = lambda (x,y).
// k = k1 + k2 lambda (mod n) with |k1|, |k2| < 2^128 (lattice basis and rounding constants g1, g2 as
// in libsecp256k1's scalar_split_lambda; re-derived and bound-checked in tests/test_oracle.py):
//   c1 = round(k b2 / n) = (k g1 + 2^383) >> 384,  c2 = round(-k b1 / n) = (k g2 + 2^383) >> 384
//   k1 = k - c1 a1 - c2 a2,  k2 = -c1 b1 - c2 b2      (exact integers; computed mod 2^256, two's complement)
// Halves the doublings of the variable-base loop: 124 doublings + 66 mixed additions instead of 252 + 63.
// -------------------------------------------------------------------------------------
template <int NA, int NB>
FEC_DEV void mul_limbs(u32* r, const u32* a, const u32* b) {  // r[NA+NB] = a[NA] * b[NB]
#pragma unroll
  for (int i = 0; i < NA + NB; i++) r[i] = 0u;
#pragma unroll
  for (int i = 0; i < NA; i++) {
    u64 carry = 0;
#pragma unroll
    for (int j = 0; j < NB; j++) {
      u64 t = (u64)a[i] * b[j] + r[i + j] + carry;
      r[i + j] = (u32)t;
      carry = t >> 32;
    }
    r[i + NB] = (u32)carry;
  }
}
FEC_DEV void k1_beta(u32 b[8]) {
  b[0] = 0x719501EEu; b[1] = 0xC1396C28u; b[2] = 0x12F58995u; b[3] = 0x9CF04975u;
  b[4] = 0xAC3434E9u; b[5] = 0x6E64479Eu; b[6] = 0x657C0710u; b[7] = 0x7AE96A2Bu;
}
// magnitudes s1, s2 (128 bits, 4 limbs) and sign masks of k1, k2
FEC_DEV void glv_split(u32 s1[4], u32& neg1, u32 s2[4], u32& neg2, const u32 k[8]) {
  const u32 G1[8] = {0x45DBB031u, 0xE893209Au, 0x71E8CA7Fu, 0x3DAA8A14u, 0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u};
  const u32 G2[8] = {0x8AC47F71u, 0x1571B4AEu, 0x9DF506C6u, 0x221208ACu, 0x0ABFE4C4u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u};
  const u32 A1[4] = {0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u};           // a1 = b2
  const u32 B1N[4] = {0x0ABFE4C3u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u};          // -b1
  const u32 A2[5] = {0x9D44CFD8u, 0x57C1108Du, 0xA8E2F3F6u, 0x14CA50F7u, 0x00000001u};
  u32 prod[16], c1[4], c2[4];
  mul_limbs<8, 8>(prod, k, G1);
  {
    u64 t = (u64)prod[11] + 0x80000000u;
    u64 c = t >> 32;
#pragma unroll
    for (int i = 0; i < 4; i++) { t = (u64)prod[12 + i] + c; c1[i] = (u32)t; c = t >> 32; }
  }
  mul_limbs<8, 8>(prod, k, G2);
  {
    u64 t = (u64)prod[11] + 0x80000000u;
    u64 c = t >> 32;
#pragma unroll
    for (int i = 0; i < 4; i++) { t = (u64)prod[12 + i] + c; c2[i] = (u32)t; c = t >> 32; }
  }
  u32 t1[8], t2[8], u2[9], k1[8], k2[8], z[8], nk[8];
  mul_limbs<4, 4>(t1, c1, B1N);  // -c1 b1
  mul_limbs<4, 4>(t2, c2, A1);   //  c2 b2
  sub8(k2, t1, t2);
  mul_limbs<4, 4>(t1, c1, A1);   //  c1 a1
  mul_limbs<4, 5>(u2, c2, A2);   //  c2 a2 (low 256 bits are enough)
  sub8(k1, k, t1);
  sub8(k1, k1, u2);
  set_small(z, 0u);
  neg1 = 0u - (k1[7] >> 31);
  sub8(nk, z, k1);
  select8(k1, neg1, nk, k1);
  neg2 = 0u - (k2[7] >> 31);
  sub8(nk, z, k2);
  select8(k2, neg2, nk, k2);
#pragma unroll
  for (int i = 0; i < 4; i++) { s1[i] = k1[i]; s2[i] = k2[i]; }
}
// 128-bit odd s -> K = 2^127 | (s >> 1); digits as in next_digit, 32 of them
FEC_DEV void recode_reg4(u32 K[4], const u32 s[4]) {
#pragma unroll
  for (int i = 0; i < 3; i++) K[i] = funnel_l(s[i], s[i + 1], 31);
  K[3] = (s[3] >> 1) | 0x80000000u;
}
FEC_DEV void next_digit4(u32 K[4], u32& idx, u32& sgn) {
  u32 u = K[3] >> 28;
#pragma unroll
  for (int i = 3; i >= 1; i--) K[i] = funnel_l(K[i - 1], K[i], 4);
  K[0] <<= 4;
  sgn = (u >> 3) - 1u;
  idx = (u ^ sgn) & 7u;
}
// mixed addition with a skip mask, exceptional cases repaired under a warp-uniform branch
template <class C>
FEC_DEV void jac_madd_safe_skip(Jac& q, const u32 tx[8], const u32 ty[8], u32 skip) {
  u32 exc = jac_madd<C>(q, tx, ty, skip);
  if (FEC_UNLIKELY(warp_any(exc))) {
    Jac d;
    copy8(d.X, tx); copy8(d.Y, ty); C::F::one(d.Z);
    jac_dbl<C>(d);
    u32 isdbl = 0u - ((exc >> 1) & 1u);
    u32 isinf = (0u - (exc & 1u)) & ~isdbl;
    u32 one[8], zero[8];
    C::F::one(one); set_small(zero, 0u);
    select8(q.X, isdbl, d.X, q.X); select8(q.Y, isdbl, d.Y, q.Y); select8(q.Z, isdbl, d.Z, q.Z);
    select8(q.X, isinf, one, q.X); select8(q.Y, isinf, one, q.Y); select8(q.Z, isinf, zero, q.Z);
  }
}
// k in [1, n-1], (px,py) on the curve.  Result: Jacobian (Q.X, Q.Y, Q.Z * zs) -- the loop runs on the curve
// isomorphic by zs = Z(2P) * Z(15P'), where all eight table entries are affine ("effective affine"; a = 0 only),
// so building the table needs no inversion at all.
template <class Tab>
FEC_DEV void wmul_var_glv_core(Jac& Q, u32 zs[8], const u32 k[8], const u32 px[8], const u32 py[8], Tab& tab) {
  typedef CurveK1 C;
  typedef FeK1 F;
  {
    Jac D, A;
    u32 C2[8], C3[8], s[8];
    u32 zr[7][8];
    copy8(D.X, px); copy8(D.Y, py); set_small(D.Z, 1u);
    jac_dbl<C>(D);
    F::sqr(C2, D.Z);
    F::mul(C3, C2, D.Z);
    F::mul(A.X, px, C2);
    F::mul(A.Y, py, C3);
    set_small(A.Z, 1u);
    tab.store(0, A.X, A.Y);
#pragma unroll 1
    for (int j = 1; j < 8; j++) {
      jac_madd_zr<C>(A, D.X, D.Y, zr[j - 1]);
      tab.store(j, A.X, A.Y);
    }
    F::mul(zs, A.Z, D.Z);
    copy8(s, zr[6]);
#pragma unroll 1
    for (int j = 6; j >= 0; j--) {
      u32 x[8], y[8], s2[8];
      tab.load(j, x, y);
      F::sqr(s2, s);
      F::mul(x, x, s2);
      F::mul(s2, s2, s);
      F::mul(y, y, s2);
      tab.store(j, x, y);
      if (j > 0) F::mul(s, s, zr[j - 1]);
    }
  }
  u32 s1[4], s2[4], neg1, neg2, K1r[4], K2r[4], beta[8];
  glv_split(s1, neg1, s2, neg2, k);
  const u32 corr1 = (s1[0] & 1u) - 1u, corr2 = (s2[0] & 1u) - 1u;  // even -> use s+1, subtract the base afterwards
  s1[0] |= 1u;
  s2[0] |= 1u;
  recode_reg4(K1r, s1);
  recode_reg4(K2r, s2);
  k1_beta(beta);
  u32 idx, sgn, tx[8], ty[8], ny[8];
  next_digit4(K1r, idx, sgn);
  tab.load(idx, Q.X, Q.Y);
  F::neg(ny, Q.Y);
  select8(Q.Y, sgn ^ neg1, ny, Q.Y);
  set_small(Q.Z, 1u);
  next_digit4(K2r, idx, sgn);
  tab.load(idx, tx, ty);
  F::mul(tx, tx, beta);
  F::neg(ny, ty);
  select8(ty, sgn ^ neg2, ny, ty);
  jac_madd_safe<C>(Q, tx, ty);
#if defined(FEC_HOT_INLINE) && FEC_HOT_INLINE == 2
  typedef CurveK1Inl CH;
#elif defined(FEC_HOT_INLINE) && FEC_HOT_INLINE == 1
  typedef CurveK1InlSqr CH;
#else
  typedef CurveK1 CH;
#endif
#pragma unroll 1
  for (int i = 1; i < 32; i++) {
#pragma unroll 1
    for (int d = 0; d < 4; d++) jac_dbl<CH>(Q);
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
      if (h == 0) next_digit4(K1r, idx, sgn); else next_digit4(K2r, idx, sgn);
      tab.load(idx, tx, ty);
      if (h == 1) F::mul(tx, tx, beta);
      F::neg(ny, ty);
      select8(ty, sgn ^ (h == 0 ? neg1 : neg2), ny, ty);
      jac_madd_safe<CH>(Q, tx, ty);
    }
  }
  // parity corrections: subtract sign(k1) P and/or sign(k2) lambda P where s was bumped to s + 1
#pragma unroll 1
  for (int h = 0; h < 2; h++) {
    tab.load(0, tx, ty);
    if (h == 1) F::mul(tx, tx, beta);
    F::neg(ny, ty);
    select8(ty, ~(h == 0 ? neg1 : neg2), ny, ty);
    jac_madd_safe_skip<C>(Q, tx, ty, ~(h == 0 ? corr1 : corr2));
  }
}

// Jacobian -> affine with a supplied 1/Z, conditional negation, canonical limbs in the field's internal
// representation (table entries, intermediate points); jac_to_affine_plain gives the integers that are encoded
template <class C>
FEC_DEV void jac_to_affine(u32 x[8], u32 y[8], const Jac& Q, const u32 zi[8], u32 negm) {
  typedef typename C::F F;
  u32 z2[8], t[8], ny[8];
  F::sqr(z2, zi);
  F::mul(t, Q.X, z2);
  F::canon(x, t);
  F::mul(z2, z2, zi);
  F::mul(t, Q.Y, z2);
  F::neg(ny, t);
  select8(t, negm, ny, t);
  F::canon(y, t);
}
template <class C>
FEC_DEV void jac_to_affine_plain(u32 x[8], u32 y[8], const Jac& Q, const u32 zi[8], u32 negm) {
  typedef typename C::F F;
  u32 z2[8], t[8], ny[8];
  F::sqr(z2, zi);
  F::mul(t, Q.X, z2);
  F::to_plain(x, t);
  F::mul(z2, z2, zi);
  F::mul(t, Q.Y, z2);
  F::neg(ny, t);
  select8(t, negm, ny, t);
  F::to_plain(y, t);
}
// affine point from its byte record (big-endian X || Y already loaded as integers): range and curve checks, then
// the internal representation.  Returns the ok mask; x, y are only meaningful where it is set.
template <class C>
FEC_DEV u32 affine_from_plain(u32 x[8], u32 y[8]) {
  typedef typename C::F F;
  const u32 ok = F::in_range(x) & F::in_range(y);
  F::from_plain(x, x);
  F::from_plain(y, y);
  return ok & on_curve<C>(x, y);
}
// The same for callers that accept the identity as an input point: `idn` is set for the all-zero X || Y record, the engine's own
// encoding of the identity (point.rs:41-47), which never satisfies the curve equation (b != 0) -- the ok mask is clear for it.
// Curve::multiply(identity, k) is the identity and multi_scalar_multiply simply adds it (forge-ec-core/src/lib.rs:820-832,
// 934-951), so such a lane is reported as LANE_INFINITY, not as an invalid point.
template <class C>
FEC_DEV u32 affine_from_plain_id(u32 x[8], u32 y[8], u32& idn) {
  idn = is_zero8(x) & is_zero8(y);
  return affine_from_plain<C>(x, y);
}
// lane status from the three masks, priority: scalar range > invalid point > identity (k = 0 or identity input)
FEC_DEV u32 lane_status(u32 st_scalar, u32 ok, u32 idn) {
  if (st_scalar == LANE_SCALAR_RANGE) return st_scalar;
  return (idn & LANE_INFINITY) | (~idn & ((~ok & LANE_POINT_INVALID) | (ok & st_scalar)));
}

// SEC1 / raw affine encodings.  fmt: 0 = 02|03 || X (33 B), 1 = 04 || X || Y (65 B), 2 = X || Y (64 B), 3 = X (32 B,
// the ECDH shared secret of KeyExchange::derive_shared_secret, secp256k1.rs:1885-1905), 4 = 06|07 || X || Y (65 B,
// PointFormat::Hybrid, secp256k1.rs:1048-1066 / p256.rs:1704-1722).
// dead != 0 -> all-zero record (the reference's identity encoding, point.rs:41-47).
FEC_DEV void sec1_encode(uint8_t* o, int fmt, const u32 x[8], const u32 y[8], u32 dead) {
  u32 xx[8], yy[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { xx[i] = x[i] & ~dead; yy[i] = y[i] & ~dead; }
  if (fmt == 0) {
    o[0] = (uint8_t)((2u | (y[0] & 1u)) & ~dead);
    store_be32(o + 1, xx);
  } else if (fmt == 1) {
    o[0] = (uint8_t)(4u & ~dead);
    store_be32(o + 1, xx);
    store_be32(o + 33, yy);
  } else if (fmt == 2) {
    store_be32(o, xx);
    store_be32(o + 32, yy);
  } else if (fmt == 4) {
    o[0] = (uint8_t)((6u | (y[0] & 1u)) & ~dead);
    store_be32(o + 1, xx);
    store_be32(o + 33, yy);
  } else {
    store_be32(o, xx);
  }
}

// SEC1 compressed -> affine candidate (decode half of point.rs:70-144): returns the ok mask.
// sec1_decompress_id additionally reports the identity record (prefix 0x00, point.rs:82-87) in `idn` (ok is clear for it).
template <class C>
FEC_DEV u32 sec1_decompress(u32 x[8], u32 y[8], const uint8_t* rec);
template <class C>
FEC_DEV u32 sec1_decompress_id(u32 x[8], u32 y[8], const uint8_t* rec, u32& idn) {
  idn = rec[0] == 0u ? 0xFFFFFFFFu : 0u;
  return sec1_decompress<C>(x, y, rec);
}
template <class C>
FEC_DEV u32 sec1_decompress(u32 x[8], u32 y[8], const uint8_t* rec) {
  typedef typename C::F F;
  u32 pre = rec[0];
  load_be32(x, rec + 1);
  u32 ok = ((pre == 2u) | (pre == 3u)) ? 0xFFFFFFFFu : 0u;
  ok &= F::in_range(x);
  F::from_plain(x, x);
  u32 rhs[8], t[8], b[8], yc[8], ny[8], yp[8];
  F::sqr(rhs, x);
  F::mul(rhs, rhs, x);
  if (C::A_MINUS3) {
    F::dbl(t, x);
    F::add(t, t, x);
    F::sub(rhs, rhs, t);
  }
  C::B(b);
  F::add(rhs, rhs, b);
  F::sqrt_cand(t, rhs);
  F::canon(yc, t);
  F::sqr(t, yc);
  F::sub(t, t, rhs);
  ok &= F::is_zero(t);
  F::neg(t, yc);
  F::canon(ny, t);
  F::to_plain(yp, yc);  // the parity that the prefix byte names is that of the integer y
  u32 flip = 0u - ((yp[0] ^ pre) & 1u);
  select8(y, flip, ny, yc);
  return ok;
}

// -------------------------------------------------------------------------------------
// X25519: clamp, mask bit 255 of u, 255 ladder steps; returns (x2 : z2)
// -------------------------------------------------------------------------------------
FEC_DEV void cswap8(u32 a[8], u32 b[8], u32 m) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u32 t = (a[i] ^ b[i]) & m;
    a[i] ^= t;
    b[i] ^= t;
  }
}
FEC_DEV void x25519_core(u32 x2[8], u32 z2[8], const u32 k_le[8], const u32 u_le[8]) {
  u32 k[8], x1[8], x3[8], z3[8];
  copy8(k, k_le);
  k[0] &= 0xFFFFFFF8u;
  k[7] = (k[7] & 0x7FFFFFFFu) | 0x40000000u;
  copy8(x1, u_le);
  x1[7] &= 0x7FFFFFFFu;
  set_small(x2, 1u); set_small(z2, 0u);
  copy8(x3, x1); set_small(z3, 1u);
  // bits 254..0, consumed from the top of a shifting register (bit 255 is clear after clamping)
#pragma unroll
  for (int i = 7; i >= 1; i--) k[i] = funnel_l(k[i - 1], k[i], 1);
  k[0] <<= 1;
  // (x2, z2) is the doubling of the previous step, (x3, z3) its differential addition (see x25519_step); prev = the
  // previous bit: the pair that holds [m]P is (x3, z3) iff prev is set
  u32 prev = 0u;
#pragma unroll 1
  for (int t = 254; t >= 0; t--) {
    u32 kt = 0u - (k[7] >> 31);
#pragma unroll
    for (int i = 7; i >= 1; i--) k[i] = funnel_l(k[i - 1], k[i], 1);
    k[0] <<= 1;
    x25519_step(x2, z2, x3, z3, x1, prev ^ kt);
    prev = kt;
  }
  select8(x2, prev, x3, x2);
  select8(z2, prev, z3, z2);
}

// -------------------------------------------------------------------------------------
// SHA-512 of a 32-byte message (one block) -- FIPS 180-4; only what Ed25519 key expansion needs
// (the reference calls the `sha2` crate: eddsa.rs:464; KATs at forge-ec-hash/src/sha2.rs:13-41)
// -------------------------------------------------------------------------------------
#ifdef __CUDACC__
__constant__ u64 FEC_SHA512_K_TABLE[80] = {FEC_SHA512_K};
#else
static const u64 FEC_SHA512_K_TABLE[80] = {FEC_SHA512_K};
#endif
FEC_DEV u64 rotr64(u64 x, int n) { return (x >> n) | (x << (64 - n)); }
// digest words h[0..8) (big-endian words as integers)
FEC_DEV void sha512_32(u64 h[8], const uint8_t msg[32]) {
  const u64 H0[8] = {FEC_SHA512_H0};
  u64 w[16];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    u64 v = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) v = (v << 8) | msg[8 * i + j];
    w[i] = v;
  }
  w[4] = 0x8000000000000000ULL;
#pragma unroll
  for (int i = 5; i < 15; i++) w[i] = 0;
  w[15] = 256;
  u64 a = H0[0], b = H0[1], c = H0[2], d = H0[3], e = H0[4], f = H0[5], g = H0[6], hh = H0[7];
#pragma unroll 1
  for (int r = 0; r < 80; r += 16) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (r > 0) {
        u64 w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
        u64 s0 = rotr64(w15, 1) ^ rotr64(w15, 8) ^ (w15 >> 7);
        u64 s1 = rotr64(w2, 19) ^ rotr64(w2, 61) ^ (w2 >> 6);
        w[i] = w[i] + s0 + w[(i + 9) & 15] + s1;
      }
      u64 S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41);
      u64 ch = (e & f) ^ (~e & g);
      u64 t1 = hh + S1 + ch + FEC_SHA512_K_TABLE[r + i] + w[i];
      u64 S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39);
      u64 mj = (a & b) ^ (a & c) ^ (b & c);
      u64 t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
  }
  h[0] = H0[0] + a; h[1] = H0[1] + b; h[2] = H0[2] + c; h[3] = H0[3] + d;
  h[4] = H0[4] + e; h[5] = H0[5] + f; h[6] = H0[6] + g; h[7] = H0[7] + hh;
}
FEC_DEV u32 bswap32(u32 v) { return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24); }
// seed -> little-endian limbs of the clamped secret scalar (first half of SHA-512(seed))
FEC_DEV void ed25519_expand_seed(u32 s[8], const uint8_t seed[32]) {
  u64 h[8];
  sha512_32(h, seed);
  // digest byte 8i+j = byte (7-j) of h[i]; little-endian limb l = bytes 4l..4l+3
#pragma unroll
  for (int i = 0; i < 4; i++) {
    s[2 * i] = bswap32((u32)(h[i] >> 32));
    s[2 * i + 1] = bswap32((u32)h[i]);
  }
  s[0] &= 0xFFFFFFF8u;
  s[7] = (s[7] & 0x7FFFFFFFu) | 0x40000000u;
}
FEC_DEV void ed_L(u32 l[8]) {  // group order 2^252 + 27742317777372353535851937790883648493
  l[0] = 0x5CF5D3EDu; l[1] = 0x5812631Au; l[2] = 0xA2F79CD6u; l[3] = 0x14DEF9DEu;
  l[4] = 0u; l[5] = 0u; l[6] = 0u; l[7] = 0x10000000u;
}
// any 256-bit s -> recoding register of an odd s' = s (mod L): s even -> s -+ L
FEC_DEV void ed_scalar_prepare(u32 K[8], const u32 s[8]) {
  u32 l[8], a[8], b[8], t[8];
  ed_L(l);
  add8(a, s, l);
  sub8(b, s, l);
  u32 hi = 0u - (s[7] >> 31);
  select8(t, hi, b, a);
  u32 even = (s[0] & 1u) - 1u;
  select8(t, even, t, s);
  recode_reg(K, t);
}
// Q = sum of COMB_WINDOWS signed odd digits times the Niels comb: niels(i, idx, ypx, ymx, t2d) = (2 idx+1) 2^(W i) B
template <class Comb, int W = COMB_W>
FEC_DEV void ed_comb_core(EdExt& Q, u32 K[8], const Comb& niels) {
  typedef Fe25519 F;
  constexpr int WINDOWS = 256 / W;
  ed_set_identity(Q);
  u32 idx, sgn, na[8], nb[8], nc[8];
  next_digit_w<W>(K, idx, sgn);
  niels(WINDOWS - 1, idx, na, nb, nc);
#pragma unroll 1
  for (int i = WINDOWS - 1; i >= 0; i--) {
    u32 ypx[8], ymx[8], t2d[8];
#ifdef FEC_ED_SIGN_BY_NEGATION
    u32 nt[8];
#endif
    const u32 sg = sgn;
    copy8(ypx, na); copy8(ymx, nb); copy8(t2d, nc);
    if (i > 0) {  // fetch the next addend while this addition runs
      next_digit_w<W>(K, idx, sgn);
      niels(i - 1, idx, na, nb, nc);
    }
#ifndef FEC_ED_SIGN_BY_NEGATION
    ed_madd_signed(Q, ypx, ymx, t2d, sg);
#else
    cswap8(ypx, ymx, sg);
    F::neg(nt, t2d);
    select8(t2d, sg, nt, t2d);
    ed_madd(Q, ypx, ymx, t2d);
#endif
  }
}
// RFC 8032 5.1.2 encoding with a supplied 1/Z: y little-endian, bit 255 = x & 1
FEC_DEV void ed_encode(u32 out_le[8], const EdExt& Q, const u32 zi[8]) {
  typedef Fe25519 F;
  u32 x[8], y[8], t[8];
  F::mul(t, Q.X, zi);
  F::canon(x, t);
  F::mul(t, Q.Y, zi);
  F::canon(y, t);
  copy8(out_le, y);
  out_le[7] |= (x[0] & 1u) << 31;
}

// -------------------------------------------------------------------------------------
// comb table entries (context set-up only; control flow depends on the public indices i, j)
// -------------------------------------------------------------------------------------
// window base 2^(W i) G, affine (one Fermat inversion; 16 of them per curve)
template <class C, int W = COMB_W>
FEC_DEV void comb_base(u32 bx[8], u32 by[8], int i) {
  typedef typename C::F F;
  Jac R;
  C::G(R.X, R.Y);
  F::one(R.Z);
  for (int d = 0; d < W * i; d++) jac_dbl<C>(R);
  u32 zi[8];
  F::inv(zi, R.Z);
  jac_to_affine<C>(bx, by, R, zi, 0u);
}
// (2j+1) (bx,by), Jacobian: MSB-first double-and-add over the COMB_W bits of 2j+1 (no exceptional case can occur:
// the running multiple is >= 2 whenever the base is added)
template <class C>
FEC_DEV void comb_entry_jac(Jac& R, const u32 bx[8], const u32 by[8], int j) {
  copy8(R.X, bx); copy8(R.Y, by); C::F::one(R.Z);
  const int m = 2 * j + 1;
  int top = COMB_W - 1;
  while (top > 0 && !((m >> top) & 1)) top--;
  for (int b = top - 1; b >= 0; b--) {
    jac_dbl<C>(R);
    if ((m >> b) & 1) jac_madd<C>(R, bx, by, 0u);
  }
}
// Ed25519 window base 2^(W i) B in extended coordinates with Z = 1
template <int W = COMB_W>
FEC_DEV void ed_comb_base(EdExt& Bi, int i) {
  typedef Fe25519 F;
  ed_base(Bi.X, Bi.Y);
  set_small(Bi.Z, 1u);
  F::mul(Bi.T, Bi.X, Bi.Y);
  for (int d = 0; d < W * i; d++) {
    EdExt c = Bi;
    ed_add(Bi, c);
  }
  u32 zi[8], x[8], y[8];
  F::inv(zi, Bi.Z);
  F::mul(x, Bi.X, zi);
  F::mul(y, Bi.Y, zi);
  F::canon(Bi.X, x);
  F::canon(Bi.Y, y);
  set_small(Bi.Z, 1u);
  F::mul(Bi.T, Bi.X, Bi.Y);
}
// (2j+1) Bi as a Niels triple (y+x, y-x, 2dxy), canonical limbs
template <int W = COMB_W>
FEC_DEV void ed_comb_entry(u32 ypx[8], u32 ymx[8], u32 t2d[8], const EdExt& Bi, int j) {
  typedef Fe25519 F;
  EdExt R = Bi;
  const int m = 2 * j + 1;
  int top = W - 1;
  while (top > 0 && !((m >> top) & 1)) top--;
  for (int b = top - 1; b >= 0; b--) {
    EdExt c = R;
    ed_add(R, c);
    if ((m >> b) & 1) ed_add(R, Bi);
  }
  u32 zi[8], x[8], y[8], t[8], d2[8];
  F::inv(zi, R.Z);
  F::mul(x, R.X, zi);
  F::mul(y, R.Y, zi);
  F::add(t, y, x); F::canon(ypx, t);
  F::sub(t, y, x); F::canon(ymx, t);
  ed_D2(d2);
  F::mul(t, x, y);
  F::mul(t, t, d2);
  F::canon(t2d, t);
}

}  // namespace fec
