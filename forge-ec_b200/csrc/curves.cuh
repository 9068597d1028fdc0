// curves.cuh -- group law on register-resident points, branch-free.
//
//  * short Weierstrass curves in Jacobian coordinates (secp256k1: a=0, P-256: a=-3):
//    doubling and mixed addition.  Replaces the per-bit `r0+r1 / 2r0 / 2r1` ladder of
//    forge-ec-curves/src/secp256k1.rs:2635-2692 (+ Add :1447-1497, double :1502-1540) and the
//    secret-branching double-and-add of p256.rs:2120-2156 (+ double :1869-1912, Add :1941-2006).
//  * twisted Edwards curve Ed25519 in extended coordinates with Niels-form table entries
//    (replaces ed25519.rs:1864-1927 `Add`, :1828-1832 `double`, :2062-2097 `multiply`).
//  * Montgomery x-only ladder step for X25519 (replaces curve25519.rs:1671-1701).
//  * block-level Montgomery-trick inversion (replaces the per-point Fermat inversion inside
//    every `to_affine`: secp256k1.rs:1342-1363, p256.rs:1835-1857, ed25519.rs:1793-1811,
//    curve25519.rs:1711-1713).
#pragma once
#include "fields.cuh"

namespace fec {

struct CurveK1 {
  typedef FeK1 F;
  typedef CurveK1 Cold;  // the instantiation used on rarely executed paths
  static constexpr int CURVE_ID = 0;
  static constexpr bool A_MINUS3 = false;
  // which pipe bounds the field arithmetic decides the point formulas of the long loops (vm.cuh): false = multiply pipe
  // (squarings are cheap, additions free), true = ALU pipe (a canonical-form addition costs a quarter of a squaring: trade
  // additions for multiplications).  The register-resident mixed addition below keeps 7M + 4S for both curves: with the
  // field routines called out of line the 8M + 3S form measured 5 % slower in the fixed-base kernel.
  static constexpr bool ALU_BOUND = false;
  FEC_DEV static void B(u32 b[8]) {
    b[0] = 7u;
#pragma unroll
    for (int i = 1; i < 8; i++) b[i] = 0u;
  }
  FEC_DEV static void N(u32 n[8]) {  // group order
    n[0] = 0xD0364141u; n[1] = 0xBFD25E8Cu; n[2] = 0xAF48A03Bu; n[3] = 0xBAAEDCE6u;
    n[4] = 0xFFFFFFFEu; n[5] = 0xFFFFFFFFu; n[6] = 0xFFFFFFFFu; n[7] = 0xFFFFFFFFu;
  }
  FEC_DEV static void G(u32 x[8], u32 y[8]) {
    x[0] = 0x16F81798u; x[1] = 0x59F2815Bu; x[2] = 0x2DCE28D9u; x[3] = 0x029BFCDBu;
    x[4] = 0xCE870B07u; x[5] = 0x55A06295u; x[6] = 0xF9DCBBACu; x[7] = 0x79BE667Eu;
    y[0] = 0xFB10D4B8u; y[1] = 0x9C47D08Fu; y[2] = 0xA6855419u; y[3] = 0xFD17B448u;
    y[4] = 0x0E1108A8u; y[5] = 0x5DA4FBFCu; y[6] = 0x26A3C465u; y[7] = 0x483ADA77u;
  }
};

struct CurveP256 {
  typedef FeP256 F;
  typedef CurveP256 Cold;
  static constexpr int CURVE_ID = 1;
  static constexpr bool A_MINUS3 = true;
  static constexpr bool ALU_BOUND = true;
  FEC_DEV static void B(u32 b[8]) {
    // b * 2^256 mod p (FeP256 works in the Montgomery domain; b = 5AC635D8 AA3A93E7 B3EBBD55 769886BC 651D06B0 CC53B0F6 3BCE3C3E 27D2604B)
    b[0] = 0x29C4BDDFu; b[1] = 0xD89CDF62u; b[2] = 0x78843090u; b[3] = 0xACF005CDu;
    b[4] = 0xF7212ED6u; b[5] = 0xE5A220ABu; b[6] = 0x04874834u; b[7] = 0xDC30061Du;
  }
  FEC_DEV static void N(u32 n[8]) {
    n[0] = 0xFC632551u; n[1] = 0xF3B9CAC2u; n[2] = 0xA7179E84u; n[3] = 0xBCE6FAADu;
    n[4] = 0xFFFFFFFFu; n[5] = 0xFFFFFFFFu; n[6] = 0x00000000u; n[7] = 0xFFFFFFFFu;
  }
  FEC_DEV static void G(u32 x[8], u32 y[8]) {
    // (Gx, Gy) * 2^256 mod p; Gx = 6B17D1F2 E12C4247 F8BCE6E5 63A440F2 77037D81 2DEB33A0 F4A13945 D898C296,
    // Gy = 4FE342E2 FE1A7F9B 8EE7EB4A 7C0F9E16 2BCE3357 6B315ECE CBB64068 37BF51F5 (FIPS 186-4; p256.rs:2094-2106)
    x[0] = 0x18A9143Cu; x[1] = 0x79E730D4u; x[2] = 0x5FEDB601u; x[3] = 0x75BA95FCu;
    x[4] = 0x77622510u; x[5] = 0x79FB732Bu; x[6] = 0xA53755C6u; x[7] = 0x18905F76u;
    y[0] = 0xCE95560Au; y[1] = 0xDDF25357u; y[2] = 0xBA19E45Cu; y[3] = 0x8B4AB8E4u;
    y[4] = 0xDD21F325u; y[5] = 0xD2E88688u; y[6] = 0x25885D85u; y[7] = 0x8571FF18u;
  }
};

// Same curves with field multiplication / squaring expanded in line (hot loops whose body is small enough to
// stay in the instruction cache); everything else is inherited.
struct FeK1Inl : FeK1 {
  FEC_DEV static void mul(u32 o[8], const u32 a[8], const u32 b[8]) { mul_inl(o, a, b); }
  FEC_DEV static void sqr(u32 o[8], const u32 a[8]) { sqr_inl(o, a); }
};
struct FeK1InlSqr : FeK1 {
  FEC_DEV static void sqr(u32 o[8], const u32 a[8]) { sqr_inl(o, a); }
};
struct CurveK1Inl : CurveK1 { typedef FeK1Inl F; };
struct CurveK1InlSqr : CurveK1 { typedef FeK1InlSqr F; };

// -------------------------------------------------------------------------------------
// Jacobian point (X:Y:Z), infinity <=> Z == 0.  Coordinates are field elements in the field's internal
// representation (F::one() is the field's 1; G() and B() return constants in that representation).
// -------------------------------------------------------------------------------------
struct Jac {
  u32 X[8], Y[8], Z[8];
};

template <class C>
FEC_DEV void jac_set_inf(Jac& q) {
#pragma unroll
  for (int i = 0; i < 8; i++) { q.X[i] = 0; q.Y[i] = 0; q.Z[i] = 0; }
  q.X[0] = 1; q.Y[0] = 1;
}

// Q <- 2Q, in place.  a=0: dbl-2009-l (2M+5S).  a=-3: 4M+4S with 2Y carried through (Z3 = 2Y Z, 4 beta = X (2Y)^2,
// 8 gamma^2 = (2Y)^4 / 2): 4 additions + 5 subtractions + a halving instead of the 9 + 7 of dbl-2001-b.  Z=0 stays Z=0.
// Written to keep few temporaries alive (Q + 4 field elements) and to use the single-pass small multiples.
template <class C>
FEC_DEV void jac_dbl(Jac& q) {
  typedef typename C::F F;
  if constexpr (!C::A_MINUS3) {
    u32 A[8], B[8], Cc[8], t[8];
    F::sqr(A, q.X);
    F::sqr(B, q.Y);
    F::sqr(Cc, B);
    F::add(t, q.X, B);
    F::sqr(t, t);
    F::sub(t, t, A);
    F::sub(t, t, Cc);
    F::mul2(B, t);              // D = 2((X+B)^2 - A - C)        (B is free)
    F::mul(q.Z, q.Y, q.Z);
    F::mul2(q.Z, q.Z);          // Z3 = 2YZ
    F::mul3(A, A);              // E = 3A
    F::sqr(q.X, A);             // F = E^2
    F::mul2(t, B);
    F::sub(q.X, q.X, t);        // X3 = F - 2D
    F::sub(t, B, q.X);
    F::mul(t, A, t);            // E(D - X3)
    F::mul8(Cc, Cc);            // 8C
    F::sub(q.Y, t, Cc);
  } else {
    u32 delta[8], g4[8], beta[8], alpha[8], t[8];
    F::sqr(delta, q.Z);
    F::sub(t, q.X, delta);
    F::add(alpha, q.X, delta);
    F::mul(alpha, t, alpha);
    F::mul3(alpha, alpha);      // alpha = 3(X-delta)(X+delta)
    F::mul2(t, q.Y);            // 2Y
    F::mul(q.Z, t, q.Z);        // Z3 = 2YZ
    F::sqr(g4, t);              // 4 gamma
    F::mul(beta, q.X, g4);      // 4 beta
    F::sqr(t, alpha);
    F::sub(t, t, beta);
    F::sub(q.X, t, beta);       // X3 = alpha^2 - 8 beta
    F::sub(t, beta, q.X);
    F::mul(t, alpha, t);
    F::sqr(g4, g4);
    F::half(g4, g4);            // 8 gamma^2
    F::sub(q.Y, t, g4);
  }
}

// Q <- Q + (x2,y2) with (x2,y2) affine: madd-2007-bl (7M+4S), IN PLACE and without any selection.
// Returns a flag word: bit 0 = the formula does not apply (Q at infinity, or Q == +-(x2,y2)),
// bit 1 = Q was at infinity, bit 2 = same y as well (doubling).  When bit 0 is set Q holds garbage and the caller
// repairs the lane (jac_madd_fix) -- under a warp-uniform branch, so no divergence depends on the data.
template <class C>
FEC_DEV u32 jac_madd_lean(Jac& q, const u32 x2[8], const u32 y2[8]) {
  typedef typename C::F F;
  u32 A[8], H[8], r[8], D[8];
  const u32 q_inf = F::is_zero(q.Z);
  F::sqr(A, q.Z);                    // Z1Z1
  F::mul(H, x2, A);                  // U2
  F::sub(H, H, q.X);                 // H = U2 - X1
  F::mul(r, q.Z, A);
  F::mul(r, y2, r);                  // S2
  F::sub(r, r, q.Y);
  const u32 h_zero = F::is_zero(H), r_zero = F::is_zero(r);
  F::mul2(r, r);                     // r = 2(S2 - Y1)
  F::add(D, q.Z, H);
  F::sqr(D, D);
  F::sub(D, D, A);
  F::sqr(A, H);                      // HH
  F::sub(q.Z, D, A);                 // Z3 = (Z1+H)^2 - Z1Z1 - HH
  F::mul4(A, A);                     // I = 4 HH
  F::mul(H, H, A);                   // J = H I
  F::mul(A, q.X, A);                 // V = X1 I
  F::sqr(q.X, r);
  F::sub(q.X, q.X, H);
  F::sub(q.X, q.X, A);
  F::sub(q.X, q.X, A);               // X3 = r^2 - J - 2V
  F::sub(A, A, q.X);
  F::mul(A, r, A);                   // r (V - X3)
  F::mul(H, q.Y, H);
  F::mul2(H, H);                     // 2 Y1 J
  F::sub(q.Y, A, H);
  return ((q_inf | h_zero) & 1u) | (q_inf & 2u) | (h_zero & r_zero & 4u);
}
// The same addition of (x2, neg_y ? -y2 : y2) with the zero tests behind a FILTER (r2): the exact flag word is computed only in a warp
// where some lane's Z1 or H passes a one-compare necessary condition (F::maybe_zero, 2^-31 per value); the sign of the addend joins
// the subtraction S2 - Y1 (one fold where the field has fused chains).  `any` is the caller's warp vote (lanes.cuh: warp_any).
template <class C, class Any>
FEC_DEV u32 jac_madd_lean_signed(Jac& q, const u32 x2[8], const u32 y2[8], u32 neg_y, Any any) {
  typedef typename C::F F;
  u32 A[8], H[8], r[8], D[8];
  const u32 sus_z = F::maybe_zero(q.Z);
  F::sqr(A, q.Z);                    // Z1Z1
  F::mul(H, x2, A);                  // U2
  F::sub(H, H, q.X);                 // H = U2 - X1
  F::mul(r, q.Z, A);
  F::mul(r, y2, r);                  // y2 Z1^3
  if constexpr (has_fused_linear<F>::value) {
    F::cneg_sub(r, r, neg_y, q.Y);   // +-y2 Z1^3 - Y1
  } else {
    u32 n[8];
    F::neg(n, r);
    select8(r, neg_y, n, r);
    F::sub(r, r, q.Y);
  }
  u32 flags = 0u;
  if (FEC_UNLIKELY(any(sus_z | F::maybe_zero(H)))) {
    const u32 q_inf = F::is_zero(q.Z), h_zero = F::is_zero(H), r_zero = F::is_zero(r);
    flags = ((q_inf | h_zero) & 1u) | (q_inf & 2u) | (h_zero & r_zero & 4u);
  }
#ifndef FEC_FIXED_MADD_7M4S
  if constexpr (has_fused_linear<F>::value) {
    // 8M + 3S with r = S2 - Y1 undoubled (r2): 5 linear operations instead of 14 for one multiplication more than a squaring
    F::mul(q.Z, q.Z, H);             // Z3 = Z1 H
    F::sqr(A, H);                    // HH
    F::mul(H, H, A);                 // HHH
    F::mul(A, q.X, A);               // V = X1 HH
    F::sqr(D, r);
    F::sub3(q.X, D, H, A);           // X3 = r^2 - HHH - 2V, one fold
    F::sub(A, A, q.X);
    F::mul(A, r, A);                 // r (V - X3)
    F::mul(H, q.Y, H);               // Y1 HHH
    F::sub(q.Y, A, H);
    return flags;
  }
#endif
  F::mul2(r, r);                     // r = 2(S2 - Y1)
  F::add(D, q.Z, H);
  F::sqr(D, D);
  F::sub(D, D, A);
  F::sqr(A, H);                      // HH
  F::sub(q.Z, D, A);                 // Z3 = (Z1+H)^2 - Z1Z1 - HH
  F::mul4(A, A);                     // I = 4 HH
  F::mul(H, H, A);                   // J = H I
  F::mul(A, q.X, A);                 // V = X1 I
  F::sqr(q.X, r);
  F::sub(q.X, q.X, H);
  F::sub(q.X, q.X, A);
  F::sub(q.X, q.X, A);               // X3 = r^2 - J - 2V
  F::sub(A, A, q.X);
  F::mul(A, r, A);                   // r (V - X3)
  F::mul(H, q.Y, H);
  F::mul2(H, H);                     // 2 Y1 J
  F::sub(q.Y, A, H);
  return flags;
}
// repair of the lanes jac_madd_lean flagged: infinity + T = T, T + T = 2T, T + (-T) = infinity
template <class C>
FEC_DEV void jac_madd_fix(Jac& q, const u32 x2[8], const u32 y2[8], u32 flags) {
  Jac d;
  copy8(d.X, x2); copy8(d.Y, y2); C::F::one(d.Z);
  Jac t = d;
  jac_dbl<C>(d);
  const u32 bad = 0u - (flags & 1u), was_inf = 0u - ((flags >> 1) & 1u), isdbl = (0u - ((flags >> 2) & 1u)) & ~was_inf;
  const u32 isinf = bad & ~was_inf & ~isdbl;
  u32 one[8], zero[8];
  C::F::one(one);
#pragma unroll
  for (int i = 0; i < 8; i++) zero[i] = 0u;
  select8(q.X, was_inf, t.X, q.X); select8(q.Y, was_inf, t.Y, q.Y); select8(q.Z, was_inf, t.Z, q.Z);
  select8(q.X, isdbl, d.X, q.X); select8(q.Y, isdbl, d.Y, q.Y); select8(q.Z, isdbl, d.Z, q.Z);
  select8(q.X, isinf, one, q.X); select8(q.Y, isinf, one, q.Y); select8(q.Z, isinf, zero, q.Z);
}

// Q <- Q + (x2,y2) with a skip mask, results selected by masks (used where the addend may be absent: corrections,
// table construction).  Returns bit0: exceptional (same x, both finite), bit1: also same y (needs doubling).
template <class C>
FEC_DEV u32 jac_madd(Jac& q, const u32 x2[8], const u32 y2[8], u32 t_inf) {
  typedef typename C::F F;
  u32 Z1Z1[8], U2[8], S2[8], H[8], HH[8], I[8], J[8], r[8], V[8], t[8];
  Jac o;
  u32 q_inf = F::is_zero(q.Z);
  F::sqr(Z1Z1, q.Z);
  F::mul(U2, x2, Z1Z1);
  F::mul(S2, q.Z, Z1Z1);
  F::mul(S2, y2, S2);
  F::sub(H, U2, q.X);
  F::sqr(HH, H);
  F::sub(r, S2, q.Y);
  u32 h_zero = F::is_zero(H);
  u32 r_zero = F::is_zero(r);
  F::mul4(I, HH);                    // I = 4 HH
  F::mul(J, H, I);
  F::mul2(r, r);                     // r = 2(S2 - Y1)
  F::mul(V, q.X, I);
  F::sqr(o.X, r);
  F::sub(o.X, o.X, J);
  F::sub(o.X, o.X, V);
  F::sub(o.X, o.X, V);               // X3 = r^2 - J - 2V
  F::sub(t, V, o.X);
  F::mul(t, r, t);
  F::mul(J, q.Y, J);
  F::mul2(J, J);
  F::sub(o.Y, t, J);                 // Y3 = r(V - X3) - 2 Y1 J
  F::add(t, q.Z, H);
  F::sqr(t, t);
  F::sub(t, t, Z1Z1);
  F::sub(o.Z, t, HH);                // Z3 = (Z1+H)^2 - Z1Z1 - HH
  u32 one[8];
  F::one(one);
  u32 keep = t_inf;                  // keep Q
  u32 take_t = q_inf & ~t_inf;       // Q was infinity: result is the addend
#pragma unroll
  for (int i = 0; i < 8; i++) {
    q.X[i] = (keep & q.X[i]) | (~keep & ((take_t & x2[i]) | (~take_t & o.X[i])));
    q.Y[i] = (keep & q.Y[i]) | (~keep & ((take_t & y2[i]) | (~take_t & o.Y[i])));
    q.Z[i] = (keep & q.Z[i]) | (~keep & ((take_t & one[i]) | (~take_t & o.Z[i])));
  }
  // exceptional: both finite and same x.  bit0: exceptional, bit1: also same y (needs doubling)
  u32 exc = h_zero & ~q_inf & ~t_inf;
  return (exc & 1u) | (exc & r_zero & 2u);
}

// P <- P + Q, both Jacobian: add-2007-bl (11M + 5S), every special case resolved by masks except P == Q, for which the
// returned mask is set and P is left unchanged (the caller doubles under a warp-uniform branch: jac_add_safe).
// Replaces `impl Add for ProjectivePoint` (secp256k1.rs:1447-1497, p256.rs:1941-2006) in reductions of many points.
template <class C>
FEC_DEV u32 jac_add(Jac& p, const Jac& q) {
  typedef typename C::F F;
  u32 Z1Z1[8], Z2Z2[8], U1[8], U2[8], S1[8], S2[8], H[8], I[8], J[8], r[8], V[8], t[8];
  Jac o;
  const u32 p_inf = F::is_zero(p.Z), q_inf = F::is_zero(q.Z);
  F::sqr(Z1Z1, p.Z);
  F::sqr(Z2Z2, q.Z);
  F::mul(U1, p.X, Z2Z2);
  F::mul(U2, q.X, Z1Z1);
  F::mul(S1, q.Z, Z2Z2);
  F::mul(S1, p.Y, S1);
  F::mul(S2, p.Z, Z1Z1);
  F::mul(S2, q.Y, S2);
  F::sub(H, U2, U1);
  F::sub(r, S2, S1);
  const u32 h_zero = F::is_zero(H), r_zero = F::is_zero(r);
  F::mul2(r, r);
  F::mul2(I, H);
  F::sqr(I, I);
  F::mul(J, H, I);
  F::mul(V, U1, I);
  F::sqr(o.X, r);
  F::sub(o.X, o.X, J);
  F::sub(o.X, o.X, V);
  F::sub(o.X, o.X, V);
  F::sub(t, V, o.X);
  F::mul(t, r, t);
  F::mul(J, S1, J);
  F::mul2(J, J);
  F::sub(o.Y, t, J);
  F::add(t, p.Z, q.Z);
  F::sqr(t, t);
  F::sub(t, t, Z1Z1);
  F::sub(t, t, Z2Z2);
  F::mul(o.Z, t, H);
  const u32 both = ~p_inf & ~q_inf;
  const u32 same = both & h_zero & r_zero;     // P == Q: doubling needed, keep P
  const u32 opp = both & h_zero & ~r_zero;     // P == -Q: infinity
  const u32 keep = q_inf | same;               // result is P
  const u32 take_q = p_inf & ~q_inf;           // result is Q
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const u32 x = (p.X[i] & keep) | (q.X[i] & take_q) | (o.X[i] & ~keep & ~take_q);
    const u32 y = (p.Y[i] & keep) | (q.Y[i] & take_q) | (o.Y[i] & ~keep & ~take_q);
    const u32 z = (p.Z[i] & keep) | (q.Z[i] & take_q) | (o.Z[i] & ~keep & ~take_q & ~opp);
    p.X[i] = x; p.Y[i] = y; p.Z[i] = z;
  }
  return same;
}

// y^2 == x^3 + a x + b  (mask)
template <class C>
FEC_DEV u32 on_curve(const u32 x[8], const u32 y[8]) {
  typedef typename C::F F;
  u32 l[8], r[8], b[8], t[8];
  F::sqr(l, y);
  F::sqr(r, x);
  F::mul(r, r, x);
  if (C::A_MINUS3) {
    F::dbl(t, x);
    F::add(t, t, x);
    F::sub(r, r, t);
  }
  C::B(b);
  F::add(r, r, b);
  F::sub(l, l, r);
  return F::is_zero(l);
}

// -------------------------------------------------------------------------------------
// Block-level simultaneous inversion (Montgomery's trick as a butterfly product tree):
// every thread of the block contributes one z != 0 and receives z^-1.  Up-sweep: 5 warp-shuffle
// levels + log2(warps) shared-memory levels, one multiplication each; ONE Fermat inversion per
// block (warp 0); down-sweep multiplies the stored sibling products back in.
// Cost per thread: 2*(5+log2 W) M  +  (1/W) of an inversion.   `sm` must hold 2*W*8 + 8 words.
// All threads of the block must call it (it contains __syncthreads()).
// -------------------------------------------------------------------------------------
#ifdef __CUDACC__
template <class F, int WARPS>
FEC_DEV void block_batch_inv(u32 zinv[8], const u32 z[8], u32* sm) {
  constexpr int LW = (WARPS == 1) ? 0 : (WARPS == 2) ? 1 : (WARPS == 4) ? 2 : (WARPS == 8) ? 3 : (WARPS == 16) ? 4 : -1;
  static_assert(LW >= 0, "WARPS must be a power of two <= 16");
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  u32 sib[5 + (LW > 0 ? LW : 1)][8];
  u32 p[8];
  copy8(p, z);
#pragma unroll
  for (int l = 0; l < 5; l++) {
#pragma unroll
    for (int i = 0; i < 8; i++) sib[l][i] = __shfl_xor_sync(0xFFFFFFFFu, p[i], 1 << l);
    F::mul(p, p, sib[l]);
  }
#pragma unroll
  for (int l = 0; l < LW; l++) {
    u32* buf = sm + (l & 1) * WARPS * 8;
    if (lane < 8) buf[warp * 8 + lane] = p[lane];
    __syncthreads();
    const unsigned partner = warp ^ (1u << l);
#pragma unroll
    for (int i = 0; i < 8; i++) sib[5 + l][i] = buf[partner * 8 + i];
    F::mul(p, p, sib[5 + l]);
    // (double-buffered: the next level writes the other half, the one after re-syncs first)
  }
  u32* slot = sm + 2 * WARPS * 8;
  __syncthreads();
  if (warp == 0) {  // warp-uniform
    u32 inv[8];
    F::inv(inv, p);
    if (lane < 8) slot[lane] = inv[lane];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; i++) p[i] = slot[i];
#pragma unroll
  for (int l = 5 + LW - 1; l >= 0; l--) F::mul(p, p, sib[l]);
  copy8(zinv, p);
}
// Warp-level simultaneous inversion: the 32 lanes of a warp each contribute one z != 0 and receive z^-1.
// 5 shuffle levels up, one Fermat inversion (computed redundantly by all lanes: no divergence, no barrier),
// 5 multiplications down.  Cost per lane: 10 M + one inversion per warp.  No shared memory, no __syncthreads.
template <class F>
FEC_DEV void warp_batch_inv(u32 zinv[8], const u32 z[8]) {
  u32 sib[5][8];
  u32 p[8];
  copy8(p, z);
#pragma unroll
  for (int l = 0; l < 5; l++) {
#pragma unroll
    for (int i = 0; i < 8; i++) sib[l][i] = __shfl_xor_sync(0xFFFFFFFFu, p[i], 1 << l);
    F::mul(p, p, sib[l]);
  }
  F::inv(p, p);
#pragma unroll
  for (int l = 4; l >= 0; l--) F::mul(p, p, sib[l]);
  copy8(zinv, p);
}
#endif  // __CUDACC__

// -------------------------------------------------------------------------------------
// Ed25519: extended coordinates (X:Y:Z:T), Niels addends (y+x, y-x, 2dxy)
// -------------------------------------------------------------------------------------
struct EdExt {
  u32 X[8], Y[8], Z[8], T[8];
};
FEC_DEV void ed_D2(u32 d2[8]) {  // 2d mod p, d = -121665/121666
  d2[0] = 0x26B2F159u; d2[1] = 0xEBD69B94u; d2[2] = 0x8283B156u; d2[3] = 0x00E0149Au;
  d2[4] = 0xEEF3D130u; d2[5] = 0x198E80F2u; d2[6] = 0x56DFFCE7u; d2[7] = 0x2406D9DCu;
}
FEC_DEV void ed_base(u32 x[8], u32 y[8]) {
  x[0] = 0x8F25D51Au; x[1] = 0xC9562D60u; x[2] = 0x9525A7B2u; x[3] = 0x692CC760u;
  x[4] = 0xFDD6DC5Cu; x[5] = 0xC0A4E231u; x[6] = 0xCD6E53FEu; x[7] = 0x216936D3u;
  y[0] = 0x66666658u; y[1] = 0x66666666u; y[2] = 0x66666666u; y[3] = 0x66666666u;
  y[4] = 0x66666666u; y[5] = 0x66666666u; y[6] = 0x66666666u; y[7] = 0x66666666u;
}
FEC_DEV void ed_set_identity(EdExt& q) {
#pragma unroll
  for (int i = 0; i < 8; i++) { q.X[i] = 0; q.Y[i] = 0; q.Z[i] = 0; q.T[i] = 0; }
  q.Y[0] = 1; q.Z[0] = 1;
}
// Q <- Q + N  (madd-2008-hwcd-3, 7M; complete: no exceptional cases, identity = (1,1,0))
// the same for the addend or its negative (neg: full mask).  -(y+x, y-x, 2dxy) = (y-x, y+x, -2dxy): the first two trade places
// as multipliers, and instead of negating the third, F = D - C and G = D + C trade places after it (32 selections instead of a
// conditional swap, a field negation and a selection)
FEC_DEV void ed_madd_signed(EdExt& q, const u32 ypx[8], const u32 ymx[8], const u32 t2d[8], u32 neg) {
  typedef Fe25519 F;
  u32 A[8], B[8], Cc[8], D[8], E[8], Fv[8], G[8], H[8], m[8];
#ifdef FEC_25519_FOLD255  // (k_ed25519.cu) X, Y, Z, T and A, B, C are results of multiplications: below 2^255 + 2^11 (fields.cuh)
  F::sub_r1(A, q.Y, q.X);
  select8(m, neg, ypx, ymx);
  F::mul(A, A, m);
  F::add_r1(B, q.Y, q.X);
  select8(m, neg, ymx, ypx);
  F::mul(B, B, m);
  F::mul(Cc, q.T, t2d);
  F::add_r1(D, q.Z, q.Z);          // 2Z: leaves 2^256 only for Z >= 2^255
  F::sub_r1(E, B, A);
  F::sub_r1(m, D, Cc);             // (sub_r1 needs only its SUBTRAHEND in the range)
  F::add(H, D, Cc);                // D is anywhere below 2^256: the general addition
  select8(Fv, neg, H, m);
  select8(G, neg, m, H);
  F::add_r1(H, B, A);
#else
  F::sub(A, q.Y, q.X);
  select8(m, neg, ypx, ymx);
  F::mul(A, A, m);
  F::add(B, q.Y, q.X);
  select8(m, neg, ymx, ypx);
  F::mul(B, B, m);
  F::mul(Cc, q.T, t2d);
  F::dbl(D, q.Z);
  F::sub(E, B, A);
  F::sub(m, D, Cc);
  F::add(H, D, Cc);
  select8(Fv, neg, H, m);
  select8(G, neg, m, H);
  F::add(H, B, A);
#endif
  F::mul(q.X, E, Fv);
  F::mul(q.Y, G, H);
  F::mul(q.T, E, H);
  F::mul(q.Z, Fv, G);
}
FEC_DEV void ed_madd(EdExt& q, const u32 ypx[8], const u32 ymx[8], const u32 t2d[8]) {
  typedef Fe25519 F;
  u32 A[8], B[8], Cc[8], D[8], E[8], Fv[8], G[8], H[8];
  F::sub(A, q.Y, q.X);
  F::mul(A, A, ymx);
  F::add(B, q.Y, q.X);
  F::mul(B, B, ypx);
  F::mul(Cc, q.T, t2d);
  F::dbl(D, q.Z);
  F::sub(E, B, A);
  F::sub(Fv, D, Cc);
  F::add(G, D, Cc);
  F::add(H, B, A);
  F::mul(q.X, E, Fv);
  F::mul(q.Y, G, H);
  F::mul(q.T, E, H);
  F::mul(q.Z, Fv, G);
}
// general extended addition with d2 (used only when building tables)
FEC_DEV void ed_add(EdExt& q, const EdExt& r) {
  typedef Fe25519 F;
  u32 A[8], B[8], Cc[8], D[8], E[8], Fv[8], G[8], H[8], t[8], d2[8];
  ed_D2(d2);
  F::sub(A, q.Y, q.X);
  F::sub(t, r.Y, r.X);
  F::mul(A, A, t);
  F::add(B, q.Y, q.X);
  F::add(t, r.Y, r.X);
  F::mul(B, B, t);
  F::mul(Cc, q.T, r.T);
  F::mul(Cc, Cc, d2);
  F::mul(D, q.Z, r.Z);
  F::dbl(D, D);
  F::sub(E, B, A);
  F::sub(Fv, D, Cc);
  F::add(G, D, Cc);
  F::add(H, B, A);
  F::mul(q.X, E, Fv);
  F::mul(q.Y, G, H);
  F::mul(q.T, E, H);
  F::mul(q.Z, Fv, G);
}

// -------------------------------------------------------------------------------------
// X25519 ladder step (RFC 7748 section 5; one bit): 5M + 4S + one multiplication by a24
// -------------------------------------------------------------------------------------
// State: D = (xd, zd) and S = (xs, zs), the doubling and the differential addition of the previous step; which of them is
// [m]P and which [m+1]P is the previous scalar bit.  The differential addition is symmetric in its two operands
// (DA + CB is, DA - CB changes sign and is squared), so RFC 7748's two conditional swaps of whole points shrink to the
// selection of the doubling's operand: sel = all-ones -> double S, else D.  (16 instead of 48 logic operations per bit.)
FEC_DEV void x25519_step(u32 xd[8], u32 zd[8], u32 xs[8], u32 zs[8], const u32 x1[8], u32 sel) {
  typedef Fe25519 F;
#ifdef FEC_25519_FOLD255  // every operand of the additions / subtractions below is a result of mul / sqr (or the masked u, 1, 0)
#define X_ADD F::add_r1
#define X_SUB F::sub_r1
#else
#define X_ADD F::add
#define X_SUB F::sub
#endif
  u32 A[8], AA[8], B[8], BB[8], E[8], Cc[8], D[8], DA[8], CB[8];
  X_ADD(A, xd, zd);
  X_SUB(B, xd, zd);
  X_ADD(Cc, xs, zs);
  X_SUB(D, xs, zs);
  F::mul(DA, D, A);
  F::mul(CB, Cc, B);
  X_ADD(xs, DA, CB);
  F::sqr(xs, xs);
  X_SUB(zs, DA, CB);
  F::sqr(zs, zs);
  F::mul(zs, x1, zs);
  select8(A, sel, Cc, A);
  select8(B, sel, D, B);
  F::sqr(AA, A);
  F::sqr(BB, B);
  X_SUB(E, AA, BB);
  F::mul(xd, AA, BB);
#ifndef FEC_X25519_NO_FUSED_A24
  F::mul_small_add(zd, E, 121665u, AA);  // AA + a24 E, one pass and one fold
#else
  F::mul_small(zd, E, 121665u);
  F::add(zd, AA, zd);
#endif
  F::mul(zd, E, zd);
#undef X_ADD
#undef X_SUB
}

}  // namespace fec
