// glv3.cuh -- secp256k1 variable-base k*P as three stages with different register / occupancy needs
// (the per-lane work of k_glv_prep, k_glv_ladder, k_glv_finish in k_wmul_glv_k1.cu; host-compilable for tests/hostsim).
//
//   prepare : 8 odd multiples of P on the curve isomorphic by zs (where they are affine), each stored as
//             (x, beta x, y) so that the endomorphism costs nothing in the loop; GLV split and recoding of k.
//   ladder  : 31 x (4 doublings + 2 mixed additions) on the accumulator machine of vm.cuh (90 % of the work).
//   finish  : parity corrections, back to the original curve (Z *= zs).  Normalisation/encoding follow in the kernel.
//
// Replaces Secp256k1::multiply (forge-ec-curves/src/secp256k1.rs:2635-2692) for batches.
#pragma once
#include "lanes.cuh"
#include "vm.cuh"

namespace fec {

constexpr int GLV_TAB_WORDS = 8 * 24;  // per lane: 8 entries x (x, beta x, y)
enum { GLV_NEG1 = 1u, GLV_NEG2 = 2u, GLV_CORR1 = 4u, GLV_CORR2 = 8u };

// tab: u32[192] of this lane (global memory on the device).  KK: recoding registers of |k1|, |k2| (4 words each).
FEC_DEV void glv_prepare(u32* tab, u32 zs[8], u32 KK[8], u32& flags, const u32 k[8], const u32 px[8], const u32 py[8]) {
  typedef CurveK1 C;
  typedef FeK1 F;
  {
    Jac D, A;
    u32 C2[8], C3[8], s[8], beta[8];
    u32 zr[7][8];
    copy8(D.X, px); copy8(D.Y, py); set_small(D.Z, 1u);
    jac_dbl<C>(D);
    F::sqr(C2, D.Z);
    F::mul(C3, C2, D.Z);
    F::mul(A.X, px, C2);
    F::mul(A.Y, py, C3);
    g_st8(tab, A.X); g_st8(tab + 16, A.Y);
#pragma unroll 1
    for (int j = 1; j < 8; j++) {  // A = (2j+1) P and D = 2P kept at one common Z: co-Z additions, 4M + 2S each (lanes.cuh)
      coz_addu<C>(A.X, A.Y, D.X, D.Y, zr[j - 1]);
      g_st8(tab + 24 * j, A.X); g_st8(tab + 24 * j + 16, A.Y);
    }
    copy8(zs, D.Z);  // Z(2P) of the doubling; the product of the ratios joins it below
    k1_beta(beta);
    // entry 7 is already affine on the final curve; entry j < 7 is scaled by s = zr[j] .. zr[6]
    {
      u32 x[8], bx[8];
      g_ld8(x, tab + 24 * 7);
      F::mul(bx, x, beta);
      g_st8(tab + 24 * 7 + 8, bx);
    }
    copy8(s, zr[6]);
#pragma unroll 1
    for (int j = 6; j >= 0; j--) {
      u32 x[8], y[8], s2[8], bx[8];
      g_ld8(x, tab + 24 * j); g_ld8(y, tab + 24 * j + 16);
      F::sqr(s2, s);
      F::mul(x, x, s2);
      F::mul(s2, s2, s);
      F::mul(y, y, s2);
      F::mul(bx, x, beta);
      g_st8(tab + 24 * j, x); g_st8(tab + 24 * j + 8, bx); g_st8(tab + 24 * j + 16, y);
      if (j > 0) F::mul(s, s, zr[j - 1]);
    }
    F::mul(zs, zs, s);  // s is now the product of all seven ratios = Z(15 P) on the isomorphic curve
  }
  u32 s1[4], s2[4], neg1, neg2;
  glv_split(s1, neg1, s2, neg2, k);
  const u32 corr1 = (s1[0] & 1u) - 1u, corr2 = (s2[0] & 1u) - 1u;  // even -> use s + 1, subtract the base afterwards
  s1[0] |= 1u;
  s2[0] |= 1u;
  recode_reg4(KK, s1);
  recode_reg4(KK + 4, s2);
  flags = (neg1 & GLV_NEG1) | (neg2 & GLV_NEG2) | (corr1 & GLV_CORR1) | (corr2 & GLV_CORR2);
}

// start the copy of table entry idx of half h (0: (x, y), 1: (beta x, y)) into the addend slots
template <int T>
FEC_DEV void glv_prefetch(saddr S, const u32* tab, u32 idx, int h) {
  typedef Vm<FeK1, T> V;
  const u32* e = tab + idx * 24;
  const u32* x = e + (h ? 8 : 0);
  const saddr TX = V::slot(S, VS_TX), TY = V::slot(S, VS_TY);
  cp_async16(TX, x); cp_async16(sa_add(TX, V::HALF), x + 4);
  cp_async16(TY, e + 16); cp_async16(sa_add(TY, V::HALF), e + 20);
}

// rare repair of a lane whose mixed addition hit Q = infinity or Q = +-T (warp-uniform branch in the caller)
template <int T>
FEC_DEV void glv_madd_fix(saddr S, const u32* tab, u32 idx, int h, u32 neg_y, u32 flags) {
  typedef Vm<FeK1, T> V;
  typedef FeK1 F;
  u32 tx[8], ty[8], ny[8];
  g_ld8(tx, tab + idx * 24 + (h ? 8 : 0));
  g_ld8(ty, tab + idx * 24 + 16);
  F::neg(ny, ty);
  select8(ty, neg_y, ny, ty);
  Jac q;
  V::ldv(q.X, V::slot(S, VS_X)); V::ldv(q.Y, V::slot(S, VS_Y)); V::ldv(q.Z, V::slot(S, VS_Z));
  jac_madd_fix<CurveK1>(q, tx, ty, flags);
  fe t;
  copy8(t.v, q.X); V::st(V::slot(S, VS_X), t);
  copy8(t.v, q.Y); V::st(V::slot(S, VS_Y), t);
  copy8(t.v, q.Z); V::st(V::slot(S, VS_Z), t);
}

// the loop: on return slots X, Y, Z hold k1' P + k2' lambda P on the isomorphic curve (k' = recoded odd magnitudes)
template <int T>
FEC_DEV void glv_ladder(saddr S, const u32* tab, const u32 KK[8], u32 flags) {
  typedef FeK1 F;
  typedef Vm<F, T> V;
  u32 K1r[4], K2r[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { K1r[i] = KK[i]; K2r[i] = KK[4 + i]; }
  const u32 neg1 = 0u - (flags & 1u), neg2 = 0u - ((flags >> 1) & 1u);
  u32 idx, sgn;
  next_digit4(K1r, idx, sgn);
  glv_prefetch<T>(S, tab, idx, 0);
  cp_async_wait();
  {
    fe a = V::ld(V::slot(S, VS_TX));
    V::st(V::slot(S, VS_X), a);
    a = V::ld(V::slot(S, VS_TY));
    a = vm_cneg<F>(a, sgn ^ neg1);
    V::st(V::slot(S, VS_Y), a);
    set_small(a.v, 1u);
    V::st(V::slot(S, VS_Z), a);
  }
  next_digit4(K2r, idx, sgn);
  glv_prefetch<T>(S, tab, idx, 1);
#pragma unroll 1
  for (int i = 0; i < 32; i++) {
    if (i > 0) {
#pragma unroll 1
#ifndef FEC_K1_NO_LAZY_Z  // Z3 = Y Z in the four doublings and one Z <- 16 Z after them (r2: three folds fewer per window)
      for (int d = 0; d < 4; d++) vm_jac_dbl_a0<F, T, true>(S);
      vm_scale_z<F, T, 4>(S);
#else
      for (int d = 0; d < 4; d++) vm_jac_dbl_a0<F, T>(S);
#endif
    }
#pragma unroll 1
    for (int h = (i == 0 ? 1 : 0); h < 2; h++) {
      const u32 neg_y = sgn ^ (h == 0 ? neg1 : neg2);
      const u32 cur = idx;
      cp_async_wait();
#ifndef FEC_K1_MADD_7M4S  // the 8M + 3S form (8 linear operations instead of 14; r2: 59.7 -> 60.2 M/s); A/B: madd-2007-bl
      constexpr bool LEAN = true;
#else
      constexpr bool LEAN = false;
#endif
#ifndef FEC_K1_EXACT_FLAGS  // a two-compare filter on the fast path, the exact tests only in a warp it flags (r2)
      u32 fl = vm_jac_madd_1<F, T, LEAN, true>(S, neg_y);
      if (FEC_UNLIKELY(warp_any(fl))) fl = vm_madd_flags<F, T>(S);
#else
      const u32 fl = vm_jac_madd_1<F, T, LEAN>(S, neg_y);
#endif
      // the addend slots are free again: start fetching the next addend (none after the very last addition)
      if (!(i == 31 && h == 1)) {
        if (h == 0) next_digit4(K2r, idx, sgn); else next_digit4(K1r, idx, sgn);
        glv_prefetch<T>(S, tab, idx, h ^ 1);
      }
      vm_jac_madd_2<F, T, LEAN>(S);
      if (FEC_UNLIKELY(warp_any(fl))) glv_madd_fix<T>(S, tab, cur, h, neg_y, fl);
    }
  }
}

// The same loop with the point in REGISTERS and the field operations expanded in line (A/B against the accumulator machine,
// FORGE_EC_GLV_LADDER=regs): no slots, no calls, no shared memory -- the instruction scheduler sees whole point formulas, as in
// the X25519 step -- at the price of a loop body (one doubling, one mixed addition) of ~2.9 k instructions and ~150 registers.
// Table entries arrive by 128-bit loads one addition ahead of their use.
template <class CH>
FEC_DEV void glv_ladder_regs(Jac& Q, const u32* tab, const u32 KK[8], u32 flags) {
  typedef FeK1 F;
  u32 K1r[4], K2r[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { K1r[i] = KK[i]; K2r[i] = KK[4 + i]; }
  const u32 neg1 = 0u - (flags & 1u), neg2 = 0u - ((flags >> 1) & 1u);
  u32 idx, sgn, tx[8], ty[8];
  next_digit4(K1r, idx, sgn);
  g_ld8(Q.X, tab + idx * 24);
  g_ld8(ty, tab + idx * 24 + 16);
  {
    u32 ny[8];
    F::neg(ny, ty);
    select8(Q.Y, sgn ^ neg1, ny, ty);
  }
  set_small(Q.Z, 1u);
  next_digit4(K2r, idx, sgn);
  g_ld8(tx, tab + idx * 24 + 8);
  g_ld8(ty, tab + idx * 24 + 16);
#pragma unroll 1
  for (int i = 0; i < 32; i++) {
    if (i > 0) {
#pragma unroll 1
      for (int d = 0; d < 4; d++) jac_dbl<CH>(Q);
    }
#pragma unroll 1
    for (int h = (i == 0 ? 1 : 0); h < 2; h++) {
      const u32 neg_y = sgn ^ (h == 0 ? neg1 : neg2);
      u32 cx[8], cy[8];
      copy8(cx, tx);
      {
        u32 ny[8];
        F::neg(ny, ty);
        select8(cy, neg_y, ny, ty);
      }
      if (!(i == 31 && h == 1)) {  // the next addend: its loads overlap this addition
        if (h == 0) next_digit4(K2r, idx, sgn); else next_digit4(K1r, idx, sgn);
        g_ld8(tx, tab + idx * 24 + (h == 0 ? 8 : 0));
        g_ld8(ty, tab + idx * 24 + 16);
      }
      const u32 fl = jac_madd_lean<CH>(Q, cx, cy);
      if (FEC_UNLIKELY(warp_any(fl))) jac_madd_fix<CurveK1>(Q, cx, cy, fl);
    }
  }
}

// corrections for the scalars that were bumped to odd, then back to the original curve
FEC_DEV void glv_finish(Jac& Q, const u32* tab, const u32 zs[8], u32 flags) {
  typedef CurveK1 C;
  typedef FeK1 F;
  const u32 neg1 = 0u - (flags & 1u), neg2 = 0u - ((flags >> 1) & 1u);
  const u32 corr1 = 0u - ((flags >> 2) & 1u), corr2 = 0u - ((flags >> 3) & 1u);
#pragma unroll 1
  for (int h = 0; h < 2; h++) {
    u32 tx[8], ty[8], ny[8];
    g_ld8(tx, tab + (h ? 8 : 0));
    g_ld8(ty, tab + 16);
    F::neg(ny, ty);
    select8(ty, ~(h == 0 ? neg1 : neg2), ny, ty);
    jac_madd_safe_skip<C>(Q, tx, ty, ~(h == 0 ? corr1 : corr2));
  }
  F::mul(Q.Z, Q.Z, zs);
}

}  // namespace fec
