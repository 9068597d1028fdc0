// p256v3.cuh -- P-256 variable-base k*P as three stages (the per-lane work of k_p3_prep, k_p3_ladder, k_p3_finish in
// k_wmul_vm_p256.cu; host-compilable for tests/hostsim), the same split as glv3.cuh makes for secp256k1:
//
//   prepare : 8 affine odd multiples of P (built on the curve isomorphic by Z(2P), brought back with ONE inversion
//             shared by the lanes that call it together), regular signed recoding of k.
//   ladder  : 63 x (4 doublings a = -3, 4M+4S + 1 mixed addition, 8M+3S) on the accumulator machine of vm.cuh (95 % of
//             the work); the formulas are the ones with the fewest additions, because this field is ALU-bound.
//   finish  : conditional negation (even scalars were replaced by n - k).  Normalisation/encoding follow in the kernel.
//
// P-256 has no efficient endomorphism, so there is one 64-digit scalar and no isomorphic-curve loop (the a = -3
// doubling does not survive the isomorphism).  Replaces P256::multiply (forge-ec-curves/src/p256.rs:2120-2156).
#pragma once
#include "vm_points.cuh"

namespace fec {

constexpr int P3_TAB_WORDS = 8 * 16;  // per lane: 8 entries x (x, y)

// the lane's table in global memory behind the Tab interface of wmul_var_table
struct GlobalLaneTab {
  u32* e;
  FEC_DEV void store(int j, const u32 x[8], const u32 y[8]) { g_st8(e + 16 * j, x); g_st8(e + 16 * j + 8, y); }
  FEC_DEV void load(u32 j, u32 x[8], u32 y[8]) const { g_ld8(x, e + 16 * j); g_ld8(y, e + 16 * j + 8); }
};

// the loop: on return slots X, Y, Z hold k' P, k' the odd scalar behind the recoding register Kin
template <int T>
FEC_DEV void p3_ladder(saddr S, const u32* tab, const u32 Kin[8]) {
  typedef CurveP256 C;
  typedef FeP256 F;
  typedef Vm<F, T> V;
  u32 K[8];
  copy8(K, Kin);
  u32 idx, sgn;
  next_digit(K, idx, sgn);  // top digit: always positive
  vm_prefetch_xy<F, T>(S, tab + idx * 16);
  cp_async_wait();
  {
    fe a = V::ld(V::slot(S, VS_TX));
    V::st(V::slot(S, VS_X), a);
    a = V::ld(V::slot(S, VS_TY));
    V::st(V::slot(S, VS_Y), a);
    F::one(a.v);
    V::st(V::slot(S, VS_Z), a);
  }
  next_digit(K, idx, sgn);
  vm_prefetch_xy<F, T>(S, tab + idx * 16);
#pragma unroll 1
  for (int i = 0; i < 63; i++) {
#pragma unroll 1
    for (int d = 0; d < 4; d++) vm_jac_dbl_am3<F, T>(S);
    const u32 cur = idx, neg_y = sgn;
    cp_async_wait();
#ifndef FEC_VM_EXACT_FLAGS  // the zero tests of the addition behind a one-compare filter (vm.cuh), exact flags only in a flagged warp
    u32 fl = vm_jac_madd_1<F, T, C::ALU_BOUND, true>(S, neg_y);
    if (FEC_UNLIKELY(warp_any(fl))) fl = vm_madd_flags<F, T>(S);
#else
    const u32 fl = vm_jac_madd_1<F, T, C::ALU_BOUND>(S, neg_y);
#endif
    // the addend slots are free again: start fetching the next addend (none after the last addition)
    if (i < 62) {
      next_digit(K, idx, sgn);
      vm_prefetch_xy<F, T>(S, tab + idx * 16);
    }
    vm_jac_madd_2<F, T, C::ALU_BOUND>(S);
    if (FEC_UNLIKELY(warp_any(fl))) vm_madd_fix<C, T>(S, tab + cur * 16, neg_y, fl);
  }
}

}  // namespace fec
