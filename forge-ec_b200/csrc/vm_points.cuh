// vm_points.cuh -- point-level routines on the accumulator machine (vm.cuh) that need the curve layer (lanes.cuh):
// the rare-case repair of a mixed addition and the fixed-base loop.  Host-compilable for tests/hostsim.
#pragma once
#include "lanes.cuh"
#include "vm.cuh"

namespace fec {

// rare repair of a lane whose mixed addition hit Q = infinity or Q = +-T (warp-uniform branch in the caller)
template <class C, int T>
FEC_DEV void vm_madd_fix(saddr S, const u32* e, u32 neg_y, u32 flags) {
  typedef typename C::F F;
  typedef Vm<F, T> V;
  u32 tx[8], ty[8], ny[8];
  g_ld8(tx, e);
  g_ld8(ty, e + 8);
  F::neg(ny, ty);
  select8(ty, neg_y, ny, ty);
  Jac q;
  V::ldv(q.X, V::slot(S, VS_X)); V::ldv(q.Y, V::slot(S, VS_Y)); V::ldv(q.Z, V::slot(S, VS_Z));
  jac_madd_fix<C>(q, tx, ty, flags);
  fe t;
  copy8(t.v, q.X); V::st(V::slot(S, VS_X), t);
  copy8(t.v, q.Y); V::st(V::slot(S, VS_Y), t);
  copy8(t.v, q.Z); V::st(V::slot(S, VS_Z), t);
}

// fixed base on the accumulator machine: Q = sum of the COMB_WINDOWS signed odd 16-bit digits of K times the comb table;
// entry(i, idx) returns the address of (2 idx + 1) 2^(16 i) G as 16 words (x, y) -- global memory on the device, fetched
// with cp.async straight into the addend slots while the previous addition is still running.  On return the slots
// X, Y, Z hold the Jacobian result.  Same digits and additions as wmul_fixed_core (lanes.cuh), different residence:
// one field element in registers instead of a whole point and its temporaries (96 instead of 168 registers, 20 instead
// of 12 resident warps per SM, field routines out of line).
template <class C, int T, class Entry>
FEC_DEV void wmul_fixed_vm_core(saddr S, u32 K[8], const Entry& entry) {
  typedef typename C::F F;
  typedef Vm<F, T> V;
  u32 idx, sgn;
  next_digit_w<COMB_W>(K, idx, sgn);  // top digit: always positive
  vm_prefetch_xy<F, T>(S, entry(COMB_WINDOWS - 1, idx));
  cp_async_wait();
  {
    fe a = V::ld(V::slot(S, VS_TX));
    V::st(V::slot(S, VS_X), a);
    a = V::ld(V::slot(S, VS_TY));
    V::st(V::slot(S, VS_Y), a);
    F::one(a.v);
    V::st(V::slot(S, VS_Z), a);
  }
  next_digit_w<COMB_W>(K, idx, sgn);
  const u32* e = entry(COMB_WINDOWS - 2, idx);
  vm_prefetch_xy<F, T>(S, e);
#pragma unroll 1
  for (int i = COMB_WINDOWS - 2; i >= 0; i--) {
    const u32* cur = e;
    const u32 neg_y = sgn;
    cp_async_wait();
#ifndef FEC_VM_EXACT_FLAGS  // the zero tests of the addition behind a one-compare filter (vm.cuh), exact flags only in a flagged warp
    u32 fl = vm_jac_madd_1<F, T, C::ALU_BOUND, true>(S, neg_y);
    if (FEC_UNLIKELY(warp_any(fl))) fl = vm_madd_flags<F, T>(S);
#else
    const u32 fl = vm_jac_madd_1<F, T, C::ALU_BOUND>(S, neg_y);
#endif
    if (i > 0) {  // the addend slots are free again: fetch the next addend while this addition runs
      next_digit_w<COMB_W>(K, idx, sgn);
      e = entry(i - 1, idx);
      vm_prefetch_xy<F, T>(S, e);
    }
    vm_jac_madd_2<F, T, C::ALU_BOUND>(S);
    if (FEC_UNLIKELY(warp_any(fl))) vm_madd_fix<C, T>(S, cur, neg_y, fl);
  }
}

}  // namespace fec
