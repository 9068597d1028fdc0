// x25519_vm.cuh -- the RFC 7748 ladder (x25519(), forge-ec-curves/src/curve25519.rs:1624-1716) on the accumulator
// machine of vm.cuh (host-compilable for tests/hostsim).
//
// Why it exists, and why it is NOT the default.  With x1, (x2, z2), (x3, z3) and the temporaries of a step in registers the
// ladder needs 128 registers per thread (16 warps per SM); here one field element is in registers and the lane's nine
// others are shared-memory slots, so the kernel runs 20 warps per SM at 96 registers like the Weierstrass ladders.
// Measured (B200, 2^22 lanes): 45.0 M/s against 48.9 M/s for the register-resident kernel -- this loop has no table and
// fits its registers, and with the field routines in line ptxas overlaps one product's reduction with the next product's
// rows, which calls through a single accumulator forbid.  Kept behind FORGE_EC_X_IMPL=vm for A/B and as the host-sim
// cross-check of the ladder (two independent formulations of the conditional swap must agree).
//
// Slot map of a lane: 0 = x1, 1..4 and 5..8 = two sets {x2, z2, x3, z3}.  A step reads set P and writes set Q, then the
// sets change roles (no copies).  The conditional swap of RFC 7748 costs no instruction: lane-private slots of the
// same half are a multiple of 128 bytes apart, so a lane reads "its" (x2, z2) from slots 0,1 or 2,3 of P by ADDRESS and
// every 128-bit access of the warp stays conflict-free.  (The addresses depend on scalar bits exactly as the table
// look-ups of the windowed kernels do; control flow stays warp-uniform.)
//
//   step:  A = x2 + z2, B = x2 - z2, C = x3 + z3, D = x3 - z3              -> P0..P3  (their inputs are dead)
//          DA = D A, CB = C B;  x3' = (DA + CB)^2;  z3' = x1 (DA - CB)^2
//          AA = A^2, BB = B^2;  x2' = AA BB;  E = AA - BB;  z2' = E (AA + a24 E)      5M + 4S + 1 small multiplication
#pragma once
#include "lanes.cuh"
#include "vm.cuh"

namespace fec {

enum { XS_X1 = 0, XS_SET0 = 1, XS_SET1 = 5, XS_COUNT = 9 };

// one ladder step: P (x2, z2, x3, z3; swapped per lane when off != 0) -> Q.  off = 0 or 2 * SLOT bytes.
template <int T>
FEC_DEV void x25519_vm_step(saddr X1, saddr P, saddr Q, int off) {
  typedef Fe25519 F;
  typedef Vm<F, T> V;
  const saddr P0 = P, P1 = V::slot(P, 1), P2 = V::slot(P, 2), P3 = V::slot(P, 3);
  fe a;
  {
    u32 xs[8], zs[8], xo[8], zo[8];
    V::ldv(xs, sa_add(P0, off)); V::ldv(zs, sa_add(P1, off));
    V::ldv(xo, sa_add(P2, -off)); V::ldv(zo, sa_add(P3, -off));
    fe t;
    F::add(t.v, xs, zs); V::st(P0, t);             // A
    F::sub(t.v, xs, zs); V::st(P1, t);             // B
    F::add(t.v, xo, zo); V::st(P2, t);             // C
    F::sub(a.v, xo, zo);                           // D
  }
  a = vm_mul<F, T>(a, P0); V::st(P3, a);           // DA
  a = V::ld(P2);
  a = vm_mul<F, T>(a, P1); V::st(P2, a);           // CB
  a = vm_add<F, T>(a, P3);
  a = vm_sqr<F>(a); V::st(V::slot(Q, 2), a);       // x3' = (DA + CB)^2
  a = V::ld(P3);
  a = vm_sub<F, T>(a, P2);
  a = vm_sqr<F>(a);
  a = vm_mul<F, T>(a, X1); V::st(V::slot(Q, 3), a);  // z3' = x1 (DA - CB)^2
  a = V::ld(P0);
  a = vm_sqr<F>(a); V::st(P0, a);                  // AA
  a = V::ld(P1);
  a = vm_sqr<F>(a); V::st(P1, a);                  // BB
  a = vm_mul<F, T>(a, P0); V::st(Q, a);            // x2' = AA BB
  a = V::ld(P0);
  a = vm_sub<F, T>(a, P1); V::st(P2, a);           // E
  {
    fe t;
    u32 aa[8];
    V::ldv(aa, P0);
    F::mul_small_add(t.v, a.v, 121665u, aa);       // AA + a24 E
    a = t;
  }
  a = vm_mul<F, T>(a, P2); V::st(V::slot(Q, 1), a);  // z2' = E (AA + a24 E)
}

// the whole ladder of one lane: (x2, z2) with X25519(k, u) = x2 / z2
template <int T>
FEC_DEV void x25519_vm_core(saddr S, u32 x2[8], u32 z2[8], const u32 k_le[8], const u32 u_le[8]) {
  typedef Fe25519 F;
  typedef Vm<F, T> V;
  u32 k[8];
  copy8(k, k_le);
  k[0] &= 0xFFFFFFF8u;
  k[7] = (k[7] & 0x7FFFFFFFu) | 0x40000000u;
  const saddr X1 = V::slot(S, XS_X1);
  saddr P = V::slot(S, XS_SET0), Q = V::slot(S, XS_SET1);
  {
    fe a;
    copy8(a.v, u_le);
    a.v[7] &= 0x7FFFFFFFu;
    V::st(X1, a); V::st(V::slot(P, 2), a);                   // x1, x3 = x1
    set_small(a.v, 1u);
    V::st(P, a); V::st(V::slot(P, 3), a);                    // x2 = 1, z3 = 1
    set_small(a.v, 0u);
    V::st(V::slot(P, 1), a);                                 // z2 = 0
  }
  // bits 254..0, consumed from the top of a shifting register (bit 255 is clear after clamping)
#pragma unroll
  for (int i = 7; i >= 1; i--) k[i] = funnel_l(k[i - 1], k[i], 1);
  k[0] <<= 1;
  u32 prev = 0u;
#pragma unroll 1
  for (int t = 254; t >= 0; t--) {
    const u32 kt = k[7] >> 31;
#pragma unroll
    for (int i = 7; i >= 1; i--) k[i] = funnel_l(k[i - 1], k[i], 1);
    k[0] <<= 1;
    const int off = (int)((prev ^ kt) * (u32)(2 * V::SLOT));
    prev = kt;
    x25519_vm_step<T>(X1, P, Q, off);
    const saddr tmp = P; P = Q; Q = tmp;
  }
  const int off = (int)(prev * (u32)(2 * V::SLOT));           // the final swap
  V::ldv(x2, sa_add(P, off));
  V::ldv(z2, sa_add(V::slot(P, 1), off));
}

}  // namespace fec
