// vm.cuh -- "accumulator machine" for long field-arithmetic programs (the doubling/addition loops).
//
// Why.  With whole points held in registers the scalar-multiplication loops need ~130-170 registers per thread
// (12-16 warps per SM, or spills), and with the field multiplications expanded in line the loop no longer fits the
// instruction cache (both measured).  Here a lane's field elements live in SHARED MEMORY
// slots and ONE element, the accumulator, lives in registers:
//
//     acc = ld(slot)  |  st(slot, acc)  |  acc = acc * [slot]  |  acc = acc^2  |  acc = acc +- [slot]  | ...
//
// Multiplication and squaring are out-of-line routines on the accumulator, the linear operations are expanded in
// line; operands arrive through the load/store pipe (which the integer pipes do not share) as 128-bit shared-memory
// accesses, the whole loop is ~2.2k instructions, and a thread needs 96 registers without spilling anything that is
// live in the loop, so 20 warps per SM are resident.
//
// Slot layout (per block of T threads): slot s = two "halves" of 4 words per thread,
//     byte address = base + ((2 s + h) T + thread) * 16        -> consecutive threads, consecutive 16-byte
// cells: every 128-bit access of a warp is conflict-free.
//
// On the host (tests/hostsim, g++) a slot address is a plain pointer and T = 1.
#pragma once
#include "fields.cuh"

namespace fec {

struct fe { u32 v[8]; };

#if defined(__CUDACC__)
typedef u32 saddr;  // byte address in the shared window
// FEC_VM_INLINE_MUL expands multiplication and squaring in line as well (short loop bodies: the fixed-base kernel)
#ifdef FEC_VM_INLINE_MUL
#define FEC_VMOP __device__ __forceinline__
#else
#define FEC_VMOP __device__ __noinline__
#endif
// the linear operations (~25 instructions each) are expanded in line: measured 3.5 % faster than calling them, and the
// loop body (~2.2 k instructions with them) still fits the instruction cache; define FEC_VM_CALL_LINEAR to call them
#ifdef FEC_VM_CALL_LINEAR
#define FEC_VMLIN __device__ __noinline__
#else
#define FEC_VMLIN __device__ __forceinline__
#endif
#define FEC_VMRARE __device__ __noinline__  // rare paths: kept out of the loop body
FEC_DEV void lds128(u32 r[4], saddr p) {
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(p) : "memory");
}
FEC_DEV void sts128(saddr p, const u32 r[4]) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
// 16 bytes global -> shared without passing through registers (completion: cp_async_wait)
FEC_DEV void cp_async16(saddr dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
FEC_DEV void cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
FEC_DEV saddr sa_add(saddr p, int bytes) { return p + (u32)bytes; }
#else
typedef u32* saddr;
#define FEC_VMOP inline
#define FEC_VMLIN inline
#define FEC_VMRARE inline
FEC_DEV void lds128(u32 r[4], saddr p) { for (int i = 0; i < 4; i++) r[i] = p[i]; }
FEC_DEV void sts128(saddr p, const u32 r[4]) { for (int i = 0; i < 4; i++) p[i] = r[i]; }
FEC_DEV void cp_async16(saddr dst, const void* src) { for (int i = 0; i < 4; i++) dst[i] = static_cast<const u32*>(src)[i]; }
FEC_DEV void cp_async_wait() {}
FEC_DEV saddr sa_add(saddr p, int bytes) { return p + bytes / 4; }
#endif

// 8 words <-> 32-byte aligned global memory (two 128-bit accesses on the device)
FEC_DEV void g_st8(u32* p, const u32 v[8]) {
#ifdef __CUDACC__
  reinterpret_cast<uint4*>(p)[0] = make_uint4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<uint4*>(p)[1] = make_uint4(v[4], v[5], v[6], v[7]);
#else
  copy8(p, v);
#endif
}
FEC_DEV void g_ld8(u32 v[8], const u32* p) {
#ifdef __CUDACC__
  const uint4 a = reinterpret_cast<const uint4*>(p)[0], b = reinterpret_cast<const uint4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#else
  copy8(v, p);
#endif
}

// F: field (FeK1, Fe25519, FeP256), T: threads per block (1 on the host)
template <class F, int T>
struct Vm {
  static constexpr int HALF = T * 16;      // bytes between the two halves of a slot
  static constexpr int SLOT = 2 * T * 16;  // bytes between slots
  FEC_DEV static saddr slot(saddr base, int s) { return sa_add(base, s * SLOT); }
  FEC_DEV static void ldv(u32 r[8], saddr p) { lds128(r, p); lds128(r + 4, sa_add(p, HALF)); }
  FEC_DEV static fe ld(saddr p) { fe a; ldv(a.v, p); return a; }
  FEC_DEV static void st(saddr p, const fe& a) { sts128(p, a.v); sts128(sa_add(p, HALF), a.v + 4); }
};

// ---- the out-of-line operations (one copy per kernel image) -----------------------------------------------
template <class F, int T> FEC_VMOP fe vm_mul(fe a, saddr p) {  // a * [p]
  u32 b[8];
  Vm<F, T>::ldv(b, p);
  fe r;
  F::mul_inl(r.v, a.v, b);
  return r;
}
template <class F> FEC_VMOP fe vm_sqr(fe a) {
  fe r;
  F::sqr_inl(r.v, a.v);
  return r;
}
// Two INDEPENDENT operations in one routine (FEC_VM_DUAL).  A warp executes in order: alone, a product
// is ~70 IMAD.WIDE followed by ~60 dependent additions, and only other warps fill the multiply pipe during the second phase.
// With two independent data flows in one straight-line routine ptxas interleaves the reduction of one with the rows of the
// other (as it does in the in-line X25519 step), so a single warp keeps both pipes busy.
struct fe2 { fe a, b; };
// two multiplications: the field's interleaved routine where it has one (FeP256::mul2_inl), else one after the other
template <class F> FEC_DEV void field_mul2(F, u32 o0[8], const u32 a0[8], const u32 b0[8], u32 o1[8], const u32 a1[8], const u32 b1[8]) {
  F::mul_inl(o0, a0, b0);
  F::mul_inl(o1, a1, b1);
}
FEC_DEV void field_mul2(FeP256, u32 o0[8], const u32 a0[8], const u32 b0[8], u32 o1[8], const u32 a1[8], const u32 b1[8]) {
  FeP256::mul2_inl(o0, a0, b0, o1, a1, b1);
}
template <class F, int T> FEC_VMOP fe2 vm_mul2(fe a0, saddr p0, fe a1, saddr p1) {  // a0 * [p0], a1 * [p1]
  u32 b0[8], b1[8];
  Vm<F, T>::ldv(b0, p0);
  Vm<F, T>::ldv(b1, p1);
  fe2 r;
  field_mul2(F(), r.a.v, a0.v, b0, r.b.v, a1.v, b1);
  return r;
}
template <class F> FEC_VMOP fe2 vm_sqr2(fe a0, fe a1) {  // a0^2, a1^2
  fe2 r;
  F::sqr_inl(r.a.v, a0.v);
  F::sqr_inl(r.b.v, a1.v);
  return r;
}
template <class F, int T> FEC_VMOP fe2 vm_mulsqr(fe a0, saddr p0, fe a1) {  // a0 * [p0], a1^2
  u32 b0[8];
  Vm<F, T>::ldv(b0, p0);
  fe2 r;
  F::mul_inl(r.a.v, a0.v, b0);
  F::sqr_inl(r.b.v, a1.v);
  return r;
}
template <class F, int T> FEC_VMLIN fe vm_add(fe a, saddr p) {  // a + [p]
  u32 b[8];
  Vm<F, T>::ldv(b, p);
  fe r;
  F::add(r.v, a.v, b);
  return r;
}
template <class F, int T> FEC_VMLIN fe vm_sub(fe a, saddr p) {  // a - [p]
  u32 b[8];
  Vm<F, T>::ldv(b, p);
  fe r;
  F::sub(r.v, a.v, b);
  return r;
}
template <class F, int T> FEC_VMLIN fe vm_rsub(fe a, saddr p) {  // [p] - a
  u32 b[8];
  Vm<F, T>::ldv(b, p);
  fe r;
  F::sub(r.v, b, a.v);
  return r;
}
// Chains of linear operations.  Fields that provide them (FeK1: HAS_FUSED_LINEAR) run a chain with ONE fold of the wrapped
// value (fields.cuh: fold_signed), the others as the operations they stand for.
template <class F, int T> FEC_VMLIN fe vm_sub2_dbl(fe a, saddr p, saddr q) {  // 2 (a - [p] - [q])
  u32 b[8], c[8];
  Vm<F, T>::ldv(b, p);
  Vm<F, T>::ldv(c, q);
  fe r;
  if constexpr (has_fused_linear<F>::value) {
    F::sub2_dbl(r.v, a.v, b, c);
  } else {
    F::sub(r.v, a.v, b); F::sub(r.v, r.v, c); F::mul2(r.v, r.v);
  }
  return r;
}
// (raw forms: fields.cuh -- [p] holds the unfolded value, its wrap count v travels in a register)
template <class F, int T> FEC_VMLIN fe vm_sub2_dbl_raw(fe a, saddr p, saddr q, u32& v) {
  u32 b[8], c[8];
  Vm<F, T>::ldv(b, p);
  Vm<F, T>::ldv(c, q);
  fe r;
  F::sub2_dbl_raw(r.v, v, a.v, b, c);
  return r;
}
template <class F, int T> FEC_VMLIN fe vm_sub_twice_raw(fe a, saddr p, u32 vp) {  // a - 2 ([p] + vp 2^256)
  u32 b[8];
  Vm<F, T>::ldv(b, p);
  fe r;
  F::sub_twice_raw(r.v, a.v, b, vp);
  return r;
}
template <class F, int T> FEC_VMLIN fe vm_rsub_raw(fe a, saddr p, u32 vp) {  // ([p] + vp 2^256) - a
  u32 b[8];
  Vm<F, T>::ldv(b, p);
  fe r;
  F::rsub_raw(r.v, b, vp, a.v);
  return r;
}
template <class F, int T> FEC_VMLIN fe vm_sub_twice(fe a, saddr p) {  // a - 2 [p]
  u32 b[8];
  Vm<F, T>::ldv(b, p);
  fe r;
  if constexpr (has_fused_linear<F>::value) {
    F::sub_twice(r.v, a.v, b);
  } else {
    F::sub(r.v, a.v, b); F::sub(r.v, r.v, b);
  }
  return r;
}
template <class F, int T> FEC_VMLIN fe vm_sub3(fe a, saddr p, saddr q) {  // a - [p] - 2 [q]
  u32 b[8], c[8];
  Vm<F, T>::ldv(b, p);
  Vm<F, T>::ldv(c, q);
  fe r;
  if constexpr (has_fused_linear<F>::value) {
    F::sub3(r.v, a.v, b, c);
  } else {
    F::sub(r.v, a.v, b); F::sub(r.v, r.v, c); F::sub(r.v, r.v, c);
  }
  return r;
}
template <class F, int T> FEC_VMLIN fe vm_sub_mul8(fe a, saddr p) {  // a - 8 [p]
  u32 b[8];
  Vm<F, T>::ldv(b, p);
  fe r;
  if constexpr (has_fused_linear<F>::value) {
    F::sub_mul8(r.v, a.v, b);
  } else {
    u32 t[8];
    F::mul8(t, b); F::sub(r.v, a.v, t);
  }
  return r;
}
#ifndef FEC_VM_NO_RELOAD
#define VM_RELOAD(a, slot) a = V::ld(slot)
#else
#define VM_RELOAD(a, slot)
#endif
template <class F, int T> FEC_VMLIN fe vm_cneg_sub(fe a, u32 m, saddr p) {  // (m ? -a : a) - [p]
  u32 b[8];
  Vm<F, T>::ldv(b, p);
  fe r;
  if constexpr (has_fused_linear<F>::value) {
    F::cneg_sub(r.v, a.v, m, b);
  } else {
    u32 n[8], t[8];
    F::neg(n, a.v);
    select8(t, m, n, a.v);
    F::sub(r.v, t, b);
  }
  return r;
}
template <class F, int K> FEC_VMLIN fe vm_small(fe a) {  // a * K, K in {2,3,4,8}
  fe r;
  if (K == 2) F::mul2(r.v, a.v);
  else if (K == 3) F::mul3(r.v, a.v);
  else if (K == 4) F::mul4(r.v, a.v);
  else F::mul8(r.v, a.v);
  return r;
}
template <class F, int T, int SH> FEC_VMLIN void vm_scale_z(saddr S) {  // Z <- 2^SH Z: the factors Z_LAZY doublings left out
  typedef Vm<F, T> V;
  fe a = V::ld(V::slot(S, 2));
  fe r;
  F::template mul_pow2<SH>(r.v, a.v);
  V::st(V::slot(S, 2), r);
}
template <class F> FEC_VMLIN fe vm_half(fe a) {  // a / 2 (fields that provide it)
  fe r;
  F::half(r.v, a.v);
  return r;
}
template <class F> FEC_VMLIN fe vm_cneg(fe a, u32 m) {  // m ? -a : a
  fe r;
  u32 n[8];
  F::neg(n, a.v);
  select8(r.v, m, n, a.v);
  return r;
}

// -------------------------------------------------------------------------------------------------------------
// Jacobian programs for a = 0 curves (secp256k1).  Slot map of a lane:
//   0 X   1 Y   2 Z   3..6 temporaries   7 TX   8 TY  (the prefetched affine addend)
// -------------------------------------------------------------------------------------------------------------
enum { VS_X = 0, VS_Y = 1, VS_Z = 2, VS_A = 3, VS_B = 4, VS_C = 5, VS_D = 6, VS_TX = 7, VS_TY = 8, VS_COUNT = 9 };

// Q <- 2Q  (dbl-2009-l, 2M + 5S), in place
// Z_LAZY: Z3 = Y Z without the factor 2 -- the caller owes Z a factor 2 per doubling and pays the run of them at once
// (vm_scale_z); nothing else in a doubling reads Z.
template <class F, int T, bool Z_LAZY = false>
FEC_DEV void vm_jac_dbl_a0(saddr S) {
  typedef Vm<F, T> V;
  const saddr X = V::slot(S, VS_X), Y = V::slot(S, VS_Y), Z = V::slot(S, VS_Z), A = V::slot(S, VS_A), B = V::slot(S, VS_B),
              Cs = V::slot(S, VS_C), D = V::slot(S, VS_D);
#ifdef FEC_VM_DUAL
  {  // the same 2M + 5S as two paired squarings, a multiplication next to a squaring, and one multiplication
    fe a0 = V::ld(X), a1 = V::ld(Y);
    fe2 s = vm_sqr2<F>(a0, a1);
    V::st(A, s.a); V::st(B, s.b);                  // A = X^2, B = Y^2
    a1 = V::ld(X);
    a1 = vm_add<F, T>(a1, B);
    fe2 t = vm_sqr2<F>(s.b, a1);
    V::st(Cs, t.a);                                // C = B^2
    a1 = vm_sub<F, T>(t.b, A);
    a1 = vm_sub<F, T>(a1, Cs);
    a1 = vm_small<F, 2>(a1); V::st(D, a1);         // D = 2((X+B)^2 - A - C)
    a1 = V::ld(A);
    a1 = vm_small<F, 3>(a1); V::st(A, a1);         // E = 3A
    a0 = V::ld(Y);
    fe2 u = vm_mulsqr<F, T>(a0, Z, a1);            // Y Z, E^2
    a0 = u.a;
    if (!Z_LAZY) a0 = vm_small<F, 2>(a0);
    V::st(Z, a0);                                  // Z3 = 2YZ
    a1 = vm_sub<F, T>(u.b, D);
    a1 = vm_sub<F, T>(a1, D); V::st(X, a1);        // X3 = E^2 - 2D
    a1 = vm_rsub<F, T>(a1, D);                     // D - X3
    a1 = vm_mul<F, T>(a1, A);                      // E (D - X3)
    a0 = V::ld(Cs);
    a0 = vm_small<F, 8>(a0);
    fe y;
    F::sub(y.v, a1.v, a0.v); V::st(Y, y);          // Y3 = E(D - X3) - 8C
    return;
  }
#endif
#ifdef FEC_K1_DBL_3M4S
  // A/B variant: D = 4 X B by a multiplication (3M + 4S, 8 linear operations instead of 11)
  {
    fe a = V::ld(Y);
    a = vm_sqr<F>(a); V::st(B, a);                 // B = Y^2
    a = vm_sqr<F>(a); V::st(Cs, a);                // C = B^2
    a = V::ld(B);
    a = vm_mul<F, T>(a, X);
    a = vm_small<F, 4>(a); V::st(D, a);            // D = 4 X B
    a = V::ld(X);
    a = vm_sqr<F>(a);
    a = vm_small<F, 3>(a); V::st(A, a);            // E = 3 X^2
    a = vm_sqr<F>(a);
    a = vm_sub<F, T>(a, D);
    a = vm_sub<F, T>(a, D); V::st(X, a);           // X3 = E^2 - 2D
    a = vm_rsub<F, T>(a, D);
    a = vm_mul<F, T>(a, A); V::st(B, a);           // E (D - X3)
    a = V::ld(Y);
    a = vm_mul<F, T>(a, Z);
    if (!Z_LAZY) a = vm_small<F, 2>(a);
    V::st(Z, a);                                   // Z3 = 2YZ
    a = V::ld(Cs);
    a = vm_small<F, 8>(a);
    a = vm_rsub<F, T>(a, B); V::st(Y, a);          // Y3 = E(D - X3) - 8C
    return;
  }
#endif
  // (a routine returns its result in other registers than it takes its argument in: feeding one straight into the next costs
  // 8 moves, reading the slot just written costs 2 loads -- VM_RELOAD)
  fe a = V::ld(Y);
  a = vm_sqr<F>(a); V::st(B, a);                 // B = Y^2
  VM_RELOAD(a, B);
  a = vm_sqr<F>(a); V::st(Cs, a);                // C = B^2
  a = V::ld(X);
  a = vm_sqr<F>(a); V::st(A, a);                 // A = X^2
  a = V::ld(X);
  a = vm_add<F, T>(a, B);
  a = vm_sqr<F>(a);
#if !defined(FEC_VM_NO_FUSED_LINEAR) && !defined(FEC_VM_FOLD_D)
  u32 vD = 0u;
  if constexpr (has_fused_linear<F>::value) {
    a = vm_sub2_dbl_raw<F, T>(a, A, Cs, vD);     // D = 2((X+B)^2 - A - C) UNFOLDED: it only enters the two chains below
  } else {
    a = vm_sub2_dbl<F, T>(a, A, Cs);
  }
  V::st(D, a);
#elif !defined(FEC_VM_NO_FUSED_LINEAR)
  a = vm_sub2_dbl<F, T>(a, A, Cs); V::st(D, a);  // D = 2((X+B)^2 - A - C), one fold
#else
  a = vm_sub<F, T>(a, A);
  a = vm_sub<F, T>(a, Cs);
  a = vm_small<F, 2>(a); V::st(D, a);            // D = 2((X+B)^2 - A - C)
#endif
  a = V::ld(Y);
  a = vm_mul<F, T>(a, Z);
  if (!Z_LAZY) a = vm_small<F, 2>(a);
  V::st(Z, a);                                   // Z3 = 2YZ
  a = V::ld(A);
  a = vm_small<F, 3>(a); V::st(A, a);            // E = 3A
  a = vm_sqr<F>(a);
#if !defined(FEC_VM_NO_FUSED_LINEAR) && !defined(FEC_VM_FOLD_D)
  if constexpr (has_fused_linear<F>::value) {
    a = vm_sub_twice_raw<F, T>(a, D, vD); V::st(X, a);  // X3 = E^2 - 2D
    a = vm_rsub_raw<F, T>(a, D, vD);                    // D - X3
  } else {
    a = vm_sub_twice<F, T>(a, D); V::st(X, a);
    a = vm_rsub<F, T>(a, D);
  }
  a = vm_mul<F, T>(a, A);                        // E (D - X3)
  a = vm_sub_mul8<F, T>(a, Cs); V::st(Y, a);     // Y3 = E(D - X3) - 8C, one fold, nothing parked in a slot
#elif !defined(FEC_VM_NO_FUSED_LINEAR)
  a = vm_sub_twice<F, T>(a, D); V::st(X, a);     // X3 = E^2 - 2D, one fold
  a = vm_rsub<F, T>(a, D);                       // D - X3
  a = vm_mul<F, T>(a, A);                        // E (D - X3)
  a = vm_sub_mul8<F, T>(a, Cs); V::st(Y, a);     // Y3 = E(D - X3) - 8C, one fold, nothing parked in a slot
#else
  a = vm_sub<F, T>(a, D);
  a = vm_sub<F, T>(a, D); V::st(X, a);           // X3 = E^2 - 2D
  a = vm_rsub<F, T>(a, D);                       // D - X3
  a = vm_mul<F, T>(a, A); V::st(B, a);           // E (D - X3)
  a = V::ld(Cs);
  a = vm_small<F, 8>(a);
  a = vm_rsub<F, T>(a, B); V::st(Y, a);          // Y3 = E(D - X3) - 8C
#endif
}

// Q <- 2Q for a = -3 curves (P-256), in place: 4M + 4S with 2Y carried through (see jac_dbl in curves.cuh) -- the
// canonical-form additions of P-256 run on the ALU pipe, which bounds this field, so they are traded for a multiplication
template <class F, int T>
FEC_DEV void vm_jac_dbl_am3(saddr S) {
  typedef Vm<F, T> V;
  const saddr X = V::slot(S, VS_X), Y = V::slot(S, VS_Y), Z = V::slot(S, VS_Z), A = V::slot(S, VS_A), B = V::slot(S, VS_B),
              Cs = V::slot(S, VS_C), D = V::slot(S, VS_D);
#ifdef FEC_VM_DUAL
  {  // alpha's product next to Z3 = 2Y Z (one interleaved routine); the rest as below
    fe a = V::ld(Z);
    a = vm_sqr<F>(a); V::st(A, a);                 // delta = Z^2
    a = vm_rsub<F, T>(a, X); V::st(D, a);          // X - delta
    fe a0 = V::ld(X);
    a0 = vm_add<F, T>(a0, A);
    fe a1 = V::ld(Y);
    a1 = vm_small<F, 2>(a1); V::st(B, a1);         // 2Y
    fe2 m = vm_mul2<F, T>(a0, D, a1, Z);           // (X + delta)(X - delta), 2Y Z
    V::st(Z, m.b);                                 // Z3
    a = vm_small<F, 3>(m.a); V::st(D, a);          // alpha
    a = V::ld(B);
    a = vm_sqr<F>(a); V::st(B, a);                 // 4 gamma
    a = vm_mul<F, T>(a, X); V::st(Cs, a);          // 4 beta
    a = V::ld(D);
    a = vm_sqr<F>(a);
    a = vm_sub<F, T>(a, Cs);
    a = vm_sub<F, T>(a, Cs); V::st(X, a);          // X3 = alpha^2 - 8 beta
    a = vm_rsub<F, T>(a, Cs);                      // 4 beta - X3
    a = vm_mul<F, T>(a, D); V::st(A, a);           // alpha (4 beta - X3)
    a = V::ld(B);
    a = vm_sqr<F>(a);
    a = vm_half<F>(a);                             // 8 gamma^2
    a = vm_rsub<F, T>(a, A); V::st(Y, a);          // Y3
    return;
  }
#endif
  fe a = V::ld(Z);
  a = vm_sqr<F>(a); V::st(A, a);                 // delta = Z^2
  a = vm_rsub<F, T>(a, X); V::st(D, a);          // X - delta
  a = V::ld(X);
  a = vm_add<F, T>(a, A);
  a = vm_mul<F, T>(a, D);
  a = vm_small<F, 3>(a); V::st(D, a);            // alpha = 3 (X - delta)(X + delta)
  a = V::ld(Y);
  a = vm_small<F, 2>(a); V::st(B, a);            // 2Y
  a = vm_mul<F, T>(a, Z); V::st(Z, a);           // Z3 = 2Y Z
  a = V::ld(B);
  a = vm_sqr<F>(a); V::st(B, a);                 // 4 gamma
  VM_RELOAD(a, B);
  a = vm_mul<F, T>(a, X); V::st(Cs, a);          // 4 beta
  a = V::ld(D);
  a = vm_sqr<F>(a);
  a = vm_sub<F, T>(a, Cs);
  a = vm_sub<F, T>(a, Cs); V::st(X, a);          // X3 = alpha^2 - 8 beta
  a = vm_rsub<F, T>(a, Cs);                      // 4 beta - X3
  a = vm_mul<F, T>(a, D); V::st(A, a);           // alpha (4 beta - X3)
  a = V::ld(B);
  a = vm_sqr<F>(a);
  a = vm_half<F>(a);                             // 8 gamma^2
  a = vm_rsub<F, T>(a, A); V::st(Y, a);          // Y3 = alpha (4 beta - X3) - 8 gamma^2
}

// start the copy of an affine table entry (x at e, y at e + 8 words, global memory) into the addend slots
template <class F, int T>
FEC_DEV void vm_prefetch_xy(saddr S, const u32* e) {
  typedef Vm<F, T> V;
  const saddr TX = V::slot(S, VS_TX), TY = V::slot(S, VS_TY);
  cp_async16(TX, e); cp_async16(sa_add(TX, V::HALF), e + 4);
  cp_async16(TY, e + 8); cp_async16(sa_add(TY, V::HALF), e + 12);
}

// Q <- Q + (TX, +-TY), madd-2007-bl (7M + 4S), in place, first half: up to r = 2(S2 - Y1).  After it the
// addend slots are no longer needed (the caller may start the next prefetch into them).
// Returns the flag word of jac_madd_lean: bit0 formula does not apply, bit1 Q was at infinity, bit2 doubling.
// SUBS_ONLY (fields bound by the ALU pipe): the 8M + 3S form of jac_madd_lean, r = S2 - Y1 undoubled.
template <class F, int T, bool SUBS_ONLY = false, bool FILTER = false>
FEC_DEV u32 vm_jac_madd_1(saddr S, u32 neg_y) {
  typedef Vm<F, T> V;
  const saddr X = V::slot(S, VS_X), Y = V::slot(S, VS_Y), Z = V::slot(S, VS_Z), A = V::slot(S, VS_A), H = V::slot(S, VS_B),
              R = V::slot(S, VS_C), TX = V::slot(S, VS_TX), TY = V::slot(S, VS_TY);
#ifdef FEC_VM_DUAL
  if (!SUBS_ONLY) {  // (secp256k1 form) Z1Z1; U2 next to Z1^3; S2 next to HH
    fe a = V::ld(Z);
    const u32 q_inf = F::is_zero(a.v);
    a = vm_sqr<F>(a); V::st(A, a);                 // Z1Z1
    fe2 m = vm_mul2<F, T>(a, TX, a, Z);            // U2 = x2 Z1Z1, Z1^3
    fe h = vm_sub<F, T>(m.a, X); V::st(H, h);      // H = U2 - X1
    const u32 h_zero = F::is_zero(h.v);
    fe2 n = vm_mulsqr<F, T>(m.b, TY, h);           // y2 Z1^3, HH
    V::st(V::slot(S, VS_D), n.b);                  // HH
    a = vm_cneg<F>(n.a, neg_y);                    // S2 = +-y2 Z1^3
    a = vm_sub<F, T>(a, Y);
    const u32 r_zero = F::is_zero(a.v);
    a = vm_small<F, 2>(a);
    V::st(R, a);                                   // r = 2(S2 - Y1)
    return ((q_inf | h_zero) & 1u) | (q_inf & 2u) | (h_zero & r_zero & 4u);
  } else {  // (P-256 form) U2 next to Z1^3 in one interleaved routine
    fe a = V::ld(Z);
    const u32 q_inf = F::is_zero(a.v);
    a = vm_sqr<F>(a); V::st(A, a);                 // Z1Z1
    fe2 m = vm_mul2<F, T>(a, TX, a, Z);            // U2 = x2 Z1Z1, Z1^3
    fe h = vm_sub<F, T>(m.a, X); V::st(H, h);      // H = U2 - X1
    const u32 h_zero = F::is_zero(h.v);
    a = vm_mul<F, T>(m.b, TY);
    a = vm_cneg<F>(a, neg_y);                      // S2 = +-y2 Z1^3
    a = vm_sub<F, T>(a, Y);
    const u32 r_zero = F::is_zero(a.v);
    V::st(R, a);                                   // r = S2 - Y1
    return ((q_inf | h_zero) & 1u) | (q_inf & 2u) | (h_zero & r_zero & 4u);
  }
#endif
  fe a = V::ld(Z);
  // FILTER: a two-compare necessary condition instead of the three exact tests (a value that is 0 mod p has its top limb all
  // zeros or all ones); when any lane of the warp passes it the caller gets the exact word from vm_madd_flags (the slots Z, H,
  // R still hold what the tests read) -- a false alarm costs that warp ~60 instructions once in ~2^26 additions
  u32 q_inf;
  if (FILTER) q_inf = F::maybe_zero(a.v); else q_inf = F::is_zero(a.v);
  a = vm_sqr<F>(a); V::st(A, a);                 // Z1Z1
  VM_RELOAD(a, A);
  a = vm_mul<F, T>(a, TX);
  a = vm_sub<F, T>(a, X); V::st(H, a);           // H = U2 - X1
  u32 h_zero;
  if (FILTER) h_zero = F::maybe_zero(a.v); else h_zero = F::is_zero(a.v);
  a = V::ld(Z);
  a = vm_mul<F, T>(a, A);
#ifndef FEC_VM_NO_RELOAD
  if (SUBS_ONLY) {
    V::st(A, a);                                 // Z1^3 through the slot of Z1Z1 (dead from here in the 8M + 3S form)
    a = V::ld(TY);
    a = vm_mul<F, T>(a, A);
  } else
#endif
    a = vm_mul<F, T>(a, TY);
#ifndef FEC_VM_NO_FUSED_LINEAR
  a = vm_cneg_sub<F, T>(a, neg_y, Y);            // +-y2 Z1^3 - Y1, negation and subtraction under one fold
#else
  a = vm_cneg<F>(a, neg_y);                      // S2 = +-y2 Z1^3
  a = vm_sub<F, T>(a, Y);
#endif
  if (FILTER) {
    if (!SUBS_ONLY) a = vm_small<F, 2>(a);
    V::st(R, a);
#ifdef FEC_VM_FILTER_ALWAYS  // test builds: every addition takes the rare path (tests/test_host.py)
    return 1u;
#else
    return q_inf | h_zero;
#endif
  }
  const u32 r_zero = F::is_zero(a.v);
  if (!SUBS_ONLY) a = vm_small<F, 2>(a);
  V::st(R, a);                                   // r = 2(S2 - Y1), or S2 - Y1
  return ((q_inf | h_zero) & 1u) | (q_inf & 2u) | (h_zero & r_zero & 4u);
}
// the exact flag word of vm_jac_madd_1, from the slots it left (Z1, H, r): the rare path behind its FILTER
template <class F, int T>
FEC_VMRARE u32 vm_madd_flags(saddr S) {
  typedef Vm<F, T> V;
  u32 t[8];
  V::ldv(t, V::slot(S, VS_Z));
  const u32 q_inf = F::is_zero(t);
  V::ldv(t, V::slot(S, VS_B));
  const u32 h_zero = F::is_zero(t);
  V::ldv(t, V::slot(S, VS_C));
  const u32 r_zero = F::is_zero(t);
  return ((q_inf | h_zero) & 1u) | (q_inf & 2u) | (h_zero & r_zero & 4u);
}
template <class F, int T, bool SUBS_ONLY = false>
FEC_DEV void vm_jac_madd_2(saddr S) {
  typedef Vm<F, T> V;
  const saddr X = V::slot(S, VS_X), Y = V::slot(S, VS_Y), Z = V::slot(S, VS_Z), A = V::slot(S, VS_A), H = V::slot(S, VS_B),
              R = V::slot(S, VS_C), HH = V::slot(S, VS_D);
#ifdef FEC_VM_DUAL
  if (SUBS_ONLY) {  // Z3 next to HHH, r (V - X3) next to Y1 HHH (interleaved routines)
    fe a = V::ld(H);
    a = vm_sqr<F>(a); V::st(HH, a);                // HH
    fe a0 = V::ld(Z);
    fe2 m = vm_mul2<F, T>(a0, H, a, H);            // Z3 = Z1 H, HHH
    V::st(Z, m.a); V::st(H, m.b);
    a = V::ld(X);
    a = vm_mul<F, T>(a, HH); V::st(A, a);          // V = X1 HH
    a = V::ld(R);
    a = vm_sqr<F>(a);
    a = vm_sub<F, T>(a, H);
    a = vm_sub<F, T>(a, A);
    a = vm_sub<F, T>(a, A); V::st(X, a);           // X3 = r^2 - HHH - 2V
    a = vm_rsub<F, T>(a, A);                       // V - X3
    fe2 w = vm_mul2<F, T>(a, R, m.b, Y);           // r (V - X3), Y1 HHH
    fe y;
    F::sub(y.v, w.a.v, w.b.v); V::st(Y, y);        // Y3
    return;
  }
#endif
  if (SUBS_ONLY) {
    fe a = V::ld(Z);
    a = vm_mul<F, T>(a, H); V::st(Z, a);           // Z3 = Z1 H
    a = V::ld(H);
    a = vm_sqr<F>(a); V::st(HH, a);                // HH
    VM_RELOAD(a, HH);
    a = vm_mul<F, T>(a, H); V::st(H, a);           // HHH
    a = V::ld(X);
    a = vm_mul<F, T>(a, HH); V::st(A, a);          // V = X1 HH
    a = V::ld(R);
    a = vm_sqr<F>(a);
#ifndef FEC_VM_NO_FUSED_LINEAR
    a = vm_sub3<F, T>(a, H, A); V::st(X, a);       // X3 = r^2 - HHH - 2V, one fold where the field has it
#else
    a = vm_sub<F, T>(a, H);
    a = vm_sub<F, T>(a, A);
    a = vm_sub<F, T>(a, A); V::st(X, a);           // X3 = r^2 - HHH - 2V
#endif
    a = vm_rsub<F, T>(a, A);                       // V - X3
    a = vm_mul<F, T>(a, R); V::st(R, a);           // r (V - X3)
    a = V::ld(Y);
    a = vm_mul<F, T>(a, H);                        // Y1 HHH
    a = vm_rsub<F, T>(a, R); V::st(Y, a);          // Y3
    return;
  }
#ifdef FEC_VM_DUAL
  {  // HH is in its slot already (vm_jac_madd_1);  (Z1+H)^2 next to r^2;  J next to V;  r (V - X3) next to Y1 J
    fe a0 = V::ld(Z);
    a0 = vm_add<F, T>(a0, H);
    fe a1 = V::ld(R);
    fe2 q = vm_sqr2<F>(a0, a1);
    a0 = vm_sub<F, T>(q.a, A);
    a0 = vm_sub<F, T>(a0, HH); V::st(Z, a0);       // Z3 = (Z1+H)^2 - Z1Z1 - HH
    a0 = V::ld(HH);
    a0 = vm_small<F, 4>(a0);                       // I = 4 HH
    V::st(HH, q.b);                                // r^2 (HH is dead)
    fe2 jv = vm_mul2<F, T>(a0, H, a0, X);          // J = H I, V = X1 I
    V::st(H, jv.a); V::st(A, jv.b);
    a0 = V::ld(HH);
    a0 = vm_sub<F, T>(a0, H);
    a0 = vm_sub<F, T>(a0, A);
    a0 = vm_sub<F, T>(a0, A); V::st(X, a0);        // X3 = r^2 - J - 2V
    a0 = vm_rsub<F, T>(a0, A);                     // V - X3
    fe2 w = vm_mul2<F, T>(a0, R, jv.a, Y);         // r (V - X3), Y1 J
    a1 = vm_small<F, 2>(w.b);
    fe y;
    F::sub(y.v, w.a.v, a1.v); V::st(Y, y);         // Y3 = r (V - X3) - 2 Y1 J
    return;
  }
#endif
  fe a = V::ld(H);
  a = vm_sqr<F>(a); V::st(HH, a);                // HH
#ifdef FEC_K1_MADD_Z_MUL
  a = vm_small<F, 4>(a); V::st(HH, a);           // I = 4 HH       (A/B variant: Z3 = 2 Z1 H by a multiplication)
  a = V::ld(Z);
  a = vm_mul<F, T>(a, H);
  a = vm_small<F, 2>(a); V::st(Z, a);
  a = V::ld(HH);
#else
  a = V::ld(Z);
  a = vm_add<F, T>(a, H);
  a = vm_sqr<F>(a);
  a = vm_sub<F, T>(a, A);
  a = vm_sub<F, T>(a, HH); V::st(Z, a);          // Z3 = (Z1+H)^2 - Z1Z1 - HH
  a = V::ld(HH);
  a = vm_small<F, 4>(a); V::st(HH, a);           // I = 4 HH
#endif
  a = vm_mul<F, T>(a, H); V::st(H, a);           // J = H I
  a = V::ld(X);
  a = vm_mul<F, T>(a, HH); V::st(A, a);          // V = X1 I
  a = V::ld(R);
  a = vm_sqr<F>(a);
  a = vm_sub<F, T>(a, H);
  a = vm_sub<F, T>(a, A);
  a = vm_sub<F, T>(a, A); V::st(X, a);           // X3 = r^2 - J - 2V
  a = vm_rsub<F, T>(a, A);                       // V - X3
  a = vm_mul<F, T>(a, R); V::st(R, a);           // r (V - X3)
  a = V::ld(Y);
  a = vm_mul<F, T>(a, H);
  a = vm_small<F, 2>(a);                         // 2 Y1 J
  a = vm_rsub<F, T>(a, R); V::st(Y, a);          // Y3
}

}  // namespace fec
