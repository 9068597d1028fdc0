// u256.cuh -- 256-bit integer building blocks on 8 x u32 limbs held in registers (sm_100a).
//
// Every 32x32->64 multiply is written as a `mad.lo.cc.u32` / `madc.hi.cc.u32` pair on the
// same operands whose results land in an even-aligned register pair; ptxas fuses each pair
// into ONE `IMAD.WIDE.U32(.X)` (checked with cuobjdump -sass, see profiles/).  A full 8x8
// schoolbook product is therefore 64 IMAD.WIDE plus ~22 carry fix-ups on the ALU pipe, a
// square is 36.  Column parities are kept in two accumulators (`ev`: limb positions i+j
// even, `od`: odd) so that every product is pair-aligned, and are merged once at the end.
//
// Replaces (as a building block) the 4 x u64 -> u128 schoolbook loops of the reference:
// forge-ec-curves/src/secp256k1.rs:449-466, p256.rs:501-533, curve25519.rs:230-313,
// ed25519.rs:522-545.
#pragma once
#include <cstdint>

namespace fec {

typedef uint32_t u32;
typedef uint64_t u64;

#if defined(__CUDACC__)
#define FEC_DEV __device__ __forceinline__
#else
#define FEC_DEV inline
#endif
// a warp-uniform branch that (almost) never runs: its body belongs out of the instruction stream of the hot path
#ifdef FEC_NO_BRANCH_HINTS
#define FEC_UNLIKELY(x) (x)
#else
#define FEC_UNLIKELY(x) __builtin_expect(!!(x), 0)
#endif

#if defined(__CUDA_ARCH__) || defined(__CUDACC__)
// ---- carry-flag primitives (one PTX instruction each; `volatile` keeps their order, the
// compiler itself never emits .cc instructions so the flag cannot be clobbered in between)
FEC_DEV u32 add_cc(u32 a, u32 b) { u32 r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FEC_DEV u32 addc_cc(u32 a, u32 b) { u32 r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FEC_DEV u32 addc(u32 a, u32 b) { u32 r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FEC_DEV u32 sub_cc(u32 a, u32 b) { u32 r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FEC_DEV u32 subc_cc(u32 a, u32 b) { u32 r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FEC_DEV u32 subc(u32 a, u32 b) { u32 r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FEC_DEV u32 mul_lo(u32 a, u32 b) { u32 r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FEC_DEV u32 mul_hi(u32 a, u32 b) { u32 r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FEC_DEV u32 mad_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
FEC_DEV u32 madc_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
FEC_DEV u32 madc_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
FEC_DEV u32 madc_hi(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
FEC_DEV u32 funnel_l(u32 lo, u32 hi, u32 s) { return __funnelshift_l(lo, hi, s); }
#else
// ---- TEST-ONLY host emulation of the PTX carry flag (tests/hostsim compiles these headers with
// g++ so the limb arithmetic can be checked on a machine without a GPU).  It is never part of
// the shipped library: forge_ec_cuda.cu is compiled by nvcc only and has no CPU code path.
static thread_local u32 fec_cc_flag = 0;  // (FEC_DEV is plain `inline` on the host)
FEC_DEV u32 fec_setcc(u64 t) { fec_cc_flag = (u32)(t >> 32) & 1u; return (u32)t; }
FEC_DEV u32 add_cc(u32 a, u32 b) { return fec_setcc((u64)a + b); }
FEC_DEV u32 addc_cc(u32 a, u32 b) { return fec_setcc((u64)a + b + fec_cc_flag); }
FEC_DEV u32 addc(u32 a, u32 b) { return (u32)((u64)a + b + fec_cc_flag); }
FEC_DEV u32 sub_cc(u32 a, u32 b) { return fec_setcc((u64)a - b); }  // borrow = bit 32 of the wrapped difference
FEC_DEV u32 subc_cc(u32 a, u32 b) { return fec_setcc((u64)a - b - fec_cc_flag); }
FEC_DEV u32 subc(u32 a, u32 b) { return (u32)((u64)a - b - fec_cc_flag); }
FEC_DEV u32 mul_lo(u32 a, u32 b) { return (u32)((u64)a * b); }
FEC_DEV u32 mul_hi(u32 a, u32 b) { return (u32)(((u64)a * b) >> 32); }
FEC_DEV u32 mad_lo_cc(u32 a, u32 b, u32 c) { return fec_setcc((u64)(u32)((u64)a * b) + c); }
FEC_DEV u32 madc_lo_cc(u32 a, u32 b, u32 c) { return fec_setcc((u64)(u32)((u64)a * b) + c + fec_cc_flag); }
FEC_DEV u32 madc_hi_cc(u32 a, u32 b, u32 c) { return fec_setcc((((u64)a * b) >> 32) + c + fec_cc_flag); }
FEC_DEV u32 madc_hi(u32 a, u32 b, u32 c) { return (u32)((((u64)a * b) >> 32) + c + fec_cc_flag); }
FEC_DEV u32 funnel_l(u32 lo, u32 hi, u32 s) { return s ? ((hi << s) | (lo >> (32 - s))) : hi; }
#endif

// acc[0..2*NP) += {a0,a1,..}[k] * b laid out as NP (lo,hi) pairs, one carry chain.
// `fresh` = how many of the top words of the span have never been written (their old value is
// taken as 0); when fresh==0 the chain's carry-out is returned (else 0 -- it cannot carry).
// An always-zero value that ptxas cannot fold (two byte permutations that select the same byte), or plain 0.  The zero
// half of a (carry word, 0) addend pair is then `z0 & carry word`, a LOP3 on the ALU pipe written straight into the
// pair's register, where ptxas would otherwise place an IMAD.MOV -- on the multiply pipe, the one that bounds the
// kernels (9 of the 136 instructions of a secp256k1 multiplication).
FEC_DEV u32 opaque_zero(u32 x) {
#if defined(__CUDA_ARCH__) && defined(FEC_ZERO_ON_ALU)
  u32 p, q;
  asm("prmt.b32 %0, %1, %1, 0x0000;" : "=r"(p) : "r"(x));
  asm("prmt.b32 %0, %1, %1, 0x4444;" : "=r"(q) : "r"(x));
  return p ^ q;
#else
  (void)x;
  return 0u;
#endif
}
FEC_DEV u32 opaque_zero_always(u32 x) {  // the same, not switched by FEC_ZERO_ON_ALU (ordering tokens of FeP256::mul2_inl)
#if defined(__CUDA_ARCH__)
  u32 p, q;
  asm("prmt.b32 %0, %1, %1, 0x0000;" : "=r"(p) : "r"(x));
  asm("prmt.b32 %0, %1, %1, 0x4444;" : "=r"(q) : "r"(x));
  return p ^ q;
#else
  (void)x;
  return 0u;
#endif
}
template <int NP, int FRESH>
FEC_DEV u32 mad_row(u32* acc, const u32* a /*stride 2*/, u32 b, u32 z0 = 0u) {
  static_assert(FRESH >= 0 && FRESH <= 2, "fresh words");
#pragma unroll
  for (int k = 0; k < NP; k++) {
    const u32 ak = a[2 * k];
    const bool lo_fresh = (2 * k >= 2 * NP - FRESH);
    const bool hi_fresh = (2 * k + 1 >= 2 * NP - FRESH);
    const u32 lo_in = lo_fresh ? 0u : acc[2 * k];
    // FRESH == 1: the pair is (carry word of an earlier row, never written).  See opaque_zero.
    const u32 hi_in = hi_fresh ? ((FRESH == 1) ? (z0 & acc[2 * k]) : 0u) : acc[2 * k + 1];
    acc[2 * k] = (k == 0) ? mad_lo_cc(ak, b, lo_in) : madc_lo_cc(ak, b, lo_in);
    if (k == NP - 1 && FRESH > 0)
      acc[2 * k + 1] = madc_hi(ak, b, hi_in);
    else
      acc[2 * k + 1] = madc_hi_cc(ak, b, hi_in);
  }
  if (FRESH == 0) return addc(0u, 0u);
  return 0u;
}

// The same row on a span whose top pair has been written before (FRESH == 0), with the chain's carry-out ABSORBED into `sink`
// instead of being materialised as a word of its own: `sink` is the word of the OTHER column parity at the carry's position,
// which this row has just written as the high word of a fresh top pair -- at most 2^32 - 2 ((2^32-1)^2 + 1 = 2^64 - 2^33 + 2),
// so the +1 cannot overflow.  Saves, per carry, the SEL that makes the carry word and the zeroed partner register the next
// row's top product would need (14 of the 65 non-multiply instructions of a multiplication were such pairs).
template <int NP>
FEC_DEV void mad_row_absorb(u32* acc, const u32* a /*stride 2*/, u32 b, u32& sink) {
#pragma unroll
  for (int k = 0; k < NP; k++) {
    const u32 ak = a[2 * k];
    acc[2 * k] = (k == 0) ? mad_lo_cc(ak, b, acc[2 * k]) : madc_lo_cc(ak, b, acc[2 * k]);
    acc[2 * k + 1] = madc_hi_cc(ak, b, acc[2 * k + 1]);
  }
  sink = addc(sink, 0u);
}

// r[0..16) = a[0..8) * b[0..8), schoolbook
// Tokens for callers that want to order later work behind the product (FeP256::mul_inl): tok_e + tok_o (the two top partial
// words, final before the merge) never overflows 32 bits; with TOKC, tok_c is the carry out of the merge chain -- zero
// for every input, because the product is below 2^512.
template <bool TOKC>
FEC_DEV void mul_wide_school_tok(u32 r[16], const u32 a[8], const u32 b[8], u32& tok_e, u32& tok_o, u32& tok_c) {
  const u32 z0 = opaque_zero(a[0]);
  u32 ev[16];  // ev[k]: weight 2^(32k)
  u32 od[15];  // od[k]: weight 2^(32(k+1))
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    ev[j] = mul_lo(a[j], b[0]);
    ev[j + 1] = mul_hi(a[j], b[0]);
    od[j] = mul_lo(a[j + 1], b[0]);
    od[j + 1] = mul_hi(a[j + 1], b[0]);
  }
#ifdef FEC_CARRY_WORDS  // the form before r2: every full chain's carry becomes a word of its own (A/B)
  // odd rows: a_odd*b[i] -> even positions, a_even*b[i] -> odd positions (full -> carry word)
  // even rows: a_even*b[i] -> even positions (full -> carry word), a_odd*b[i] -> odd positions
  mad_row<4, 2>(ev + 2, a + 1, b[1]);  od[8]  = mad_row<4, 0>(od + 0, a, b[1]);
  ev[10] = mad_row<4, 0>(ev + 2, a, b[2]);  mad_row<4, 1>(od + 2, a + 1, b[2], z0);
  mad_row<4, 1>(ev + 4, a + 1, b[3], z0);  od[10] = mad_row<4, 0>(od + 2, a, b[3]);
  ev[12] = mad_row<4, 0>(ev + 4, a, b[4]);  mad_row<4, 1>(od + 4, a + 1, b[4], z0);
  mad_row<4, 1>(ev + 6, a + 1, b[5], z0);  od[12] = mad_row<4, 0>(od + 4, a, b[5]);
  ev[14] = mad_row<4, 0>(ev + 6, a, b[6]);  mad_row<4, 1>(od + 6, a + 1, b[6], z0);
  mad_row<4, 1>(ev + 8, a + 1, b[7], z0);  od[14] = mad_row<4, 0>(od + 6, a, b[7]);
#else
  // Row i = two carry chains: the products that land on the parity whose top pair is NEW this row (a_odd * b[i]: fresh pair,
  // no carry out), then the others over a span that is full (a_even * b[i]), whose carry-out sits exactly at the position of the
  // fresh pair's high word in the other parity and is absorbed there (mad_row_absorb).  No carry words, no zeroed partners.
  (void)z0;
  mad_row<4, 2>(ev + 2, a + 1, b[1]);  mad_row_absorb<4>(od + 0, a, b[1], ev[9]);
  mad_row<4, 2>(od + 2, a + 1, b[2]);  mad_row_absorb<4>(ev + 2, a, b[2], od[9]);
  mad_row<4, 2>(ev + 4, a + 1, b[3]);  mad_row_absorb<4>(od + 2, a, b[3], ev[11]);
  mad_row<4, 2>(od + 4, a + 1, b[4]);  mad_row_absorb<4>(ev + 4, a, b[4], od[11]);
  mad_row<4, 2>(ev + 6, a + 1, b[5]);  mad_row_absorb<4>(od + 4, a, b[5], ev[13]);
  mad_row<4, 2>(od + 6, a + 1, b[6]);  mad_row_absorb<4>(ev + 6, a, b[6], od[13]);
  mad_row<4, 2>(ev + 8, a + 1, b[7]);  mad_row_absorb<4>(od + 6, a, b[7], ev[15]);
  od[14] = 0u;  // (no word of the odd parity exists at position 15 any more)
#endif
  // merge: r = ev + (od << 32).  ev[15] was written by row 7 (fresh), od[14] by row 7's carry (zero in the absorbing form).
  r[0] = ev[0];
  r[1] = add_cc(ev[1], od[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) r[k] = addc_cc(ev[k], od[k - 1]);
  tok_c = 0u;
  if (TOKC) {
    r[15] = addc_cc(ev[15], od[14]);
    tok_c = addc(0u, 0u);
  } else {
    r[15] = addc(ev[15], od[14]);
  }
  tok_e = ev[15]; tok_o = od[14];
}
// The product left in its two column-parity accumulators, NOT merged: a * b = sum ev[k] 2^(32k) + sum od[k] 2^(32(k+1)), ev[0..16),
// od[0..14), every (x[2j], x[2j+1]) an aligned register pair.  For reductions that fold the high half back as the addends of
// IMAD.WIDE chains on these very pairs (Fe25519::mul_inl): merging first costs 8 additions more and moves to re-align pairs.
FEC_DEV void mul_wide_evod(u32 ev[16], u32 od[14], const u32 a[8], const u32 b[8]) {
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    ev[j] = mul_lo(a[j], b[0]);
    ev[j + 1] = mul_hi(a[j], b[0]);
    od[j] = mul_lo(a[j + 1], b[0]);
    od[j + 1] = mul_hi(a[j + 1], b[0]);
  }
  mad_row<4, 2>(ev + 2, a + 1, b[1]);  mad_row_absorb<4>(od + 0, a, b[1], ev[9]);
  mad_row<4, 2>(od + 2, a + 1, b[2]);  mad_row_absorb<4>(ev + 2, a, b[2], od[9]);
  mad_row<4, 2>(ev + 4, a + 1, b[3]);  mad_row_absorb<4>(od + 2, a, b[3], ev[11]);
  mad_row<4, 2>(od + 4, a + 1, b[4]);  mad_row_absorb<4>(ev + 4, a, b[4], od[11]);
  mad_row<4, 2>(ev + 6, a + 1, b[5]);  mad_row_absorb<4>(od + 4, a, b[5], ev[13]);
  mad_row<4, 2>(od + 6, a + 1, b[6]);  mad_row_absorb<4>(ev + 6, a, b[6], od[13]);
  mad_row<4, 2>(ev + 8, a + 1, b[7]);  mad_row_absorb<4>(od + 6, a, b[7], ev[15]);
}
FEC_DEV void mul_wide_school(u32 r[16], const u32 a[8], const u32 b[8]) {
  u32 te, to, tc;
  mul_wide_school_tok<false>(r, a, b, te, to, tc);
}

// The same product in STAGES, for routines that interleave two independent field operations at carry-chain granularity
// (FeP256::mul2_inl): stage 0 = row b[0] (plain products), stages 1..7 = rows b[1..7] (two carry chains each),
// prod_merge = r = ev + (od << 32) with the carry out of the chain in tok_c (zero for every input).
struct ProdSt { u32 ev[16]; u32 od[15]; };
template <int R>
FEC_DEV void prod_row(ProdSt& s, const u32 a[8], const u32 b[8]) {
  u32* ev = s.ev;
  u32* od = s.od;
  if (R == 0) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      ev[j] = mul_lo(a[j], b[0]);
      ev[j + 1] = mul_hi(a[j], b[0]);
      od[j] = mul_lo(a[j + 1], b[0]);
      od[j + 1] = mul_hi(a[j + 1], b[0]);
    }
  } else if (R == 1) {
    mad_row<4, 2>(ev + 2, a + 1, b[1]);  od[8] = mad_row<4, 0>(od + 0, a, b[1]);
  } else if (R == 2) {
    ev[10] = mad_row<4, 0>(ev + 2, a, b[2]);  mad_row<4, 1>(od + 2, a + 1, b[2]);
  } else if (R == 3) {
    mad_row<4, 1>(ev + 4, a + 1, b[3]);  od[10] = mad_row<4, 0>(od + 2, a, b[3]);
  } else if (R == 4) {
    ev[12] = mad_row<4, 0>(ev + 4, a, b[4]);  mad_row<4, 1>(od + 4, a + 1, b[4]);
  } else if (R == 5) {
    mad_row<4, 1>(ev + 6, a + 1, b[5]);  od[12] = mad_row<4, 0>(od + 4, a, b[5]);
  } else if (R == 6) {
    ev[14] = mad_row<4, 0>(ev + 6, a, b[6]);  mad_row<4, 1>(od + 6, a + 1, b[6]);
  } else {
    mad_row<4, 1>(ev + 8, a + 1, b[7]);  od[14] = mad_row<4, 0>(od + 6, a, b[7]);
  }
}
FEC_DEV void prod_merge(u32 r[16], u32& tok_c, const ProdSt& s) {
  r[0] = s.ev[0];
  r[1] = add_cc(s.ev[1], s.od[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) r[k] = addc_cc(s.ev[k], s.od[k - 1]);
  r[15] = addc_cc(s.ev[15], s.od[14]);
  tok_c = addc(0u, 0u);
}

// r[0..8) = a[0..4) * b[0..4): the same even/odd accumulation on half-size operands
FEC_DEV void mul_wide4(u32 r[8], const u32 a[4], const u32 b[4]) {
  const u32 z0 = 0u;  // (Karatsuba halves: A/B variant only)
  u32 ev[8], od[7];
  ev[0] = mul_lo(a[0], b[0]); ev[1] = mul_hi(a[0], b[0]);
  ev[2] = mul_lo(a[2], b[0]); ev[3] = mul_hi(a[2], b[0]);
  od[0] = mul_lo(a[1], b[0]); od[1] = mul_hi(a[1], b[0]);
  od[2] = mul_lo(a[3], b[0]); od[3] = mul_hi(a[3], b[0]);
  mad_row<2, 2>(ev + 2, a + 1, b[1]);  od[4] = mad_row<2, 0>(od + 0, a, b[1]);
  ev[6] = mad_row<2, 0>(ev + 2, a, b[2]);  mad_row<2, 1>(od + 2, a + 1, b[2], z0);
  mad_row<2, 1>(ev + 4, a + 1, b[3], z0);  od[6] = mad_row<2, 0>(od + 2, a, b[3]);
  r[0] = ev[0];
  r[1] = add_cc(ev[1], od[0]);
#pragma unroll
  for (int k = 2; k < 7; k++) r[k] = addc_cc(ev[k], od[k - 1]);
  r[7] = addc(ev[7], od[6]);
}
// |x - y| for 4-limb operands; returns the sign mask (all ones if x < y)
FEC_DEV u32 absdiff4(u32 d[4], const u32 x[4], const u32 y[4]) {
  u32 t[4];
  t[0] = sub_cc(x[0], y[0]);
  t[1] = subc_cc(x[1], y[1]);
  t[2] = subc_cc(x[2], y[2]);
  t[3] = subc_cc(x[3], y[3]);
  const u32 m = subc(0u, 0u);  // 0 or 0xFFFFFFFF
  d[0] = add_cc(t[0] ^ m, m & 1u);
  d[1] = addc_cc(t[1] ^ m, 0u);
  d[2] = addc_cc(t[2] ^ m, 0u);
  d[3] = addc(t[3] ^ m, 0u);
  return m;
}
// r[0..16) = a * b by one level of (subtractive) Karatsuba: 48 multiplications instead of 64, ~70 more additions.
// a1 b0 + a0 b1 = a0 b0 + a1 b1 - (a0 - a1)(b0 - b1).  Branch-free.
FEC_DEV void mul_wide_karatsuba(u32 r[16], const u32 a[8], const u32 b[8]) {
  u32 z0[8], z2[8], m[8], da[4], db[4], t[9], z1[9];
  mul_wide4(z0, a, b);
  mul_wide4(z2, a + 4, b + 4);
  const u32 s = absdiff4(da, a, a + 4) ^ absdiff4(db, b, b + 4);  // all ones: the cross product enters with a plus sign
  mul_wide4(m, da, db);
  t[0] = add_cc(z0[0], z2[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) t[i] = addc_cc(z0[i], z2[i]);
  t[8] = addc(0u, 0u);
  const u32 ns = ~s;                       // subtracting: add ~m + 1 with an all-ones extension word
  add_cc(ns, ns);                          // carry flag = (ns != 0): the "+ 1" of the two's complement
#pragma unroll
  for (int i = 0; i < 8; i++) z1[i] = addc_cc(t[i], m[i] ^ ns);
  z1[8] = addc(t[8], ns);
#pragma unroll
  for (int i = 0; i < 4; i++) r[i] = z0[i];
  r[4] = add_cc(z0[4], z1[0]);
  r[5] = addc_cc(z0[5], z1[1]);
  r[6] = addc_cc(z0[6], z1[2]);
  r[7] = addc_cc(z0[7], z1[3]);
  r[8] = addc_cc(z2[0], z1[4]);
  r[9] = addc_cc(z2[1], z1[5]);
  r[10] = addc_cc(z2[2], z1[6]);
  r[11] = addc_cc(z2[3], z1[7]);
  r[12] = addc_cc(z2[4], z1[8]);
  r[13] = addc_cc(z2[5], 0u);
  r[14] = addc_cc(z2[6], 0u);
  r[15] = addc(z2[7], 0u);
}

// r[0..16) = a[0..8)^2 : 28 off-diagonal products (doubled by a 1-bit funnel shift) + 8 diagonals.
// (TOKC: tok_c receives the carry out of the final chain -- zero for every input, the square is below 2^512)
template <bool TOKC>
FEC_DEV void sqr_wide_tok(u32 r[16], const u32 a[8], u32& tok_c) {
  const u32 z0 = opaque_zero(a[0]);
  // off-diagonal sum S = sum_{i<j} a[i]a[j] 2^(32(i+j)); positions 1..14 (+carry word 15 stays 0)
  u32 ev[16];  // even positions 2..14 used: ev[k] weight 2^(32k)
  u32 od[15];  // odd positions: od[k] weight 2^(32(k+1)), positions 1..13 -> od[0..13]
  // row 0: a0*a1 (1,2) a0*a3 (3,4) a0*a5 (5,6) a0*a7 (7,8) -> od[0..7]; a0*a2 (2,3) a0*a4 (4,5) a0*a6 (6,7) -> ev[2..7]
  od[0] = mul_lo(a[0], a[1]); od[1] = mul_hi(a[0], a[1]);
  od[2] = mul_lo(a[0], a[3]); od[3] = mul_hi(a[0], a[3]);
  od[4] = mul_lo(a[0], a[5]); od[5] = mul_hi(a[0], a[5]);
  od[6] = mul_lo(a[0], a[7]); od[7] = mul_hi(a[0], a[7]);
  ev[2] = mul_lo(a[0], a[2]); ev[3] = mul_hi(a[0], a[2]);
  ev[4] = mul_lo(a[0], a[4]); ev[5] = mul_hi(a[0], a[4]);
  ev[6] = mul_lo(a[0], a[6]); ev[7] = mul_hi(a[0], a[6]);
#ifdef FEC_CARRY_WORDS  // the form before r2 (A/B): carry words
  // row 1: a1*a2 (3,4) a1*a4 (5,6) a1*a6 (7,8) -> od[2..7] full -> carry od[8];
  //        a1*a3 (4,5) a1*a5 (6,7) a1*a7 (8,9) -> ev[4..9], ev[8],ev[9] fresh
  od[8] = mad_row<3, 0>(od + 2, a + 2, a[1]);
  mad_row<3, 2>(ev + 4, a + 3, a[1]);
  // row 2: a2*a3 (5,6) a2*a5 (7,8) a2*a7 (9,10) -> od[4..9], od[9] fresh;
  //        a2*a4 (6,7) a2*a6 (8,9) -> ev[6..9] full -> carry ev[10]
  mad_row<3, 1>(od + 4, a + 3, a[2], z0);
  ev[10] = mad_row<2, 0>(ev + 6, a + 4, a[2]);
  // row 3: a3*a4 (7,8) a3*a6 (9,10) -> od[6..9] full -> carry od[10];
  //        a3*a5 (8,9) a3*a7 (10,11) -> ev[8..11], ev[11] fresh
  od[10] = mad_row<2, 0>(od + 6, a + 4, a[3]);
  mad_row<2, 1>(ev + 8, a + 5, a[3], z0);
  // row 4: a4*a5 (9,10) a4*a7 (11,12) -> od[8..11], od[11] fresh;  a4*a6 (10,11) -> ev[10..11] full -> carry ev[12]
  mad_row<2, 1>(od + 8, a + 5, a[4], z0);
  ev[12] = mad_row<1, 0>(ev + 10, a + 6, a[4]);
  // row 5: a5*a6 (11,12) -> od[10..11] full -> carry od[12];  a5*a7 (12,13) -> ev[12..13], ev[13] fresh
  od[12] = mad_row<1, 0>(od + 10, a + 6, a[5]);
  mad_row<1, 1>(ev + 12, a + 7, a[5], z0);
  // row 6: a6*a7 (13,14) -> od[12..13], od[13] fresh
  mad_row<1, 1>(od + 12, a + 7, a[6], z0);
#else
  // Every row: first the chain whose top pair is new (fresh, no carry out), then the chain over a full span, whose carry-out
  // is absorbed into the new pair's high word of the other parity -- the same position (see mad_row_absorb).
  (void)z0;
  // row 1: a1*a3 (4,5) a1*a5 (6,7) a1*a7 (8,9) -> ev[4..9], pair (8,9) new;  a1*a2 (3,4) a1*a4 (5,6) a1*a6 (7,8) -> od[2..7], carry at 9
  mad_row<3, 2>(ev + 4, a + 3, a[1]);
  mad_row_absorb<3>(od + 2, a + 2, a[1], ev[9]);
  // row 2: a2*a3 (5,6) a2*a5 (7,8) a2*a7 (9,10) -> od[4..9], pair (9,10) new;  a2*a4 (6,7) a2*a6 (8,9) -> ev[6..9], carry at 10
  mad_row<3, 2>(od + 4, a + 3, a[2]);
  mad_row_absorb<2>(ev + 6, a + 4, a[2], od[9]);
  // row 3: a3*a5 (8,9) a3*a7 (10,11) -> ev[8..11], pair (10,11) new;  a3*a4 (7,8) a3*a6 (9,10) -> od[6..9], carry at 11
  mad_row<2, 2>(ev + 8, a + 5, a[3]);
  mad_row_absorb<2>(od + 6, a + 4, a[3], ev[11]);
  // row 4: a4*a5 (9,10) a4*a7 (11,12) -> od[8..11], pair (11,12) new;  a4*a6 (10,11) -> ev[10..11], carry at 12
  mad_row<2, 2>(od + 8, a + 5, a[4]);
  mad_row_absorb<1>(ev + 10, a + 6, a[4], od[11]);
  // row 5: a5*a7 (12,13) -> ev[12..13] new;  a5*a6 (11,12) -> od[10..11], carry at 13
  mad_row<1, 2>(ev + 12, a + 7, a[5]);
  mad_row_absorb<1>(od + 10, a + 6, a[5], ev[13]);
  // row 6: a6*a7 (13,14) -> od[12..13] new
  mad_row<1, 2>(od + 12, a + 7, a[6]);
#endif
  // merge off-diagonals: s[1..14] (s[0]=0, s[15] = 0); s fits in 15 words (sum < 2^480)
  u32 s[16];
  s[1] = od[0];
  s[2] = add_cc(ev[2], od[1]);
#pragma unroll
  for (int k = 3; k < 14; k++) s[k] = addc_cc(ev[k], od[k - 1]);
  s[14] = addc(0u, od[13]);
  // double: t = 2*s (funnel shifts), then add diagonals a[i]^2 at (2i, 2i+1) in one chain
  u32 t[16];
  t[0] = 0;
  t[1] = s[1] << 1;
#pragma unroll
  for (int k = 2; k < 15; k++) t[k] = funnel_l(s[k - 1], s[k], 1);
  t[15] = s[14] >> 31;
  r[0] = mad_lo_cc(a[0], a[0], t[0]);
  r[1] = madc_hi_cc(a[0], a[0], t[1]);
#pragma unroll
  for (int i = 1; i < 8; i++) {
    r[2 * i] = madc_lo_cc(a[i], a[i], t[2 * i]);
    if (i < 7 || TOKC)
      r[2 * i + 1] = madc_hi_cc(a[i], a[i], t[2 * i + 1]);
    else
      r[2 * i + 1] = madc_hi(a[i], a[i], t[2 * i + 1]);
  }
  tok_c = TOKC ? addc(0u, 0u) : 0u;
}
FEC_DEV void sqr_wide(u32 r[16], const u32 a[8]) {
  u32 tc;
  sqr_wide_tok<false>(r, a, tc);
}

FEC_DEV void mul_wide(u32 r[16], const u32 a[8], const u32 b[8]) {
#ifdef FEC_KARATSUBA
  mul_wide_karatsuba(r, a, b);
#else
  mul_wide_school(r, a, b);
#endif
}

// ---- 8-limb helpers -------------------------------------------------------------------
// r = a + b, returns carry
FEC_DEV u32 add8(u32 r[8], const u32 a[8], const u32 b[8]) {
  r[0] = add_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) r[i] = addc_cc(a[i], b[i]);
  return addc(0u, 0u);
}
// r = a - b, returns borrow (0/1)
FEC_DEV u32 sub8(u32 r[8], const u32 a[8], const u32 b[8]) {
  r[0] = sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) r[i] = subc_cc(a[i], b[i]);
  return subc(0u, 0u) & 1u;
}
FEC_DEV void copy8(u32 r[8], const u32 a[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = a[i];
}
// r = m ? a : b   (m is an all-ones / all-zero mask)
FEC_DEV void select8(u32 r[8], u32 m, const u32 a[8], const u32 b[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = (a[i] & m) | (b[i] & ~m);
}
FEC_DEV u32 is_zero8(const u32 a[8]) {  // all-ones mask if a == 0
  u32 o = a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7];
  return (o == 0) ? 0xFFFFFFFFu : 0u;
}
FEC_DEV u32 eq8(const u32 a[8], const u32 b[8]) {
  u32 o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a[i] ^ b[i];
  return (o == 0) ? 0xFFFFFFFFu : 0u;
}
// a >= b (unsigned 256-bit) as mask
FEC_DEV u32 geq8(const u32 a[8], const u32 b[8]) {
  u32 t[8];
  u32 bw = sub8(t, a, b);
  return bw ? 0u : 0xFFFFFFFFu;
}

// 32-byte big-endian <-> limbs (limb 0 = least significant)
FEC_DEV void load_be32(u32 r[8], const uint8_t* p) {  // p need not be aligned
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint8_t* q = p + 28 - 4 * i;
    r[i] = ((u32)q[0] << 24) | ((u32)q[1] << 16) | ((u32)q[2] << 8) | (u32)q[3];
  }
}
FEC_DEV void store_be32(uint8_t* p, const u32 a[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint8_t* q = p + 28 - 4 * i;
    q[0] = (uint8_t)(a[i] >> 24); q[1] = (uint8_t)(a[i] >> 16); q[2] = (uint8_t)(a[i] >> 8); q[3] = (uint8_t)a[i];
  }
}
FEC_DEV void load_le32(u32 r[8], const uint8_t* p) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint8_t* q = p + 4 * i;
    r[i] = ((u32)q[3] << 24) | ((u32)q[2] << 16) | ((u32)q[1] << 8) | (u32)q[0];
  }
}
FEC_DEV void store_le32(uint8_t* p, const u32 a[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint8_t* q = p + 4 * i;
    q[0] = (uint8_t)a[i]; q[1] = (uint8_t)(a[i] >> 8); q[2] = (uint8_t)(a[i] >> 16); q[3] = (uint8_t)(a[i] >> 24);
  }
}

}  // namespace fec
