// k_nonce.cu -- signing-side nonce commitment: the fixed-base step R = k*G / [r]B of a signature and the
// scalar that goes with it, batched.  The multiplication itself is the fixed-base kernel (k_wmul_fixed*.cuh, k_ed25519.cu);
// the kernels here are the short tails in front of / behind it:
//
//   ECDSA    (forge-ec-signature/src/ecdsa.rs:108-131)   R = k G;  r = x(R) read as a scalar -- the reference takes
//            Scalar::from_bytes(x), which FAILS for x >= n (Error::InvalidScalar, :116-121), and rejects r = 0 (:125-129)
//   BIP-340  (schnorr.rs:372-385)                         R = k G;  x-only R;  k <- n - k when y(R) is odd
//   EdDSA    (eddsa.rs:120-131, 322-330; RFC 8032 5.1.6)  r = the 64-byte little-endian hash value mod L;  R = [r]B
//            (the reference reduces only the first 32 bytes of the hash -- one of its deviations, SURVEY.md section 0;
//            this is the RFC 8032 value)
#include "kernels.cuh"

namespace fec {

constexpr int NONCE_THREADS = 128;

__device__ __forceinline__ void store_rec32_be(uint8_t* p, const u32 r[8]) {  // p 16-byte aligned
  u32 t[8];
#pragma unroll
  for (int i = 0; i < 8; i++) t[i] = __byte_perm(r[7 - i], 0, 0x0123);
  store_rec32_le(p, t);
}

// xy: n x 64 B affine R = k G (all-zero where the lane is dead), st_in: the lane status of that multiplication
template <class C>
__global__ void __launch_bounds__(NONCE_THREADS)
k_nonce_tail(int mode, const uint8_t* __restrict__ xy, const uint8_t* __restrict__ k_be, const uint8_t* st_in, size_t n,
             uint8_t* __restrict__ out_R, uint8_t* __restrict__ out_r, uint8_t* status /* may alias st_in */) {
  __shared__ __align__(16) uint8_t io_all[(NONCE_THREADS / 32) * 33 * 32 + 16];
  const unsigned lane_in_warp = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint8_t* io = io_all + warp * (33 * 32);
  const size_t first = ((size_t)blockIdx.x * (NONCE_THREADS / 32) + warp) * 32;
  if (first >= n) return;  // warp-uniform
  const int cnt_lanes = (int)((n - first) < 32 ? (n - first) : 32);
  const bool live = (int)lane_in_warp < cnt_lanes;
  const size_t lane = live ? first + lane_in_warp : n - 1;
  u32 x[8], y[8], k[8], nn[8];
  load_rec32_be(x, xy + lane * 64);
  load_rec32_be(y, xy + lane * 64 + 32);
  load_rec32_be(k, k_be + lane * 32);
  C::N(nn);
  u32 st = st_in[lane];
  u32 r[8];
  if (mode == FEC_NONCE_ECDSA) {
    // r = x(R) as a scalar: out of range where the reference's Scalar::from_bytes fails, identity where r = 0
    const u32 ge = geq8(x, nn), zero = is_zero8(x);
    if (st == 0u) st = (ge & LANE_SCALAR_RANGE) | (~ge & zero & LANE_INFINITY);
    copy8(r, x);
  } else {
    // BIP-340: the signer continues with -k when y(R) is odd
    u32 nk[8];
    sub8(nk, nn, k);
    select8(r, 0u - (y[0] & 1u), nk, k);
  }
  const u32 dead = st != 0u ? 0xFFFFFFFFu : 0u;
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] &= ~dead;
  if (live) {
    store_rec32_be(out_r + lane * 32, r);
    if (status != nullptr) status[lane] = (uint8_t)st;
  }
  if (mode == FEC_NONCE_ECDSA) {  // SEC1 compressed R through the warp's staging area (33-byte records)
    if (live) sec1_encode(io + lane_in_warp * 33, 0, x, y, dead);
    __syncwarp();
    warp_copy(out_R + first * 33, io, (size_t)cnt_lanes * 33);
  } else if (live) {
    u32 xx[8];
#pragma unroll
    for (int i = 0; i < 8; i++) xx[i] = x[i] & ~dead;
    store_rec32_be(out_R + lane * 32, xx);
  }
}

// 64-byte little-endian value mod L = 2^252 + 27742317777372353535851937790883648493, byte-wise Horner from the top:
// with r < L, v = 256 r + b < 2^261, q = v >> 252 < 2^9 and v - q L = (v mod 2^252) - q c lies in (-2^134, 2^252):
// one conditional addition of L brings it back below L.  Branch-free.
__global__ void __launch_bounds__(NONCE_THREADS)
k_sc_reduce64(const uint8_t* __restrict__ in64, size_t n, uint8_t* __restrict__ out_le) {
  const size_t lane = (size_t)blockIdx.x * NONCE_THREADS + threadIdx.x;
  if (lane >= n) return;
  const u32 C0 = 0x5CF5D3EDu, C1 = 0x5812631Au, C2 = 0xA2F79CD6u, C3 = 0x14DEF9DEu;  // c = L - 2^252, little-endian words
  u32 w[16];
  {
    const uint4* p = reinterpret_cast<const uint4*>(in64 + lane * 64);
#pragma unroll
    for (int i = 0; i < 4; i++) { const uint4 v = p[i]; w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
  }
  // the top 31 bytes (248 bits) are below 2^252 already
  u32 r[8];
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = (i < 7 ? (w[8 + i] >> 8) | (w[9 + i] << 24) : (w[15] >> 8));
#pragma unroll 1
  for (int byte = 32; byte >= 0; byte--) {
    const u32 b = (w[byte >> 2] >> (8 * (byte & 3))) & 0xFFu;
    // v = 256 r + b: 9 words; q = v >> 252; t = v mod 2^252
    const u32 q = r[7] >> 20;  // bits 252.. of v = bits 244.. of r (r < 2^253: at most 9 bits)
    u32 t[8];
#pragma unroll
    for (int i = 7; i >= 1; i--) t[i] = (r[i] << 8) | (r[i - 1] >> 24);
    t[0] = (r[0] << 8) | b;
    t[7] &= 0x0FFFFFFFu;
    // s = q c (q < 2^9, c < 2^125): five words
    const u64 m0 = (u64)q * C0, m1 = (u64)q * C1 + (m0 >> 32), m2 = (u64)q * C2 + (m1 >> 32), m3 = (u64)q * C3 + (m2 >> 32);
    const u32 s0 = (u32)m0, s1 = (u32)m1, s2 = (u32)m2, s3 = (u32)m3, s4 = (u32)(m3 >> 32);
    u32 d[8];
    d[0] = sub_cc(t[0], s0); d[1] = subc_cc(t[1], s1); d[2] = subc_cc(t[2], s2); d[3] = subc_cc(t[3], s3); d[4] = subc_cc(t[4], s4);
    d[5] = subc_cc(t[5], 0u); d[6] = subc_cc(t[6], 0u); d[7] = subc_cc(t[7], 0u);
    const u32 neg = subc(0u, 0u);  // all ones if the difference is negative: add L
    r[0] = add_cc(d[0], neg & C0); r[1] = addc_cc(d[1], neg & C1); r[2] = addc_cc(d[2], neg & C2); r[3] = addc_cc(d[3], neg & C3);
    r[4] = addc_cc(d[4], 0u); r[5] = addc_cc(d[5], 0u); r[6] = addc_cc(d[6], 0u); r[7] = addc(d[7], neg & 0x10000000u);
  }
  // r < L here except for the representative in [2^252, L) ... which is below L as well; nothing else to do
  store_rec32_le(out_le + lane * 32, r);
}

// Standalone point decoding / validation (CompressedPoint::to_affine, forge-ec-encoding/src/point.rs:70-144, and
// Curve::validate_point, forge-ec-core/src/lib.rs:905-923 -- on these cofactor-1 curves "on the curve" is the whole check):
// n x {33 B SEC1 compressed | 64 B X || Y} -> n records of out_fmt + status (0 valid, 2 rejected, 3 the identity record).
template <class C>
__global__ void __launch_bounds__(NONCE_THREADS)
k_point_decode(const uint8_t* __restrict__ points, int points_fmt, size_t n, int out_fmt, uint8_t* __restrict__ out,
               uint8_t* __restrict__ status) {
  typedef typename C::F F;
  __shared__ __align__(16) uint8_t in_all[(NONCE_THREADS / 32) * 33 * 32 + 16];
  __shared__ __align__(16) uint8_t out_all[(NONCE_THREADS / 32) * 65 * 32 + 16];
  const unsigned lane_in_warp = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint8_t* in = in_all + warp * (33 * 32);
  uint8_t* io = out_all + warp * (65 * 32);
  const size_t first = ((size_t)blockIdx.x * (NONCE_THREADS / 32) + warp) * 32;
  if (first >= n) return;  // warp-uniform
  const int cnt_lanes = (int)((n - first) < 32 ? (n - first) : 32);
  const bool live = (int)lane_in_warp < cnt_lanes;
  const size_t lane = first + (live ? lane_in_warp : 0u);
  u32 px[8], py[8], ok, idn = 0u;
  if (points_fmt == FEC_PT_SEC1_COMPRESSED) {
    warp_copy(in, points + first * 33, (size_t)cnt_lanes * 33);
    __syncwarp();
    ok = sec1_decompress_id<C>(px, py, in + (live ? lane_in_warp : 0u) * 33, idn);
  } else {
    load_rec32_be(px, points + lane * 64);
    load_rec32_be(py, points + lane * 64 + 32);
    ok = affine_from_plain_id<C>(px, py, idn);
  }
  const u32 st = lane_status(0u, ok, idn);
  u32 x[8], y[8];
  F::to_plain(x, px);
  F::to_plain(y, py);
  const int osz = out_size(out_fmt);
  if (live) sec1_encode(io + lane_in_warp * osz, out_fmt, x, y, st != 0u ? 0xFFFFFFFFu : 0u);
  __syncwarp();
  warp_copy(out + first * osz, io, (size_t)cnt_lanes * osz);
  if (live && status != nullptr) status[lane] = (uint8_t)st;
}

cudaError_t point_decode_launch(int curve, cudaStream_t st, const uint8_t* points, int points_fmt, size_t n, int out_fmt, uint8_t* out,
                                uint8_t* status) {
  const unsigned g = (unsigned)((n + NONCE_THREADS - 1) / NONCE_THREADS);
  if (curve == FEC_SECP256K1) k_point_decode<CurveK1><<<g, NONCE_THREADS, 0, st>>>(points, points_fmt, n, out_fmt, out, status);
  else k_point_decode<CurveP256><<<g, NONCE_THREADS, 0, st>>>(points, points_fmt, n, out_fmt, out, status);
  return cudaGetLastError();
}
cudaError_t nonce_tail_launch(int curve, int mode, cudaStream_t st, const uint8_t* xy, const uint8_t* k_be, const uint8_t* st_in, size_t n,
                              uint8_t* out_R, uint8_t* out_r, uint8_t* status) {
  const unsigned g = (unsigned)((n + NONCE_THREADS - 1) / NONCE_THREADS);
  if (curve == FEC_SECP256K1) k_nonce_tail<CurveK1><<<g, NONCE_THREADS, 0, st>>>(mode, xy, k_be, st_in, n, out_R, out_r, status);
  else k_nonce_tail<CurveP256><<<g, NONCE_THREADS, 0, st>>>(mode, xy, k_be, st_in, n, out_R, out_r, status);
  return cudaGetLastError();
}
cudaError_t sc_reduce64_launch(cudaStream_t st, const uint8_t* in64, size_t n, uint8_t* out_le) {
  k_sc_reduce64<<<(unsigned)((n + NONCE_THREADS - 1) / NONCE_THREADS), NONCE_THREADS, 0, st>>>(in64, n, out_le);
  return cudaGetLastError();
}

}  // namespace fec
