// k_wmul_var.cuh -- variable-base kernel template; instantiated per curve in k_wmul_var_{k1,p256}.cu
//
// Every WARP is independent (no __syncthreads anywhere): it walks tiles of 32 consecutive lanes, keeps its
// Jacobian results of up to DEFER tiles in thread-local memory together with the running product of their Z
// (Montgomery's trick along the tile axis), and only then performs ONE warp-level inversion (shuffle butterfly +
// one Fermat inversion) for all 32 x DEFER results, normalises, encodes and writes them.  The inversion cost per
// result is 3 M + (10 M + I) / DEFER and no warp ever waits for another one.
//   secp256k1: GLV split + table on the isomorphic curve where it is affine (no inversion for the table)
//   P-256    : 64 windows; the a = -3 doubling does not survive the isomorphism, so its table is normalised with
//              a warp-level inversion per tile.
#pragma once
#include "kernels.cuh"

namespace fec {

constexpr int VAR_DEFER = 8;        // tiles per warp-level inversion
constexpr int VAR_SLOT_WORDS = 36;  // X, Y, Z, prefix (8 words each) + meta + padding (16-byte multiples)

// per-lane table of 8 affine points in thread-local memory (the hardware interleaves local memory across the lanes
// of a warp, so every access is coalesced; it is cached in L1/L2 and private to the launch)
struct LocalTab {
  u32 e[8][16];
  __device__ __forceinline__ void store(int j, const u32 x[8], const u32 y[8]) {
#pragma unroll
    for (int w = 0; w < 8; w++) { e[j][w] = x[w]; e[j][8 + w] = y[w]; }
  }
  __device__ __forceinline__ void load(u32 j, u32 x[8], u32 y[8]) const {
#pragma unroll
    for (int w = 0; w < 8; w++) { x[w] = e[j][w]; y[w] = e[j][8 + w]; }
  }
};

template <class F>
struct WarpInv {
  __device__ __forceinline__ void operator()(u32 zi[8], const u32 z[8]) { warp_batch_inv<F>(zi, z); }
};

template <class C>
__global__ void __launch_bounds__(VAR_THREADS, VAR_BLOCKS_PER_SM)
k_wmul_var(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ points, int points_fmt, size_t n, int out_fmt,
           uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  typedef typename C::F F;
  extern __shared__ __align__(16) u32 smem[];
  const unsigned lane_in_warp = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  LocalTab tab;
  uint8_t* io = reinterpret_cast<uint8_t*>(smem) + warp * VAR_IO_BYTES_PER_WARP;  // this warp's staging area
  const int osz = out_size(out_fmt);
  const size_t ntiles = (n + 31) / 32;
  const size_t gwarp = (size_t)blockIdx.x * (VAR_THREADS / 32) + warp, nwarps = (size_t)gridDim.x * (VAR_THREADS / 32);
  u32 slots[VAR_DEFER][VAR_SLOT_WORDS];  // thread-local (coalesced by construction), indexed dynamically
  u32 acc[8];
  int cnt = 0;
  size_t first_tile_of_batch = 0;
  F::one(acc);

  for (size_t tile = gwarp;; tile += nwarps) {
    const bool have = tile < ntiles;                      // warp-uniform
    if (have) {
      const size_t first = tile * 32;
      const size_t lane = first + lane_in_warp;
      const int cnt_lanes = (int)((n - first) < 32 ? (n - first) : 32);
      const bool live = (int)lane_in_warp < cnt_lanes;
      u32 k[8], px[8], py[8], st, ok = 0xFFFFFFFFu, idn = 0u;
      set_small(k, 1u);
      C::G(px, py);
      if (points_fmt == FEC_PT_SEC1_COMPRESSED) {
        __syncwarp();
        warp_copy(io, points + first * 33, (size_t)cnt_lanes * 33);
        __syncwarp();
        if (live) ok = sec1_decompress_id<C>(px, py, io + lane_in_warp * 33, idn);
        __syncwarp();
      } else if (live) {
        load_rec32_be(px, points + lane * 64);
        load_rec32_be(py, points + lane * 64 + 32);
        ok = affine_from_plain_id<C>(px, py, idn);
      }
      if (live) load_rec32_be(k, scalars + lane * 32);
      {
        u32 gx[8], gy[8];
        C::G(gx, gy);
        select8(px, ok, px, gx);
        select8(py, ok, py, gy);
      }
      Jac Q;
      u32 negm = 0u;
      if (C::CURVE_ID == 0) {
        // secp256k1: GLV needs no odd scalar; only the range / zero substitution of scalar_prepare
        u32 nn[8], one[8], zs[8];
        C::N(nn);
        set_small(one, 1u);
        const u32 bad = geq8(k, nn), zero = is_zero8(k);
        select8(k, bad | zero, one, k);
        st = (bad & LANE_SCALAR_RANGE) | (~bad & zero & LANE_INFINITY);
        wmul_var_glv_core(Q, zs, k, px, py, tab);
        F::mul(Q.Z, Q.Z, zs);
      } else {
        u32 K[8];
        WarpInv<F> inv;
        scalar_prepare<C>(K, negm, st, k);
        wmul_var_core<C>(Q, K, px, py, tab, inv);
      }
      st = lane_status(st, ok, idn);  // range > invalid point > identity (k = 0, or the identity as the input point)
      const u32 zz = F::is_zero(Q.Z);
      st |= (st == 0u ? zz & LANE_INFINITY : 0u);
      if (status != nullptr && live) status[lane] = (uint8_t)st;
      // defer: store (X, +-Y, Z', prefix) with Z' = 1 for dead lanes, extend the running product
      u32 ny[8], one[8], z[8];
      F::neg(ny, Q.Y);
      select8(Q.Y, negm, ny, Q.Y);
      F::one(one);
      select8(z, (st != 0u) ? 0xFFFFFFFFu : 0u, one, Q.Z);
      if (cnt == 0) first_tile_of_batch = tile;
      u32* s = slots[cnt];
#pragma unroll
      for (int i = 0; i < 8; i++) { s[i] = Q.X[i]; s[8 + i] = Q.Y[i]; s[16 + i] = z[i]; s[24 + i] = acc[i]; }
      s[32] = st;
      if (cnt == 0) copy8(acc, z); else F::mul(acc, acc, z);
      cnt++;
    }
    if (cnt == VAR_DEFER || (!have && cnt > 0)) {
      // one inversion for the whole batch of this warp, then unwind
      u32 inv[8];
      warp_batch_inv<F>(inv, acc);
#pragma unroll 1
      for (int t = cnt - 1; t >= 0; t--) {
        const u32* s = slots[t];
        Jac Q;
        u32 z[8], pre[8], zi[8], x[8], y[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { Q.X[i] = s[i]; Q.Y[i] = s[8 + i]; z[i] = s[16 + i]; pre[i] = s[24 + i]; }
        const u32 st = s[32];
        if (t > 0) { F::mul(zi, inv, pre); F::mul(inv, inv, z); } else { copy8(zi, inv); }
        jac_to_affine_plain<C>(x, y, Q, zi, 0u);
        const size_t tile_t = first_tile_of_batch + (size_t)t * nwarps;
        const size_t first = tile_t * 32;
        const int cnt_lanes = (int)((n - first) < 32 ? (n - first) : 32);
        __syncwarp();
        if ((int)lane_in_warp < cnt_lanes) sec1_encode(io + lane_in_warp * osz, out_fmt, x, y, st ? 0xFFFFFFFFu : 0u);
        __syncwarp();
        warp_copy(out + first * osz, io, (size_t)cnt_lanes * osz);
      }
      cnt = 0;
      F::one(acc);
    }
    if (!have) break;
  }
}

template <class C>
cudaError_t wmul_var_setup_t() {
  return cudaFuncSetAttribute(k_wmul_var<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, VAR_SMEM_BYTES);
}
template <class C>
cudaError_t wmul_var_launch_t(int sms, cudaStream_t st, const uint8_t* scalars, const uint8_t* points, int points_fmt, size_t n,
                              int out_fmt, uint8_t* out, uint8_t* status) {
  const int g = persistent_grid(sms, VAR_BLOCKS_PER_SM, (n + VAR_THREADS - 1) / VAR_THREADS);
  k_wmul_var<C><<<g, VAR_THREADS, VAR_SMEM_BYTES, st>>>(scalars, points, points_fmt, n, out_fmt, out, status);
  return cudaGetLastError();
}
cudaError_t wmul_var_setup_k1();
cudaError_t wmul_var_setup_p256();
cudaError_t wmul_var_launch_k1(int, cudaStream_t, const uint8_t*, const uint8_t*, int, size_t, int, uint8_t*, uint8_t*);
cudaError_t wmul_var_launch_p256(int, cudaStream_t, const uint8_t*, const uint8_t*, int, size_t, int, uint8_t*, uint8_t*);

}  // namespace fec
