// k_wmul_fixed.cuh -- fixed-base kernel template and comb-table builder; instantiated per curve
#pragma once
#include <type_traits>
#include "kernels.cuh"

namespace fec {

// ======================================================================================
// kernel: fixed-base k*G  (configs 1, 5a)
// ======================================================================================

// Every WARP is independent (no __syncthreads): it walks tiles of 32 consecutive lanes, keeps the Jacobian results of
// FIX_DEFER tiles in thread-local memory together with the running product of their Z, then performs ONE warp-level
// inversion for all 32 x FIX_DEFER results (Montgomery's trick along the tile axis), normalises, encodes and writes
// them.  With 15 additions per scalar the inversion would otherwise be a third of the work.
// r2: 64 tiles per inversion instead of 16 (the Fermat chain is 270 field operations per WARP: 17 per tile at 16, 4 at 64 against the
// ~170 of a tile's additions): k*G 511 -> 543 M/s, Ed25519 keygen 588 -> 622 M/s at 2^22 lanes, nothing lost at 2^18
// (profiles/r2v_defer_ab.log).  The price is thread-local memory: 64 x 36 words = 9 KiB per thread.
#ifndef FEC_FIX_DEFER
#define FEC_FIX_DEFER 64
#endif
constexpr int FIX_DEFER = FEC_FIX_DEFER;
constexpr int FIX_SLOT_WORDS = 36;  // X, Y, Z, prefix (8 words each) + status + padding

// CT = true: the constant-time table form (ScanComb, lanes.cuh): comb_g is the 32 KiB 4-bit comb, copied to shared memory first
template <class C, bool CT = false>
__global__ void __launch_bounds__(FIX_THREADS, FIX_BLOCKS_PER_SM)
k_wmul_fixed(const uint8_t* __restrict__ scalars, const u32* __restrict__ comb_g, size_t n, int out_fmt,
             uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  typedef typename C::F F;
  extern __shared__ __align__(16) u32 smem[];
  const unsigned lane_in_warp = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint8_t* io = reinterpret_cast<uint8_t*>(smem) + warp * FIX_IO_BYTES_PER_WARP;
  typename std::conditional<CT, ScanComb, GlobalComb>::type comb{comb_g};
  if (CT) {
    u32* tab = smem + FIX_SMEM_BYTES / 4;
    for (int i = threadIdx.x; i < CT_TABLE_WORDS; i += FIX_THREADS) tab[i] = comb_g[i];
    __syncthreads();
    comb.base = tab;
  }
  const int osz = out_size(out_fmt);
  const size_t ntiles = (n + 31) / 32;
  const size_t gwarp = (size_t)blockIdx.x * (FIX_THREADS / 32) + warp, nwarps = (size_t)gridDim.x * (FIX_THREADS / 32);
  u32 slots[FIX_DEFER][FIX_SLOT_WORDS];
  u32 acc[8];
  int cnt = 0;
  size_t first_tile_of_batch = 0;
  const bool small = small_batch_mode(ntiles, nwarps, 16);  // kernel-uniform
  u32* inv_scratch = smem + (FIX_THREADS / 32) * FIX_IO_BYTES_PER_WARP / 4;
  F::one(acc);
  for (size_t tile = gwarp;; tile += nwarps) {
    const bool have = tile < ntiles;  // warp-uniform
    if (have) {
      const size_t lane = tile * 32 + lane_in_warp;
      const bool live = lane < n;
      u32 k[8], K[8], negm, st;
      set_small(k, 1u);
      if (live) load_rec32_be(k, scalars + lane * 32);
      scalar_prepare<C>(K, negm, st, k);
      Jac Q;
      if (CT) wmul_fixed_core<C, decltype(comb), CT_W>(Q, K, comb);
      else wmul_fixed_core<C>(Q, K, comb);
      const u32 zz = F::is_zero(Q.Z);
      st |= (st == 0u ? zz & LANE_INFINITY : 0u);
      if (status != nullptr && live) status[lane] = (uint8_t)st;
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
    if (small ? !have : (cnt == FIX_DEFER || (!have && cnt > 0))) {
      u32 inv[8];
      if (small) block_batch_inv<F, FIX_THREADS / 32>(inv, acc, inv_scratch);  // every warp of the block, exactly once
      else warp_batch_inv<F>(inv, acc);
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
        const size_t first = (first_tile_of_batch + (size_t)t * nwarps) * 32;
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

// comb table construction (context set-up): the 16 window bases, then one thread per entry
template <class C>
__global__ void __launch_bounds__(32) k_build_comb_bases(u32* __restrict__ bases) {
  if (threadIdx.x >= COMB_WINDOWS) return;
  u32 bx[8], by[8];
  comb_base<C>(bx, by, (int)threadIdx.x);
  u32* o = bases + threadIdx.x * 16;
#pragma unroll
  for (int w = 0; w < 8; w++) { o[w] = bx[w]; o[8 + w] = by[w]; }
}
template <class C>
__global__ void __launch_bounds__(128) k_build_comb(u32* __restrict__ comb_g, const u32* __restrict__ bases) {
  typedef typename C::F F;
  __shared__ u32 scratch[INV_SCRATCH_WORDS];
  const size_t e = (size_t)blockIdx.x * 128 + threadIdx.x;  // (window, entry); the grid covers the table exactly
  const int i = (int)(e / COMB_ENTRIES), j = (int)(e % COMB_ENTRIES);
  u32 bx[8], by[8];
#pragma unroll
  for (int w = 0; w < 8; w++) { bx[w] = bases[i * 16 + w]; by[w] = bases[i * 16 + 8 + w]; }
  Jac R;
  comb_entry_jac<C>(R, bx, by, j);
  u32 zi[8], x[8], y[8];
  block_batch_inv<F, 4>(zi, R.Z, scratch);
  jac_to_affine<C>(x, y, R, zi, 0u);
  u32* o = comb_g + e * 16;
#pragma unroll
  for (int w = 0; w < 8; w++) { o[w] = x[w]; o[8 + w] = y[w]; }
}

// the 4-bit comb of the constant-time form: one thread per entry (i, j) = (2j+1) 2^(4i) G, affine (its own Fermat inversion)
template <class C>
__global__ void __launch_bounds__(128) k_build_comb_ct(u32* __restrict__ comb4) {
  typedef typename C::F F;
  const int e = blockIdx.x * 128 + threadIdx.x;  // the grid covers the 512 entries exactly
  const int i = e / CT_ENTRIES, j = e % CT_ENTRIES;
  u32 bx[8], by[8];
  comb_base<C, CT_W>(bx, by, i);
  Jac R;
  comb_entry_jac<C>(R, bx, by, j);
  u32 zi[8], x[8], y[8];
  F::inv(zi, R.Z);
  jac_to_affine<C>(x, y, R, zi, 0u);
  u32* o = comb4 + (size_t)e * 16;
#pragma unroll
  for (int w = 0; w < 8; w++) { o[w] = x[w]; o[8 + w] = y[w]; }
}
constexpr int FIX_CT_SMEM_BYTES = FIX_SMEM_BYTES + CT_TABLE_WORDS * 4;
template <class C>
cudaError_t wmul_fixed_setup_t() {
  cudaError_t e = cudaFuncSetAttribute(k_wmul_fixed<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FIX_CT_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_wmul_fixed<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FIX_SMEM_BYTES);
}
template <class C>
cudaError_t wmul_fixed_launch_t(int sms, cudaStream_t st, const uint8_t* scalars, const u32* comb, size_t n, int out_fmt,
                                uint8_t* out, uint8_t* status) {
  const int g = persistent_grid(sms, FIX_BLOCKS_PER_SM, (n + FIX_THREADS - 1) / FIX_THREADS);
  k_wmul_fixed<C, false><<<g, FIX_THREADS, FIX_SMEM_BYTES, st>>>(scalars, comb, n, out_fmt, out, status);
  return cudaGetLastError();
}
template <class C>
cudaError_t wmul_fixed_ct_launch_t(int sms, cudaStream_t st, const uint8_t* scalars, const u32* comb4, size_t n, int out_fmt,
                                   uint8_t* out, uint8_t* status) {
  const int g = persistent_grid(sms, FIX_BLOCKS_PER_SM, (n + FIX_THREADS - 1) / FIX_THREADS);
  k_wmul_fixed<C, true><<<g, FIX_THREADS, FIX_CT_SMEM_BYTES, st>>>(scalars, comb4, n, out_fmt, out, status);
  return cudaGetLastError();
}
template <class C>
cudaError_t build_comb_ct_launch_t(cudaStream_t st, u32* comb4) {
  k_build_comb_ct<C><<<CT_WINDOWS * CT_ENTRIES / 128, 128, 0, st>>>(comb4);
  return cudaGetLastError();
}
template <class C>
cudaError_t build_comb_launch_t(cudaStream_t st, u32* comb, u32* bases) {
  k_build_comb_bases<C><<<1, 32, 0, st>>>(bases);
  k_build_comb<C><<<(unsigned)((size_t)COMB_WINDOWS * COMB_ENTRIES / 128), 128, 0, st>>>(comb, bases);
  return cudaGetLastError();
}
cudaError_t wmul_fixed_setup_k1();
cudaError_t wmul_fixed_setup_p256();
cudaError_t wmul_fixed_launch_k1(int, cudaStream_t, const uint8_t*, const u32*, size_t, int, uint8_t*, uint8_t*);
cudaError_t wmul_fixed_launch_p256(int, cudaStream_t, const uint8_t*, const u32*, size_t, int, uint8_t*, uint8_t*);
cudaError_t wmul_fixed_vm_setup_k1();
cudaError_t wmul_fixed_vm_setup_p256();
cudaError_t wmul_fixed_vm_launch_k1(int, cudaStream_t, const uint8_t*, const u32*, size_t, int, uint8_t*, uint8_t*);
cudaError_t wmul_fixed_vm_launch_p256(int, cudaStream_t, const uint8_t*, const u32*, size_t, int, uint8_t*, uint8_t*);
cudaError_t build_comb_launch_k1(cudaStream_t, u32*, u32*);
cudaError_t build_comb_launch_p256(cudaStream_t, u32*, u32*);
cudaError_t wmul_fixed_ct_launch_k1(int, cudaStream_t, const uint8_t*, const u32*, size_t, int, uint8_t*, uint8_t*);
cudaError_t wmul_fixed_ct_launch_p256(int, cudaStream_t, const uint8_t*, const u32*, size_t, int, uint8_t*, uint8_t*);
cudaError_t build_comb_ct_launch_k1(cudaStream_t, u32*);
cudaError_t build_comb_ct_launch_p256(cudaStream_t, u32*);

}  // namespace fec
