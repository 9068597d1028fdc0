// k_x25519.cu -- X25519 ladder kernel (config 4)
#define FEC_25519_EVOD_REDUCE  // Fe25519::mul_inl reduces the product from its unmerged accumulators (fields.cuh)
#ifndef FEC_25519_NO_FOLD255
#define FEC_25519_FOLD255      // results of mul / sqr below 2^255 + 2^11: the ladder's additions need no fold (fields.cuh, curves.cuh)
#endif
#include "kernels.cuh"

namespace fec {

// ======================================================================================
// kernel: X25519  (config 4)
// ======================================================================================
// Warp-independent tiles of 32 lanes; the final x2 / z2 of X_DEFER tiles share ONE warp-level inversion (Montgomery's
// trick along the tile axis), so no warp ever waits at a barrier while another one runs the Fermat chain.
#ifndef FEC_X_DEFER
#define FEC_X_DEFER 32
#endif
constexpr int X_DEFER = FEC_X_DEFER;
__global__ void __launch_bounds__(X_THREADS, X_BLOCKS_PER_SM)
k_x25519(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ us, size_t n, uint8_t* __restrict__ out) {
  typedef Fe25519 F;
  const unsigned lane_in_warp = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const size_t ntiles = (n + 31) / 32;
  const size_t gwarp = (size_t)blockIdx.x * (X_THREADS / 32) + warp, nwarps = (size_t)gridDim.x * (X_THREADS / 32);
  u32 slots[X_DEFER][28];  // x2, z2 (1 where it was 0), prefix product, zero flag
  u32 acc[8];
  int cnt = 0;
  size_t first_tile_of_batch = 0;
  __shared__ u32 inv_scratch[INV_SCRATCH_WORDS];
  const bool small = small_batch_mode(ntiles, nwarps, 8);  // kernel-uniform
  set_small(acc, 1u);
  for (size_t tile = gwarp;; tile += nwarps) {
    const bool have = tile < ntiles;  // warp-uniform
    if (have) {
      const size_t lane = tile * 32 + lane_in_warp;
      const bool live = lane < n;
      u32 k[8], u[8], x2[8], z2[8];
      set_small(k, 0u);
      set_small(u, 9u);
      if (live) {
        load_rec32_le(k, scalars + lane * 32);
        if (us != nullptr) load_rec32_le(u, us + lane * 32);
      }
      x25519_core(x2, z2, k, u);
      u32 z[8], one[8];
      set_small(one, 1u);
      const u32 zz = F::is_zero(z2);
      select8(z, zz, one, z2);
      if (cnt == 0) first_tile_of_batch = tile;
      u32* s = slots[cnt];
#pragma unroll
      for (int i = 0; i < 8; i++) { s[i] = x2[i]; s[8 + i] = z[i]; s[16 + i] = acc[i]; }
      s[24] = zz;
      if (cnt == 0) copy8(acc, z); else F::mul(acc, acc, z);
      cnt++;
    }
    if (small ? !have : (cnt == X_DEFER || (!have && cnt > 0))) {
      u32 inv[8];
      if (small) block_batch_inv<F, X_THREADS / 32>(inv, acc, inv_scratch);  // every warp of the block, exactly once
      else warp_batch_inv<F>(inv, acc);
#pragma unroll 1
      for (int t = cnt - 1; t >= 0; t--) {
        const u32* s = slots[t];
        u32 x2[8], z[8], pre[8], zi[8], tt[8], r[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { x2[i] = s[i]; z[i] = s[8 + i]; pre[i] = s[16 + i]; }
        const u32 zz = s[24];
        if (t > 0) { F::mul(zi, inv, pre); F::mul(inv, inv, z); } else { copy8(zi, inv); }
        F::mul(tt, x2, zi);
        F::canon(r, tt);
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] &= ~zz;  // z2 == 0 -> 0, as x * 0^(p-2)
        const size_t lane = (first_tile_of_batch + (size_t)t * nwarps) * 32 + lane_in_warp;
        if (lane < n) store_rec32_le(out + lane * 32, r);
      }
      cnt = 0;
      set_small(acc, 1u);
    }
    if (!have) break;
  }
}


cudaError_t x25519_launch(int sms, cudaStream_t st, const uint8_t* scalars, const uint8_t* us, size_t n, uint8_t* out) {
  const int g = persistent_grid(sms, X_BLOCKS_PER_SM, (n + X_THREADS - 1) / X_THREADS);
  k_x25519<<<g, X_THREADS, 0, st>>>(scalars, us, n, out);
  return cudaGetLastError();
}

}  // namespace fec
