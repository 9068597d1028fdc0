// k_wmul_fixed_vm.cuh -- fixed-base k*G on the accumulator machine (configs 1, 5a); instantiated per curve in
// k_wmul_fixed_vm_{k1,p256}.cu.  Same table, digits, deferred warp-level inversion and encoders as k_wmul_fixed.cuh;
// the running point lives in shared-memory slots (vm.cuh) and the table entries arrive there by cp.async.
#pragma once
#include "vm_points.cuh"
#include "kernels.cuh"

namespace fec {

constexpr int FIXVM_THREADS = 128;
#ifndef FEC_FIXVM_BLOCKS
#define FEC_FIXVM_BLOCKS 5
#endif
constexpr int FIXVM_BLOCKS = FEC_FIXVM_BLOCKS;
#ifndef FEC_FIXVM_DEFER
#define FEC_FIXVM_DEFER 64
#endif
constexpr int FIXVM_DEFER = FEC_FIXVM_DEFER;
constexpr int FIXVM_SLOT_BYTES = VS_COUNT * 32 * FIXVM_THREADS;
constexpr int FIXVM_IO_BYTES = (FIXVM_THREADS / 32) * 65 * 32;
constexpr int FIXVM_SMEM = FIXVM_SLOT_BYTES + FIXVM_IO_BYTES;  // slots + staging: 45 184 B, five blocks per SM (one more KB would leave four)

struct GlobalCombEntry {
  const u32* base;
  __device__ __forceinline__ const u32* operator()(int i, u32 idx) const { return base + ((size_t)i * COMB_ENTRIES + idx) * 16; }
};

template <class C>
__global__ void __launch_bounds__(FIXVM_THREADS, FIXVM_BLOCKS)
k_wmul_fixed_vm(const uint8_t* __restrict__ scalars, const u32* __restrict__ comb_g, size_t n, int out_fmt,
                uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  typedef typename C::F F;
  typedef Vm<F, FIXVM_THREADS> V;
  extern __shared__ __align__(16) u32 smem[];
  const unsigned lane_in_warp = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const saddr S = (saddr)__cvta_generic_to_shared(smem) + threadIdx.x * 16u;
  uint8_t* io = reinterpret_cast<uint8_t*>(smem) + FIXVM_SLOT_BYTES + warp * (65 * 32);
  const GlobalCombEntry entry{comb_g};
  const int osz = out_size(out_fmt);
  const size_t ntiles = (n + 31) / 32;
  const size_t gwarp = (size_t)blockIdx.x * (FIXVM_THREADS / 32) + warp, nwarps = (size_t)gridDim.x * (FIXVM_THREADS / 32);
  u32 slots[FIXVM_DEFER][36];  // X, Y, Z, prefix product, status
  u32 acc[8];
  int cnt = 0;
  size_t first_tile_of_batch = 0;
  const bool small = small_batch_mode(ntiles, nwarps, 16);  // kernel-uniform
  u32* inv_scratch = smem;  // the slot area: free once every warp of the block has left its last tile (barrier below)
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
      wmul_fixed_vm_core<C, FIXVM_THREADS>(S, K, entry);
      u32* s = slots[cnt];
      u32 z[8], y[8], ny[8], one[8];
      V::ldv(z, V::slot(S, VS_Z));
      const u32 zz = F::is_zero(z);
      st |= (st == 0u ? zz & LANE_INFINITY : 0u);
      if (status != nullptr && live) status[lane] = (uint8_t)st;
      F::one(one);
      select8(z, (st != 0u) ? 0xFFFFFFFFu : 0u, one, z);
      V::ldv(y, V::slot(S, VS_Y));
      F::neg(ny, y);
      select8(y, negm, ny, y);
      if (cnt == 0) first_tile_of_batch = tile;
#pragma unroll
      for (int i = 0; i < 8; i++) { s[8 + i] = y[i]; s[16 + i] = z[i]; s[24 + i] = acc[i]; }
      V::ldv(y, V::slot(S, VS_X));
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = y[i];
      s[32] = st;
      if (cnt == 0) copy8(acc, z); else F::mul(acc, acc, z);
      cnt++;
    }
    if (small ? !have : (cnt == FIXVM_DEFER || (!have && cnt > 0))) {
      u32 inv[8];
      if (small) {  // every warp of the block, exactly once
        __syncthreads();  // nobody is still working in the slots the scratch area overlays
        block_batch_inv<F, FIXVM_THREADS / 32>(inv, acc, inv_scratch);
      }
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

template <class C>
cudaError_t wmul_fixed_vm_setup_t() {
  cudaError_t e = cudaFuncSetAttribute(k_wmul_fixed_vm<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, FIXVM_SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_wmul_fixed_vm<C>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
template <class C>
cudaError_t wmul_fixed_vm_launch_t(int sms, cudaStream_t st, const uint8_t* scalars, const u32* comb, size_t n, int out_fmt,
                                   uint8_t* out, uint8_t* status) {
  const int g = persistent_grid(sms, FIXVM_BLOCKS, (n + FIXVM_THREADS - 1) / FIXVM_THREADS);
  k_wmul_fixed_vm<C><<<g, FIXVM_THREADS, FIXVM_SMEM, st>>>(scalars, comb, n, out_fmt, out, status);
  return cudaGetLastError();
}

}  // namespace fec
