// k_wmul_fixed.cuh -- fixed-base kernel template and comb-table builder; instantiated per curve
#pragma once
#include "kernels.cuh"

namespace fec {

// ======================================================================================
// kernel: fixed-base k*G  (configs 1, 5a)
// ======================================================================================

template <class C>
__global__ void __launch_bounds__(FIX_THREADS, FIX_BLOCKS_PER_SM)
k_wmul_fixed(const uint8_t* __restrict__ scalars, const u32* __restrict__ comb_g, size_t n, int out_fmt,
             uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  typedef typename C::F F;
  extern __shared__ __align__(16) u32 smem[];
  uint8_t* io = reinterpret_cast<uint8_t*>(smem);
  BlockInv<F, FIX_THREADS / 32> inv{smem + FIX_IO_BYTES / 4};
  GlobalComb comb{comb_g};
  const int osz = out_size(out_fmt);
  const size_t ntiles = (n + FIX_THREADS - 1) / FIX_THREADS;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t first = tile * FIX_THREADS;
    const size_t lane = first + threadIdx.x;
    const int cnt = (int)((n - first) < (size_t)FIX_THREADS ? (n - first) : (size_t)FIX_THREADS);
    const bool live = threadIdx.x < cnt;
    u32 k[8], K[8], negm, st;
    set_small(k, 1u);
    if (live) load_rec32_be(k, scalars + lane * 32);
    scalar_prepare<C>(K, negm, st, k);
    Jac Q;
    wmul_fixed_core<C>(Q, K, comb);
    u32 z[8], zi[8], one[8], x[8], y[8];
    set_small(one, 1u);
    const u32 zz = F::is_zero(Q.Z);
    select8(z, zz, one, Q.Z);
    inv(zi, z);
    jac_to_affine<C>(x, y, Q, zi, negm);
    st |= (st == 0u ? zz & LANE_INFINITY : 0u);
    if (live) sec1_encode(io + threadIdx.x * osz, out_fmt, x, y, st ? 0xFFFFFFFFu : 0u);
    __syncthreads();
    block_copy(out + first * osz, io, (size_t)cnt * osz);
    if (status != nullptr && live) status[lane] = (uint8_t)st;
    __syncthreads();
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

template <class C>
cudaError_t wmul_fixed_setup_t() {
  return cudaFuncSetAttribute(k_wmul_fixed<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, FIX_SMEM_BYTES);
}
template <class C>
cudaError_t wmul_fixed_launch_t(int sms, cudaStream_t st, const uint8_t* scalars, const u32* comb, size_t n, int out_fmt,
                                uint8_t* out, uint8_t* status) {
  const int g = persistent_grid(sms, FIX_BLOCKS_PER_SM, (n + FIX_THREADS - 1) / FIX_THREADS);
  k_wmul_fixed<C><<<g, FIX_THREADS, FIX_SMEM_BYTES, st>>>(scalars, comb, n, out_fmt, out, status);
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
cudaError_t build_comb_launch_k1(cudaStream_t, u32*, u32*);
cudaError_t build_comb_launch_p256(cudaStream_t, u32*, u32*);

}  // namespace fec
