// k_x25519.cu -- X25519 ladder kernel (config 4)
#include "kernels.cuh"

namespace fec {

// ======================================================================================
// kernel: X25519  (config 4)
// ======================================================================================
__global__ void __launch_bounds__(X_THREADS, X_BLOCKS_PER_SM)
k_x25519(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ us, size_t n, uint8_t* __restrict__ out) {
  typedef Fe25519 F;
  __shared__ u32 scratch[INV_SCRATCH_WORDS];
  const size_t ntiles = (n + X_THREADS - 1) / X_THREADS;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t lane = tile * X_THREADS + threadIdx.x;
    const bool live = lane < n;
    u32 k[8], u[8], x2[8], z2[8];
    set_small(k, 0u);
    set_small(u, 9u);
    if (live) {
      load_rec32_le(k, scalars + lane * 32);
      if (us != nullptr) load_rec32_le(u, us + lane * 32);
    }
    x25519_core(x2, z2, k, u);
    u32 z[8], zi[8], one[8], t[8], r[8];
    set_small(one, 1u);
    const u32 zz = F::is_zero(z2);
    select8(z, zz, one, z2);
    block_batch_inv<F, X_THREADS / 32>(zi, z, scratch);
    F::mul(t, x2, zi);
    F::canon(r, t);
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] &= ~zz;  // z2 == 0 -> 0, as x * 0^(p-2)
    if (live) store_rec32_le(out + lane * 32, r);
    __syncthreads();
  }
}


cudaError_t x25519_launch(int sms, cudaStream_t st, const uint8_t* scalars, const uint8_t* us, size_t n, uint8_t* out) {
  const int g = persistent_grid(sms, X_BLOCKS_PER_SM, (n + X_THREADS - 1) / X_THREADS);
  k_x25519<<<g, X_THREADS, 0, st>>>(scalars, us, n, out);
  return cudaGetLastError();
}

}  // namespace fec
