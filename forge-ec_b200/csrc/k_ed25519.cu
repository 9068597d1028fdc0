// k_ed25519.cu -- Ed25519 public-key kernel (config 5b) and its Niels comb builder
#include "kernels.cuh"

namespace fec {

// ======================================================================================
// kernel: Ed25519 public keys  (config 5b)
// ======================================================================================
__global__ void __launch_bounds__(ED_THREADS, ED_BLOCKS_PER_SM)
k_ed25519_pubkey(const uint8_t* __restrict__ in, int input_is_seed, const u32* __restrict__ niels_g, size_t n,
                 uint8_t* __restrict__ out) {
  typedef Fe25519 F;
  extern __shared__ __align__(16) u32 smem[];
  u32* scratch = smem + NIELS_WORDS;
  for (int i = threadIdx.x; i < NIELS_WORDS / 4; i += ED_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(niels_g)[i];
  __syncthreads();
  SmemNiels niels{smem};
  const size_t ntiles = (n + ED_THREADS - 1) / ED_THREADS;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t lane = tile * ED_THREADS + threadIdx.x;
    const bool live = lane < n;
    u32 s[8], K[8];
    set_small(s, 1u);
    if (live) {
      if (input_is_seed) {
        u32 w[8];
        load_rec32_le(w, in + lane * 32);
        uint8_t seed[32];
#pragma unroll
        for (int i = 0; i < 32; i++) seed[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
        ed25519_expand_seed(s, seed);
      } else {
        load_rec32_le(s, in + lane * 32);
      }
    }
    ed_scalar_prepare(K, s);
    EdExt Q;
    ed_comb_core(Q, K, niels);
    u32 zi[8], o[8];
    block_batch_inv<F, ED_THREADS / 32>(zi, Q.Z, scratch);  // Z != 0 always (complete formulas)
    ed_encode(o, Q, zi);
    if (live) store_rec32_le(out + lane * 32, o);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128) k_build_niels(u32* __restrict__ niels_g) {
  const int e = blockIdx.x * 128 + threadIdx.x;
  u32 a[8], b[8], c[8];
  ed_comb_entry(a, b, c, e >> 3, e & 7);
  u32* o = niels_g + (size_t)e * NIELS_STRIDE;
#pragma unroll
  for (int w = 0; w < 8; w++) { o[w] = a[w]; o[8 + w] = b[w]; o[16 + w] = c[w]; }
  o[24] = o[25] = o[26] = o[27] = 0u;
}


cudaError_t ed25519_setup() {
  return cudaFuncSetAttribute(k_ed25519_pubkey, cudaFuncAttributeMaxDynamicSharedMemorySize, ED_SMEM_BYTES);
}
cudaError_t ed25519_launch(int sms, cudaStream_t st, const uint8_t* in, int input_is_seed, const u32* niels, size_t n,
                           uint8_t* out) {
  const int g = persistent_grid(sms, ED_BLOCKS_PER_SM, (n + ED_THREADS - 1) / ED_THREADS);
  k_ed25519_pubkey<<<g, ED_THREADS, ED_SMEM_BYTES, st>>>(in, input_is_seed, niels, n, out);
  return cudaGetLastError();
}
cudaError_t build_niels_launch(cudaStream_t st, u32* niels) {
  k_build_niels<<<4, 128, 0, st>>>(niels);
  return cudaGetLastError();
}

}  // namespace fec
