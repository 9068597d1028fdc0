// k_ed25519.cu -- Ed25519 public-key kernel (config 5b) and its Niels comb builder
#ifndef FEC_25519_NO_FOLD255
#define FEC_25519_FOLD255  // results of mul / sqr below 2^255 + 2^11: most additions of ed_madd_signed need no fold (fields.cuh, curves.cuh)
#endif
#include "kernels.cuh"

namespace fec {

// ======================================================================================
// kernel: Ed25519 public keys  (config 5b)
// ======================================================================================
// Warp-independent tiles with a deferred warp-level inversion, as k_wmul_fixed (k_wmul_fixed.cuh).
// CT = true (FORGE_EC_CT_TABLE=1): the constant-time table form -- niels_g is the 48 KiB 4-bit comb (lanes.cuh: ScanNiels), copied
// to shared memory by every block; 64 additions with masked scans instead of 16 with gathers by secret digits.
#ifndef FEC_ED_DEFER
#define FEC_ED_DEFER 64
#endif
constexpr int ED_DEFER = FEC_ED_DEFER;
template <bool CT>
__global__ void __launch_bounds__(ED_THREADS, ED_BLOCKS_PER_SM)
k_ed25519_pubkey(const uint8_t* __restrict__ in, int mode, const u32* __restrict__ niels_g, size_t n,
                 uint8_t* __restrict__ out) {
  // mode 0: 32-byte little-endian scalars as they are -> RFC 8032 encoding of s*B;  mode 1: seeds (SHA-512 + clamp) -> same;
  // mode 2: X25519 with the base point u = 9 (RFC 7748 6.1 key generation): clamp, s*B on the birationally equivalent
  //         Edwards curve, u = (Z + Y) / (Z - Y) -- 16 table additions instead of a 255-step ladder.  Z - Y != 0 because a
  //         clamped scalar is never a multiple of the group order.
  const int input_is_seed = mode == 1;
  typedef Fe25519 F;
  typename std::conditional<CT, ScanNiels, GlobalNiels>::type niels{niels_g};
  if (CT) {
    extern __shared__ __align__(16) u32 ct_tab[];
    for (int i = threadIdx.x; i < CT_NIELS_WORDS; i += ED_THREADS) ct_tab[i] = niels_g[i];
    __syncthreads();
    niels.base = ct_tab;
  }
  const unsigned lane_in_warp = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const size_t ntiles = (n + 31) / 32;
  const size_t gwarp = (size_t)blockIdx.x * (ED_THREADS / 32) + warp, nwarps = (size_t)gridDim.x * (ED_THREADS / 32);
  u32 slots[ED_DEFER][32];  // X, Y, Z, prefix product
  u32 acc[8];
  int cnt = 0;
  size_t first_tile_of_batch = 0;
  __shared__ u32 inv_scratch[INV_SCRATCH_WORDS];
  const bool small = small_batch_mode(ntiles, nwarps, 16);  // kernel-uniform
  set_small(acc, 1u);
  for (size_t tile = gwarp;; tile += nwarps) {
    const bool have = tile < ntiles;  // warp-uniform
    if (have) {
      const size_t lane = tile * 32 + lane_in_warp;
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
          if (mode == 2) {
            s[0] &= 0xFFFFFFF8u;
            s[7] = (s[7] & 0x7FFFFFFFu) | 0x40000000u;
          }
        }
      }
      ed_scalar_prepare(K, s);
      EdExt Q;
      if (CT) ed_comb_core<decltype(niels), CT_W>(Q, K, niels);
      else ed_comb_core(Q, K, niels);  // Z != 0 always (complete formulas)
      if (mode == 2) {            // (Y, Z) <- (Z + Y, Z - Y): the unwinding multiplies Y by 1 / Z
        u32 t[8];
        F::add(t, Q.Z, Q.Y);
        F::sub(Q.Z, Q.Z, Q.Y);
        copy8(Q.Y, t);
      }
      if (cnt == 0) first_tile_of_batch = tile;
      u32* sl = slots[cnt];
#pragma unroll
      for (int i = 0; i < 8; i++) { sl[i] = Q.X[i]; sl[8 + i] = Q.Y[i]; sl[16 + i] = Q.Z[i]; sl[24 + i] = acc[i]; }
      if (cnt == 0) copy8(acc, Q.Z); else F::mul(acc, acc, Q.Z);
      cnt++;
    }
    if (small ? !have : (cnt == ED_DEFER || (!have && cnt > 0))) {
      u32 inv[8];
      if (small) block_batch_inv<F, ED_THREADS / 32>(inv, acc, inv_scratch);  // every warp of the block, exactly once
      else warp_batch_inv<F>(inv, acc);
#pragma unroll 1
      for (int t = cnt - 1; t >= 0; t--) {
        const u32* sl = slots[t];
        EdExt Q;
        u32 z[8], pre[8], zi[8], o[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { Q.X[i] = sl[i]; Q.Y[i] = sl[8 + i]; z[i] = sl[16 + i]; pre[i] = sl[24 + i]; }
        if (t > 0) { F::mul(zi, inv, pre); F::mul(inv, inv, z); } else { copy8(zi, inv); }
        if (mode == 2) {
          u32 t[8];
          F::mul(t, Q.Y, zi);
          F::canon(o, t);
        } else {
          ed_encode(o, Q, zi);
        }
        const size_t lane = (first_tile_of_batch + (size_t)t * nwarps) * 32 + lane_in_warp;
        if (lane < n) store_rec32_le(out + lane * 32, o);
      }
      cnt = 0;
      set_small(acc, 1u);
    }
    if (!have) break;
  }
}

__global__ void __launch_bounds__(32) k_build_niels_bases(u32* __restrict__ bases) {
  if (threadIdx.x >= COMB_WINDOWS) return;
  EdExt Bi;
  ed_comb_base(Bi, (int)threadIdx.x);
  u32* o = bases + threadIdx.x * 32;
#pragma unroll
  for (int w = 0; w < 8; w++) { o[w] = Bi.X[w]; o[8 + w] = Bi.Y[w]; o[16 + w] = Bi.Z[w]; o[24 + w] = Bi.T[w]; }
}
__global__ void __launch_bounds__(128) k_build_niels(u32* __restrict__ niels_g, const u32* __restrict__ bases) {
  const size_t e = (size_t)blockIdx.x * 128 + threadIdx.x;
  const int i = (int)(e / COMB_ENTRIES), j = (int)(e % COMB_ENTRIES);
  EdExt Bi;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    Bi.X[w] = bases[i * 32 + w]; Bi.Y[w] = bases[i * 32 + 8 + w]; Bi.Z[w] = bases[i * 32 + 16 + w]; Bi.T[w] = bases[i * 32 + 24 + w];
  }
  u32 a[8], b[8], c[8];
  ed_comb_entry(a, b, c, Bi, j);
  u32* o = niels_g + e * 24;
#pragma unroll
  for (int w = 0; w < 8; w++) { o[w] = a[w]; o[8 + w] = b[w]; o[16 + w] = c[w]; }
}


// the 4-bit Niels comb of the constant-time form: one thread per entry (i, j) = (2j+1) 2^(4i) B
__global__ void __launch_bounds__(128) k_build_niels_ct(u32* __restrict__ niels4) {
  const int e = blockIdx.x * 128 + threadIdx.x;
  const int i = e / CT_ENTRIES, j = e % CT_ENTRIES;
  EdExt Bi;
  ed_comb_base<CT_W>(Bi, i);
  u32 a[8], b[8], c[8];
  ed_comb_entry<CT_W>(a, b, c, Bi, j);
  u32* o = niels4 + (size_t)e * 24;
#pragma unroll
  for (int w = 0; w < 8; w++) { o[w] = a[w]; o[8 + w] = b[w]; o[16 + w] = c[w]; }
}

cudaError_t ed25519_setup() {
  return cudaFuncSetAttribute(k_ed25519_pubkey<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_NIELS_WORDS * 4);
}
cudaError_t ed25519_ct_launch(int sms, cudaStream_t st, const uint8_t* in, int mode, const u32* niels4, size_t n, uint8_t* out) {
  const int g = persistent_grid(sms, ED_BLOCKS_PER_SM, (n + ED_THREADS - 1) / ED_THREADS);
  k_ed25519_pubkey<true><<<g, ED_THREADS, CT_NIELS_WORDS * 4, st>>>(in, mode, niels4, n, out);
  return cudaGetLastError();
}
cudaError_t build_niels_ct_launch(cudaStream_t st, u32* niels4) {
  k_build_niels_ct<<<CT_WINDOWS * CT_ENTRIES / 128, 128, 0, st>>>(niels4);
  return cudaGetLastError();
}
cudaError_t ed25519_launch(int sms, cudaStream_t st, const uint8_t* in, int mode, const u32* niels, size_t n,
                           uint8_t* out) {
  const int g = persistent_grid(sms, ED_BLOCKS_PER_SM, (n + ED_THREADS - 1) / ED_THREADS);
  k_ed25519_pubkey<false><<<g, ED_THREADS, 0, st>>>(in, mode, niels, n, out);
  return cudaGetLastError();
}
cudaError_t build_niels_launch(cudaStream_t st, u32* niels, u32* bases) {
  k_build_niels_bases<<<1, 32, 0, st>>>(bases);
  k_build_niels<<<(unsigned)((size_t)COMB_WINDOWS * COMB_ENTRIES / 128), 128, 0, st>>>(niels, bases);
  return cudaGetLastError();
}

}  // namespace fec
