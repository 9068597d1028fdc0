// kernels.cuh -- shared device helpers, tile geometry and the launcher interface between the
// kernel translation units (k_*.cu, one per kernel family so they compile in parallel) and the
// host side of the C ABI (forge_ec_cuda.cu).
//
// Kernel shape (all of them): persistent blocks (grid = SMs x resident blocks), each block walks
// tiles of THREADS consecutive lanes; one lane = one thread, all field elements in registers.
// Inputs are 16-byte-aligned 32/64-byte records read with 128-bit loads; 33/65-byte records are
// staged through shared memory so that global traffic stays word-coalesced.  The only
// cross-thread step is the block-level simultaneous inversion (curves.cuh).
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>

#include "../../include/forge_ec_cuda.h"
#include "lanes.cuh"

namespace fec {

// ======================================================================================
// small device helpers
// ======================================================================================
__device__ __forceinline__ void load_rec32_le(u32 r[8], const uint8_t* p) {  // p 16-byte aligned
  const uint4 a = reinterpret_cast<const uint4*>(p)[0], b = reinterpret_cast<const uint4*>(p)[1];
  r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}
__device__ __forceinline__ void load_rec32_be(u32 r[8], const uint8_t* p) {
  u32 t[8];
  load_rec32_le(t, p);
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = __byte_perm(t[7 - i], 0, 0x0123);
}
__device__ __forceinline__ void store_rec32_le(uint8_t* p, const u32 r[8]) {
  reinterpret_cast<uint4*>(p)[0] = make_uint4(r[0], r[1], r[2], r[3]);
  reinterpret_cast<uint4*>(p)[1] = make_uint4(r[4], r[5], r[6], r[7]);
}
// block-cooperative copy, word-coalesced; src/dst 4-byte aligned
__device__ __forceinline__ void block_copy(uint8_t* dst, const uint8_t* src, size_t bytes) {
  const size_t nw = bytes >> 2;
  const u32* s = reinterpret_cast<const u32*>(src);
  u32* d = reinterpret_cast<u32*>(dst);
  for (size_t i = threadIdx.x; i < nw; i += blockDim.x) d[i] = s[i];
  for (size_t i = (nw << 2) + threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
}

// warp-cooperative copy, word-coalesced; src/dst 4-byte aligned; callers bracket it with __syncwarp()
__device__ __forceinline__ void warp_copy(uint8_t* dst, const uint8_t* src, size_t bytes) {
  const unsigned l = threadIdx.x & 31u;
  const size_t nw = bytes >> 2;
  const u32* s = reinterpret_cast<const u32*>(src);
  u32* d = reinterpret_cast<u32*>(dst);
  for (size_t i = l; i < nw; i += 32) d[i] = s[i];
  for (size_t i = (nw << 2) + l; i < bytes; i += 32) dst[i] = src[i];
}

__host__ __device__ constexpr int out_size(int fmt) { return fmt == 0 ? 33 : ((fmt == 1 || fmt == 4) ? 65 : (fmt == 2 ? 64 : 32)); }

// per-lane table of 8 affine points in shared memory, word-interleaved across the block:
// word w of entry j of thread t at [(j*16 + w) * THREADS + t]  (conflict-free for any digits)
template <int THREADS>
struct SmemTab {
  u32* base;
  __device__ __forceinline__ void store(int j, const u32 x[8], const u32 y[8]) {
    u32* e = base + (size_t)j * 16 * THREADS + threadIdx.x;
#pragma unroll
    for (int w = 0; w < 8; w++) { e[w * THREADS] = x[w]; e[(8 + w) * THREADS] = y[w]; }
  }
  __device__ __forceinline__ void load(u32 j, u32 x[8], u32 y[8]) const {
    const u32* e = base + (size_t)j * 16 * THREADS + threadIdx.x;
#pragma unroll
    for (int w = 0; w < 8; w++) { x[w] = e[w * THREADS]; y[w] = e[(8 + w) * THREADS]; }
  }
};
template <class F, int WARPS>
struct BlockInv {
  u32* scratch;
  __device__ __forceinline__ void operator()(u32 zi[8], const u32 z[8]) { block_batch_inv<F, WARPS>(zi, z, scratch); }
};
// fixed-base tables in global memory (L2-resident): entry (i,j) = (2j+1) 2^(16 i) G at word ((i * COMB_ENTRIES) + j) * 16
// (affine X, Y; four 128-bit loads), Niels entries (y+x, y-x, 2dxy) at stride 24 words (six 128-bit loads)
constexpr size_t COMB_WORDS = (size_t)COMB_WINDOWS * COMB_ENTRIES * 16;
struct GlobalComb {
  const u32* base;
  __device__ __forceinline__ void operator()(int i, u32 idx, u32 x[8], u32 y[8]) const {
    const uint4* e = reinterpret_cast<const uint4*>(base + ((size_t)i * COMB_ENTRIES + idx) * 16);
    const uint4 a = __ldg(e), b = __ldg(e + 1), c = __ldg(e + 2), d = __ldg(e + 3);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    y[0] = c.x; y[1] = c.y; y[2] = c.z; y[3] = c.w; y[4] = d.x; y[5] = d.y; y[6] = d.z; y[7] = d.w;
  }
};
constexpr size_t NIELS_WORDS = (size_t)COMB_WINDOWS * COMB_ENTRIES * 24;
struct GlobalNiels {
  const u32* base;
  __device__ __forceinline__ void operator()(int i, u32 idx, u32 a[8], u32 b[8], u32 c[8]) const {
    const uint4* e = reinterpret_cast<const uint4*>(base + ((size_t)i * COMB_ENTRIES + idx) * 24);
    const uint4 v0 = __ldg(e), v1 = __ldg(e + 1), v2 = __ldg(e + 2), v3 = __ldg(e + 3), v4 = __ldg(e + 4), v5 = __ldg(e + 5);
    a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w; a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
    b[0] = v2.x; b[1] = v2.y; b[2] = v2.z; b[3] = v2.w; b[4] = v3.x; b[5] = v3.y; b[6] = v3.z; b[7] = v3.w;
    c[0] = v4.x; c[1] = v4.y; c[2] = v4.z; c[3] = v4.w; c[4] = v5.x; c[5] = v5.y; c[6] = v5.z; c[7] = v5.w;
  }
};

// ---- tile geometry ----------------------------------------------------------------------
constexpr int INV_SCRATCH_WORDS = 2 * 16 * 8 + 8;  // block_batch_inv scratch, enough for 16 warps
constexpr int VAR_THREADS = 128;
#ifndef FEC_VAR_BLOCKS_PER_SM
#define FEC_VAR_BLOCKS_PER_SM 5
#endif
constexpr int VAR_BLOCKS_PER_SM = FEC_VAR_BLOCKS_PER_SM;
constexpr int VAR_TAB_BYTES = 8 * 16 * VAR_THREADS * 4;  // 64 KiB per block
constexpr int VAR_IO_BYTES_PER_WARP = 65 * 32;           // per-warp staging of 33/65-byte records
constexpr int VAR_SMEM_BYTES = (VAR_THREADS / 32) * VAR_IO_BYTES_PER_WARP;
constexpr int FIX_THREADS = 128;
#ifndef FEC_FIX_BLOCKS_PER_SM
#define FEC_FIX_BLOCKS_PER_SM 3
#endif
constexpr int FIX_BLOCKS_PER_SM = FEC_FIX_BLOCKS_PER_SM;
constexpr int FIX_IO_BYTES_PER_WARP = 65 * 32;  // per-warp staging of 33/65-byte records
constexpr int FIX_SMEM_BYTES = (FIX_THREADS / 32) * FIX_IO_BYTES_PER_WARP + INV_SCRATCH_WORDS * 4;  // + block-level inversion scratch
constexpr int X_THREADS = 128;
#ifndef FEC_X_BLOCKS_PER_SM
#define FEC_X_BLOCKS_PER_SM 4
#endif
constexpr int X_BLOCKS_PER_SM = FEC_X_BLOCKS_PER_SM;
constexpr int ED_THREADS = 128;
#ifndef FEC_ED_BLOCKS_PER_SM
#define FEC_ED_BLOCKS_PER_SM 3
#endif
constexpr int ED_BLOCKS_PER_SM = FEC_ED_BLOCKS_PER_SM;
constexpr int FB_THREADS = 256;
constexpr int PEAK_THREADS = 256;
constexpr int PEAK_CHAINS = 8;
constexpr int PEAK_INNER = 16;

// grid of a persistent kernel: one block per `tiles` unit of work until the device is full (sms x resident blocks).  The
// deferred-inversion kernels pass the number of 128-lane tiles, so a small batch spreads over all SMs (one tile per warp,
// one inversion each) and only a large one makes every warp walk several tiles and share an inversion among them.
// Small-batch mode of the deferred-inversion kernels: when no warp of the grid gets more than `defer` / 2 tiles, every warp
// inverts exactly once, after its last tile, and the warps of a block then share ONE inversion (block_batch_inv) instead
// of running one each.  (With more tiles per warp the warps stay independent: a barrier per batch would make three warps
// wait for the fourth's Fermat chain.)
// Below two blocks per SM (148 SMs, 4 warps per block) the device is not full and a batch is latency-bound: the barrier
// only adds latency there (measured: -3 % at 2^14 lanes), so those grids keep the warp-level inversion too.
__host__ __device__ inline bool small_batch_mode(size_t ntiles, size_t nwarps, int defer) {
  return 2 * ntiles <= nwarps * (size_t)defer && nwarps >= 2 * 148 * 4;  // up to defer / 2 tiles per warp (beyond: -1 %)
}
inline int persistent_grid(int sms, int blocks_per_sm, size_t tiles) {
  size_t g = (size_t)sms * blocks_per_sm;
  if (tiles < g) g = tiles;
  return g < 1 ? 1 : (int)g;
}

// ---- launchers (defined in the k_*.cu files); each returns cudaGetLastError() ----------------
cudaError_t wmul_var_setup();
// optional per-stage timing of the three-kernel k*P path: 4 events per chunk (start, after prepare, after ladder, after finish)
struct StageEvents {
  cudaEvent_t* ev;  // capacity `cap` events, `used` of them recorded
  int cap, used;
};
cudaError_t wmul_var_launch(int curve, int sms, cudaStream_t st, const uint8_t* scalars, const uint8_t* points, int points_fmt,
                            size_t n, int out_fmt, uint8_t* out, uint8_t* status, int* launches, StageEvents* se);
// secp256k1 k*P as prepare / ladder / finish kernels over a stream-ordered scratch allocation (k_wmul_glv_k1.cu)
// lanes per scratch allocation / kernel launch of the three-kernel paths (2^22 lanes = 3.9 GB of scratch for secp256k1);
// FORGE_EC_SCRATCH_LG=18..24 overrides it (measurements)
inline size_t glv_chunk_lanes() {
  static const size_t v = [] {
    const char* x = getenv("FORGE_EC_SCRATCH_LANES");  // exact lane count (measurements)
    if (x != nullptr && atoll(x) >= 1024) return (size_t)atoll(x) / 1024 * 1024;
    const char* e = getenv("FORGE_EC_SCRATCH_LG");
    const int lg = e != nullptr ? atoi(e) : 0;
    return size_t(1) << ((lg >= 18 && lg <= 24) ? lg : 22);
  }();
  return v;
}
#define GLV_CHUNK_LANES (glv_chunk_lanes())
// stream-ordered scratch from the library's own per-device pool (k_misc.cu); freed with cudaFreeAsync
cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st);
void scratch_trim(int dev);
cudaError_t wmul_glv_setup_k1();
cudaError_t wmul_glv_launch_k1(cudaStream_t st, const uint8_t* scalars, const uint8_t* points, int points_fmt, size_t n, int out_fmt,
                               uint8_t* out, uint8_t* status, size_t chunk, int* launches, StageEvents* se);
// P-256 k*P, the same three-kernel shape without an endomorphism (k_wmul_vm_p256.cu)
cudaError_t wmul_vm_setup_p256();
cudaError_t wmul_vm_launch_p256(cudaStream_t st, const uint8_t* scalars, const uint8_t* points, int points_fmt, size_t n, int out_fmt,
                                uint8_t* out, uint8_t* status, size_t chunk, int* launches, StageEvents* se);
// ECDH tail: n x 32-byte big-endian x-coordinates -> n x out_len bytes (kdf.cuh); dead lanes (status != 0) give zero records
cudaError_t kdf_launch(cudaStream_t st, int kdf, const uint8_t* x_be, const uint8_t* status, size_t n, const uint8_t* info, int info_len,
                       int out_len, uint8_t* out);
// multi-scalar multiplication tail: sum of n affine products (X || Y, 64 B) with their lane status -> one encoded point
int msm_blocks(int sms, size_t n);
cudaError_t msm_reduce_launch(int curve, int sms, cudaStream_t st, const uint8_t* xy, const uint8_t* lane_status, size_t n, u32* scratch,
                              int out_fmt, uint8_t* out, uint8_t* status);
// bucket method for n >= MSM_PIPPENGER_MIN_* points (k_msm_pippenger.cu); *skewed = 1 (and nothing else done) when some bucket holds
// more than skew_limit points, in which case the caller uses the products + reduction path
constexpr size_t MSM_PIPPENGER_MIN_K1 = size_t(1) << 17, MSM_PIPPENGER_MIN_P256 = size_t(1) << 16;  // measured crossovers
constexpr size_t MSM_PIPPENGER_MAX = 0xFFFFFFFFull / 17;  // 17 n (window, lane) entries are counted and addressed in 32 bits
size_t msm_pippenger_scratch_bytes(size_t n);
cudaError_t msm_pippenger_launch(int curve, cudaStream_t st, const uint8_t* scalars, const uint8_t* points, int points_fmt, size_t n,
                                 uint8_t* scratch, int out_fmt, uint8_t* out, uint8_t* status, unsigned skew_limit, int* skewed);
// signing-side nonce commitment tails (k_nonce.cu): r = x(R) / the even-y rule behind R = k G; 64-byte value mod L in front of [r]B
cudaError_t nonce_tail_launch(int curve, int mode, cudaStream_t st, const uint8_t* xy, const uint8_t* k_be, const uint8_t* st_in, size_t n,
                              uint8_t* out_R, uint8_t* out_r, uint8_t* status);
cudaError_t sc_reduce64_launch(cudaStream_t st, const uint8_t* in64, size_t n, uint8_t* out_le);
// standalone SEC1 decoding + curve membership (k_nonce.cu)
cudaError_t point_decode_launch(int curve, cudaStream_t st, const uint8_t* points, int points_fmt, size_t n, int out_fmt, uint8_t* out,
                                uint8_t* status);
cudaError_t quirk_k1_launch(cudaStream_t st, const uint8_t* scalars, const uint8_t* points, size_t n, int out_fmt, uint8_t* out,
                            uint8_t* status);
cudaError_t quirk_ed25519_launch(cudaStream_t st, const uint8_t* seeds, size_t n, uint8_t* out);
cudaError_t quirk_x25519_launch(cudaStream_t st, const uint8_t* scalars, const uint8_t* us, size_t n, uint8_t* out);
cudaError_t quirk_p256_launch(cudaStream_t st, const uint8_t* scalars, const uint8_t* points, size_t n, int out_fmt, uint8_t* out,
                              uint8_t* status);
cudaError_t wmul_fixed_setup();
cudaError_t wmul_fixed_launch(int curve, int sms, cudaStream_t st, const uint8_t* scalars, const u32* comb, size_t n,
                              int out_fmt, uint8_t* out, uint8_t* status);
// constant-time table form of k*G (FORGE_EC_CT_TABLE=1): a 64 x 8 comb of odd multiples (32 KiB per curve) in shared memory,
// every addend read by a masked scan of its window (lanes.cuh: ScanComb)
bool fixed_ct_table();
cudaError_t build_comb_ct_launch(int curve, cudaStream_t st, u32* comb4 /* CT_TABLE_WORDS words */);
cudaError_t wmul_fixed_ct_launch(int curve, int sms, cudaStream_t st, const uint8_t* scalars, const u32* comb4, size_t n, int out_fmt,
                                 uint8_t* out, uint8_t* status);
cudaError_t x25519_launch(int sms, cudaStream_t st, const uint8_t* scalars, const uint8_t* us, size_t n, uint8_t* out);
// the same ladder on the accumulator machine (k_x25519_vm.cu); FORGE_EC_X_IMPL=regs|vm forces one form (measurements)
cudaError_t x25519_vm_setup();
cudaError_t x25519_vm_launch(int sms, cudaStream_t st, const uint8_t* scalars, const uint8_t* us, size_t n, uint8_t* out);
cudaError_t ed25519_setup();
cudaError_t ed25519_launch(int sms, cudaStream_t st, const uint8_t* in, int mode /* 0 scalars, 1 seeds, 2 X25519 base point */, const u32* niels, size_t n,
                           uint8_t* out);
cudaError_t build_comb_launch(int curve, cudaStream_t st, u32* comb, u32* bases /* COMB_WINDOWS x 16 words scratch */);
cudaError_t build_niels_launch(cudaStream_t st, u32* niels, u32* bases /* COMB_WINDOWS x 32 words scratch */);
// constant-time table form of the Ed25519 / X25519 base-point multiplication (FORGE_EC_CT_TABLE=1): 64 x 8 Niels triples (48 KiB)
cudaError_t build_niels_ct_launch(cudaStream_t st, u32* niels4 /* CT_NIELS_WORDS words */);
cudaError_t ed25519_ct_launch(int sms, cudaStream_t st, const uint8_t* in, int mode, const u32* niels4, size_t n, uint8_t* out);
cudaError_t field_bench_launch(int sms, cudaStream_t st, int prime, int op, const uint8_t* a, const uint8_t* b, uint8_t* out,
                               size_t lanes, int iters);
cudaError_t imad_peak_launch(int variant, int blocks, cudaStream_t st, u32* sink, u32 seed, int trips);
cudaError_t synth_launch(int sms, cudaStream_t st, uint8_t* out, size_t n, u32 seed_lo, u32 seed_hi, u32 stream_id, u32 sub,
                         unsigned long long first_lane, int mode, int edge);

}  // namespace fec
