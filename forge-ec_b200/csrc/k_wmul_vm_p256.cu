// k_wmul_vm_p256.cu -- P-256 variable-base k*P as three kernels over a per-call scratch area in HBM (see p256v3.cuh /
// vm.cuh for the why; the secp256k1 counterpart is k_wmul_glv_k1.cu):
//
//   k_p3_prep    one lane per thread, registers: parse + validate (k, P), table of 8 affine odd multiples (one
//                warp-level inversion per tile), recoding     -> scratch: table 512 B, digits 32 B, flags 4 B per lane
//   k_p3_ladder  accumulator machine in shared memory, 9 slots of 32 B per lane, table entries fetched with cp.async
//                straight into the addend slots              -> scratch: Jacobian result 96 B per lane
//   k_p3_finish  sign, deferred warp-level inversion, affine, SEC1 / X-only encoding, status
#define FEC_NOINLINE_FIELD 1
#include "p256v3.cuh"
#include "kernels.cuh"

namespace fec {

#ifndef FEC_P3_PREP_THREADS
#define FEC_P3_PREP_THREADS 128  // (r2: 128 threads 32.97, 256 threads 32.84, 512 threads 32.86 M/s; one inversion per warp 31.84)
#endif
constexpr int P3_PREP_THREADS = FEC_P3_PREP_THREADS;
#ifndef FEC_P3_PREP_BLOCKS
#define FEC_P3_PREP_BLOCKS (512 / FEC_P3_PREP_THREADS)  // 128 registers
#endif
#ifndef FEC_P3_PREP_BLOCK_INV
#define FEC_P3_PREP_BLOCK_INV 1  // the warps of a block share ONE inversion (0: one per warp, the form before r2)
#endif
#ifndef FEC_P3_LAD_THREADS
#define FEC_P3_LAD_THREADS 128
#endif
constexpr int P3_LAD_THREADS = FEC_P3_LAD_THREADS;
#ifndef FEC_P3_LAD_BLOCKS
#define FEC_P3_LAD_BLOCKS 6  // 80 registers: 31.5 -> 31.8 M/s once the multiplication lost its predicate spills (4: 30.9)
#endif
constexpr int P3_LAD_SMEM = VS_COUNT * 32 * P3_LAD_THREADS;
constexpr int P3_FIN_THREADS = 128;
#ifndef FEC_P3_FIN_BLOCKS
#define FEC_P3_FIN_BLOCKS 5
#endif
constexpr int P3_FIN_BLOCKS = FEC_P3_FIN_BLOCKS;
#ifndef FEC_P3_FIN_DEFER
#define FEC_P3_FIN_DEFER 48
#endif
constexpr int P3_FIN_DEFER = FEC_P3_FIN_DEFER;
constexpr int P3_FIN_SMEM = (P3_FIN_THREADS / 32) * 65 * 32;

struct P3Scratch {
  u32* tab;    // [lane][128], plus 32 spare rows for the lanes past the end of the last tile
  u32* q;      // [lane][24]
  u32* kk;     // [lane][8]
  u32* flags;  // [lane]   bit 0: negate the result, bits 8-15 status
};
constexpr size_t P3_SCRATCH_BYTES_PER_LANE = (P3_TAB_WORDS + 24 + 8 + 1) * 4;

struct P3WarpInv {
  __device__ __forceinline__ void operator()(u32 zi[8], const u32 z[8]) { warp_batch_inv<FeP256>(zi, z); }
};

__global__ void __launch_bounds__(P3_PREP_THREADS, FEC_P3_PREP_BLOCKS)
k_p3_prep(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ points, int points_fmt, size_t n, P3Scratch sc) {
  typedef CurveP256 C;
  typedef FeP256 F;
  __shared__ __align__(16) uint8_t io_all[(P3_PREP_THREADS / 32) * 33 * 32 + 16];
  const unsigned lane_in_warp = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint8_t* io = io_all + warp * (33 * 32);
  const size_t first = ((size_t)blockIdx.x * (P3_PREP_THREADS / 32) + warp) * 32;
#if FEC_P3_PREP_BLOCK_INV
  // The inversion of the table's common denominator is the expensive part of this kernel (255 squarings: as much as the
  // table itself) and a warp pays for it whether one lane or 32 need it.  The warps of a block therefore share ONE inversion
  // (block_batch_inv: products across the warps through shared memory, warp 0 inverts, everybody multiplies back), and a
  // warp past the end of the batch stays for the barriers and recomputes the last lane into the spare rows.
  __shared__ u32 inv_scratch[INV_SCRATCH_WORDS];
  const int cnt_lanes = first >= n ? 0 : (int)((n - first) < 32 ? (n - first) : 32);
#else
  if (first >= n) return;  // warp-uniform
  const int cnt_lanes = (int)((n - first) < 32 ? (n - first) : 32);
#endif
  const bool live = (int)lane_in_warp < cnt_lanes;
  // lanes past the end recompute the last lane (the whole warp takes part in the inversion) into 32 spare table rows
  const size_t lane = live ? first + lane_in_warp : n - 1;
  const size_t row = live ? lane : n + lane_in_warp;
  u32 k[8], px[8], py[8], ok, idn = 0u;
  if (points_fmt == FEC_PT_SEC1_COMPRESSED) {
    // (a lane past the end reads the record of the batch's last lane, which lives in another warp's staging area or nowhere:
    // it takes the record from global memory instead)
    warp_copy(io, points + first * 33, (size_t)cnt_lanes * 33);
    __syncwarp();
    if (live) {
      ok = sec1_decompress_id<C>(px, py, io + lane_in_warp * 33, idn);
    } else {
      uint8_t rec[33];
      for (int i = 0; i < 33; i++) rec[i] = points[lane * 33 + i];
      ok = sec1_decompress_id<C>(px, py, rec, idn);
    }
  } else {
    load_rec32_be(px, points + lane * 64);
    load_rec32_be(py, points + lane * 64 + 32);
    ok = affine_from_plain_id<C>(px, py, idn);
  }
  load_rec32_be(k, scalars + lane * 32);
  {
    u32 gx[8], gy[8];
    C::G(gx, gy);
    select8(px, ok, px, gx);
    select8(py, ok, py, gy);
  }
  u32 K[8], negm, st;
  scalar_prepare<C>(K, negm, st, k);
  st = lane_status(st, ok, idn);  // range > invalid point > identity (k = 0, or the identity as the input point)
  GlobalLaneTab tab{sc.tab + row * P3_TAB_WORDS};
#if FEC_P3_PREP_BLOCK_INV
  BlockInv<F, P3_PREP_THREADS / 32> inv{inv_scratch};
#else
  P3WarpInv inv;
#endif
  wmul_var_table<C>(tab, px, py, inv);
  if (live) {
    g_st8(sc.kk + lane * 8, K);
    sc.flags[lane] = (negm & 1u) | (st << 8);
  }
}

// r2: the same with the inversion DEFERRED over P3_PREP_DEFER consecutive tiles of a warp (as the finish kernels do): the first
// half of every tile's table (co-Z chain, un-normalised entries in the scratch area) runs first, the common denominators are
// multiplied together, ONE warp-level inversion follows, and the tables are normalised on the way back (Montgomery's trick along
// the tile axis: 2 multiplications per tile).  The Fermat chain costs 270 field operations per warp whatever it inverts: a share of
// 67 per tile with one inversion per block of four tiles (three of whose warps wait at the barrier meanwhile), 34 + 2 at 8 tiles.
#ifndef FEC_P3_PREP_DEFER
#define FEC_P3_PREP_DEFER 16  // most tiles per inversion (0: always k_p3_prep above); the launcher takes fewer when the batch is small
#endif
constexpr int P3_PREP_DEFER = FEC_P3_PREP_DEFER > 0 ? FEC_P3_PREP_DEFER : 1;
__global__ void __launch_bounds__(P3_PREP_THREADS, FEC_P3_PREP_BLOCKS)
k_p3_prep_defer(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ points, int points_fmt, size_t n, P3Scratch sc,
                int tiles_per_warp /* 1 .. P3_PREP_DEFER */) {
  typedef CurveP256 C;
  typedef FeP256 F;
  __shared__ __align__(16) uint8_t io_all[(P3_PREP_THREADS / 32) * 33 * 32 + 16];
  const unsigned lane_in_warp = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint8_t* io = io_all + warp * (33 * 32);
  const size_t ntiles = (n + 31) / 32;
  const size_t tile0 = ((size_t)blockIdx.x * (P3_PREP_THREADS / 32) + warp) * (size_t)tiles_per_warp;
  if (tile0 >= ntiles) return;  // warp-uniform; no block-level barrier in this kernel
  u32 zr[P3_PREP_DEFER][7][8];  // thread-local memory
  u32 zts[P3_PREP_DEFER][8], pre[P3_PREP_DEFER][8];
  u32 acc[8];
  F::one(acc);
  int cnt = 0;
#pragma unroll 1
  for (int t = 0; t < tiles_per_warp; t++) {
    const size_t first = (tile0 + t) * 32;
    if (first >= n) break;  // warp-uniform
    const int cnt_lanes = (int)((n - first) < 32 ? (n - first) : 32);
    const bool live = (int)lane_in_warp < cnt_lanes;
    // lanes past the end recompute the last lane (the whole warp takes part in the inversion) into 32 spare table rows
    const size_t lane = live ? first + lane_in_warp : n - 1;
    const size_t row = live ? lane : n + lane_in_warp;
    u32 k[8], px[8], py[8], ok, idn = 0u;
    if (points_fmt == FEC_PT_SEC1_COMPRESSED) {
      __syncwarp();
      warp_copy(io, points + first * 33, (size_t)cnt_lanes * 33);
      __syncwarp();
      if (live) {
        ok = sec1_decompress_id<C>(px, py, io + lane_in_warp * 33, idn);
      } else {
        uint8_t rec[33];
        for (int i = 0; i < 33; i++) rec[i] = points[lane * 33 + i];
        ok = sec1_decompress_id<C>(px, py, rec, idn);
      }
    } else {
      load_rec32_be(px, points + lane * 64);
      load_rec32_be(py, points + lane * 64 + 32);
      ok = affine_from_plain_id<C>(px, py, idn);
    }
    load_rec32_be(k, scalars + lane * 32);
    {
      u32 gx[8], gy[8];
      C::G(gx, gy);
      select8(px, ok, px, gx);
      select8(py, ok, py, gy);
    }
    u32 K[8], negm, st;
    scalar_prepare<C>(K, negm, st, k);
    st = lane_status(st, ok, idn);  // range > invalid point > identity (k = 0, or the identity as the input point)
    GlobalLaneTab tab{sc.tab + row * P3_TAB_WORDS};
    u32 zt[8];
    wmul_var_table_build<C>(tab, px, py, zt, zr[t]);
    if (live) {
      g_st8(sc.kk + lane * 8, K);
      sc.flags[lane] = (negm & 1u) | (st << 8);
    }
    copy8(zts[t], zt);
    copy8(pre[t], acc);
    if (t == 0) copy8(acc, zt); else F::mul(acc, acc, zt);
    cnt++;
  }
  u32 inv[8];
  warp_batch_inv<F>(inv, acc);
#pragma unroll 1
  for (int t = cnt - 1; t >= 0; t--) {
    u32 zi[8];
    if (t > 0) { F::mul(zi, inv, pre[t]); F::mul(inv, inv, zts[t]); } else { copy8(zi, inv); }
    const size_t first = (tile0 + t) * 32;
    const int cnt_lanes = (int)((n - first) < 32 ? (n - first) : 32);
    const bool live = (int)lane_in_warp < cnt_lanes;
    const size_t row = live ? first + lane_in_warp : n + lane_in_warp;
    GlobalLaneTab tab{sc.tab + row * P3_TAB_WORDS};
    wmul_var_table_finish<C>(tab, zi, zr[t]);
  }
}

#ifdef FEC_P3_LAD_MAXNREG  // A/B: an explicit register cap instead of a resident-block count
__global__ void __maxnreg__(FEC_P3_LAD_MAXNREG)
#else
__global__ void __launch_bounds__(P3_LAD_THREADS, FEC_P3_LAD_BLOCKS)
#endif
k_p3_ladder(size_t n, P3Scratch sc) {
  extern __shared__ __align__(16) u32 smem[];
  size_t lane = (size_t)blockIdx.x * P3_LAD_THREADS + threadIdx.x;
  const bool live = lane < n;
  if (!live) lane = n - 1;  // recompute the last lane instead of branching around the whole loop
  const saddr S = (saddr)__cvta_generic_to_shared(smem) + threadIdx.x * 16u;
  u32 K[8];
  g_ld8(K, sc.kk + lane * 8);
  p3_ladder<P3_LAD_THREADS>(S, sc.tab + lane * P3_TAB_WORDS, K);
  if (live) {
    typedef Vm<FeP256, P3_LAD_THREADS> V;
    u32 t[8];
    V::ldv(t, V::slot(S, VS_X)); g_st8(sc.q + lane * 24, t);
    V::ldv(t, V::slot(S, VS_Y)); g_st8(sc.q + lane * 24 + 8, t);
    V::ldv(t, V::slot(S, VS_Z)); g_st8(sc.q + lane * 24 + 16, t);
  }
}

// Warp-independent: every warp walks tiles of 32 lanes, keeps the Jacobian results of P3_FIN_DEFER tiles in thread-local
// memory with the running product of their Z and then inverts once per warp (Montgomery's trick along the tile axis).
__global__ void __launch_bounds__(P3_FIN_THREADS, P3_FIN_BLOCKS)
k_p3_finish(size_t n, P3Scratch sc, int out_fmt, uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  typedef CurveP256 C;
  typedef FeP256 F;
  extern __shared__ __align__(16) u32 smem[];
  const unsigned lane_in_warp = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint8_t* io = reinterpret_cast<uint8_t*>(smem) + warp * (65 * 32);
  const int osz = out_size(out_fmt);
  const size_t ntiles = (n + 31) / 32;
  const size_t gwarp = (size_t)blockIdx.x * (P3_FIN_THREADS / 32) + warp, nwarps = (size_t)gridDim.x * (P3_FIN_THREADS / 32);
  u32 slots[P3_FIN_DEFER][36];  // X, Y, Z, prefix product, status
  u32 acc[8];
  int cnt = 0;
  size_t first_tile_of_batch = 0;
  F::one(acc);
  for (size_t tile = gwarp;; tile += nwarps) {
    const bool have = tile < ntiles;  // warp-uniform
    if (have) {
      size_t lane = tile * 32 + lane_in_warp;
      const bool live = lane < n;
      if (!live) lane = n - 1;
      Jac Q;
      g_ld8(Q.X, sc.q + lane * 24); g_ld8(Q.Y, sc.q + lane * 24 + 8); g_ld8(Q.Z, sc.q + lane * 24 + 16);
      const u32 fl = sc.flags[lane];
      u32 st = (fl >> 8) & 0xFFu;
      u32 ny[8];
      F::neg(ny, Q.Y);
      select8(Q.Y, 0u - (fl & 1u), ny, Q.Y);
      const u32 zz = F::is_zero(Q.Z);
      st |= (st == 0u ? zz & LANE_INFINITY : 0u);
      if (status != nullptr && live) status[lane] = (uint8_t)st;
      u32 one[8], z[8];
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
    if (cnt == P3_FIN_DEFER || (!have && cnt > 0)) {
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

static int g_sms_p3 = 0;
cudaError_t wmul_vm_setup_p256() {
  cudaError_t e = cudaFuncSetAttribute(k_p3_ladder, cudaFuncAttributeMaxDynamicSharedMemorySize, P3_LAD_SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_p3_ladder, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&g_sms_p3, cudaDevAttrMultiProcessorCount, dev);
  return cudaFuncSetAttribute(k_p3_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, P3_FIN_SMEM);
}

static unsigned p3_fin_grid(size_t m) {
  const size_t want = (m + (size_t)P3_FIN_THREADS - 1) / (size_t)P3_FIN_THREADS;  // one tile per warp until the device is full
  const size_t cap = (size_t)(g_sms_p3 > 0 ? g_sms_p3 : 148) * P3_FIN_BLOCKS;
  const size_t g = want < cap ? want : cap;
  return (unsigned)(g < 1 ? 1 : g);
}
// n lanes in chunks of at most `chunk` lanes over one stream-ordered scratch allocation
cudaError_t wmul_vm_launch_p256(cudaStream_t st, const uint8_t* scalars, const uint8_t* points, int points_fmt, size_t n, int out_fmt,
                                uint8_t* out, uint8_t* status, size_t chunk, int* launches, StageEvents* se) {
  if (n == 0) return cudaSuccess;
  const size_t cap = n < chunk ? n : chunk;
  uint8_t* base = nullptr;
  cudaError_t e = scratch_alloc(reinterpret_cast<void**>(&base), cap * P3_SCRATCH_BYTES_PER_LANE + 32 * P3_TAB_WORDS * 4 + 256, st);
  if (e != cudaSuccess) return e;
  P3Scratch sc;
  sc.tab = reinterpret_cast<u32*>(base);
  sc.q = sc.tab + (cap + 32) * P3_TAB_WORDS;
  sc.kk = sc.q + cap * 24;
  sc.flags = sc.kk + cap * 8;
  const int osz = out_size(out_fmt);
  const int psz = points_fmt == FEC_PT_SEC1_COMPRESSED ? 33 : 64;
  auto mark = [&]() {
    if (se != nullptr && se->used < se->cap) cudaEventRecord(se->ev[se->used++], st);
  };
  for (size_t off = 0; off < n; off += cap) {
    const size_t m = (n - off) < cap ? (n - off) : cap;
    mark();
    // tiles per inversion: as many as leave every resident warp of the device something to do; below two tiles per warp the
    // block-level inversion of k_p3_prep is the better deal (four tiles per inversion, all of them in flight at once)
    const size_t warps_resident = (size_t)(g_sms_p3 > 0 ? g_sms_p3 : 148) * FEC_P3_PREP_BLOCKS * (P3_PREP_THREADS / 32);
    const size_t tiles = (m + 31) / 32;
    size_t tpw = FEC_P3_PREP_DEFER > 0 ? tiles / warps_resident : 0;
    if (tpw > (size_t)P3_PREP_DEFER) tpw = P3_PREP_DEFER;
    if (tpw >= 2) {
      const size_t lanes_per_block = (size_t)P3_PREP_THREADS * tpw;
      k_p3_prep_defer<<<(unsigned)((m + lanes_per_block - 1) / lanes_per_block), P3_PREP_THREADS, 0, st>>>(
          scalars + off * 32, points + off * psz, points_fmt, m, sc, (int)tpw);
    } else {
      k_p3_prep<<<(unsigned)((m + P3_PREP_THREADS - 1) / P3_PREP_THREADS), P3_PREP_THREADS, 0, st>>>(
          scalars + off * 32, points + off * psz, points_fmt, m, sc);
    }
    mark();
    k_p3_ladder<<<(unsigned)((m + P3_LAD_THREADS - 1) / P3_LAD_THREADS), P3_LAD_THREADS, P3_LAD_SMEM, st>>>(m, sc);
    mark();
    k_p3_finish<<<p3_fin_grid(m), P3_FIN_THREADS, P3_FIN_SMEM, st>>>(m, sc, out_fmt, out + off * osz, status ? status + off : nullptr);
    mark();
    if (launches) *launches += 3;
  }
  e = cudaGetLastError();
  cudaError_t e2 = cudaFreeAsync(base, st);
  return e != cudaSuccess ? e : e2;
}

}  // namespace fec
