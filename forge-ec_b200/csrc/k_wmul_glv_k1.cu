// k_wmul_glv_k1.cu -- secp256k1 variable-base k*P (BASELINE.json configs[2], the headline workload) as three
// kernels over a per-call scratch area in HBM (see glv3.cuh / vm.cuh for the why):
//
//   k_glv_prep    one lane per thread, registers: parse + validate (k, P), table of 8 odd multiples as
//                 (x, beta x, y), GLV split           -> scratch: table 768 B, zs 32 B, digits 32 B, flags 4 B per lane
//   k_glv_ladder  accumulator machine in shared memory, 9 slots of 32 B per lane, table entries fetched with
//                 cp.async straight into the addend slots   -> scratch: Jacobian result 96 B per lane
//   k_glv_finish  parity corrections, block-level simultaneous inversion, affine, SEC1 encoding, status
//
// Scratch traffic per lane: 932 B written + ~5 KB read (64 table entries of 64 B), i.e. ~0.4 TB/s at 7e7 lanes/s:
// 6 % of HBM bandwidth, and the table of a chunk mostly stays in the 126 MB L2.
#define FEC_NOINLINE_FIELD 1
#include "glv3.cuh"
#include "kernels.cuh"

namespace fec {

constexpr int GLV_PREP_THREADS = 128;
#ifndef FEC_GLV_LAD_THREADS
#define FEC_GLV_LAD_THREADS 128
#endif
constexpr int GLV_LAD_THREADS = FEC_GLV_LAD_THREADS;
#ifndef FEC_GLV_LAD_BLOCKS
#define FEC_GLV_LAD_BLOCKS 5
#endif
constexpr int GLV_LAD_SMEM = VS_COUNT * 32 * GLV_LAD_THREADS;
constexpr int GLV_FIN_THREADS = 128;
#ifndef FEC_GLV_FIN_BLOCKS
#define FEC_GLV_FIN_BLOCKS 5
#endif
constexpr int GLV_FIN_BLOCKS = FEC_GLV_FIN_BLOCKS;
#ifndef FEC_GLV_FIN_DEFER
#define FEC_GLV_FIN_DEFER 48
#endif
constexpr int GLV_FIN_DEFER = FEC_GLV_FIN_DEFER;
constexpr int GLV_FIN_SMEM = (GLV_FIN_THREADS / 32) * 65 * 32;

struct GlvScratch {
  u32* tab;    // [lane][192]
  u32* zs;     // [lane][8]
  u32* kk;     // [lane][8]
  u32* flags;  // [lane]   bits 0-3 GLV flags, bits 8-15 status
  u32* q;      // [lane][24]
};
constexpr size_t GLV_SCRATCH_BYTES_PER_LANE = (192 + 8 + 8 + 1 + 24) * 4;

#ifndef FEC_GLV_PREP_BLOCKS
#define FEC_GLV_PREP_BLOCKS 4
#endif
__global__ void __launch_bounds__(GLV_PREP_THREADS, FEC_GLV_PREP_BLOCKS)
k_glv_prep(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ points, int points_fmt, size_t n, GlvScratch sc) {
  typedef CurveK1 C;
  typedef FeK1 F;
  __shared__ __align__(16) uint8_t io_all[(GLV_PREP_THREADS / 32) * 33 * 32 + 16];
  const unsigned lane_in_warp = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint8_t* io = io_all + warp * (33 * 32);
  const size_t first = ((size_t)blockIdx.x * (GLV_PREP_THREADS / 32) + warp) * 32;
  if (first >= n) return;  // warp-uniform
  const size_t lane = first + lane_in_warp;
  const int cnt_lanes = (int)((n - first) < 32 ? (n - first) : 32);
  const bool live = (int)lane_in_warp < cnt_lanes;
  u32 k[8], px[8], py[8], ok = 0xFFFFFFFFu, idn = 0u;
  set_small(k, 1u);
  C::G(px, py);
  if (points_fmt == FEC_PT_SEC1_COMPRESSED) {
    warp_copy(io, points + first * 33, (size_t)cnt_lanes * 33);
    __syncwarp();
    if (live) ok = sec1_decompress_id<C>(px, py, io + lane_in_warp * 33, idn);
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
  u32 nn[8], one[8];
  C::N(nn);
  set_small(one, 1u);
  const u32 bad = geq8(k, nn), zero = is_zero8(k);
  select8(k, bad | zero, one, k);
  u32 st = (bad & LANE_SCALAR_RANGE) | (~bad & zero & LANE_INFINITY);
  st = lane_status(st, ok, idn);  // range > invalid point > identity (k = 0, or the identity as the input point)
  if (!live) return;
  u32 zs[8], KK[8], flags;
  glv_prepare(sc.tab + lane * GLV_TAB_WORDS, zs, KK, flags, k, px, py);
  g_st8(sc.zs + lane * 8, zs);
  g_st8(sc.kk + lane * 8, KK);
  sc.flags[lane] = flags | (st << 8);
}

#ifdef FEC_GLV_LAD_MAXNREG  // A/B: an explicit register cap instead of a resident-block count
__global__ void __maxnreg__(FEC_GLV_LAD_MAXNREG)
#else
__global__ void __launch_bounds__(GLV_LAD_THREADS, FEC_GLV_LAD_BLOCKS)
#endif
k_glv_ladder(size_t n, GlvScratch sc) {
  extern __shared__ __align__(16) u32 smem[];
  size_t lane = (size_t)blockIdx.x * GLV_LAD_THREADS + threadIdx.x;
  const bool live = lane < n;
  if (!live) lane = n - 1;  // recompute the last lane instead of branching around the whole loop
  const saddr S = (saddr)__cvta_generic_to_shared(smem) + threadIdx.x * 16u;
  u32 KK[8];
  g_ld8(KK, sc.kk + lane * 8);
  const u32 flags = sc.flags[lane];
  glv_ladder<GLV_LAD_THREADS>(S, sc.tab + lane * GLV_TAB_WORDS, KK, flags);
  if (live) {
    typedef Vm<FeK1, GLV_LAD_THREADS> V;
    u32 t[8];
    V::ldv(t, V::slot(S, VS_X)); g_st8(sc.q + lane * 24, t);
    V::ldv(t, V::slot(S, VS_Y)); g_st8(sc.q + lane * 24 + 8, t);
    V::ldv(t, V::slot(S, VS_Z)); g_st8(sc.q + lane * 24 + 16, t);
  }
}

// A/B form of the ladder (FORGE_EC_GLV_LADDER=regs): the point in registers, field operations in line (glv_ladder_regs)
#ifndef FEC_GLV_LADR_BLOCKS
#define FEC_GLV_LADR_BLOCKS 3
#endif
#ifndef FEC_GLV_LADR_INLINE
#define FEC_GLV_LADR_INLINE 2  // 2: multiplication and squaring in line, 1: squaring only, 0: both called
#endif
__global__ void __launch_bounds__(GLV_LAD_THREADS, FEC_GLV_LADR_BLOCKS)
k_glv_ladder_regs(size_t n, GlvScratch sc) {
  size_t lane = (size_t)blockIdx.x * GLV_LAD_THREADS + threadIdx.x;
  const bool live = lane < n;
  if (!live) lane = n - 1;
  u32 KK[8];
  g_ld8(KK, sc.kk + lane * 8);
  const u32 flags = sc.flags[lane];
  Jac Q;
#if FEC_GLV_LADR_INLINE == 2
  glv_ladder_regs<CurveK1Inl>(Q, sc.tab + lane * GLV_TAB_WORDS, KK, flags);
#elif FEC_GLV_LADR_INLINE == 1
  glv_ladder_regs<CurveK1InlSqr>(Q, sc.tab + lane * GLV_TAB_WORDS, KK, flags);
#else
  glv_ladder_regs<CurveK1>(Q, sc.tab + lane * GLV_TAB_WORDS, KK, flags);
#endif
  if (live) {
    g_st8(sc.q + lane * 24, Q.X);
    g_st8(sc.q + lane * 24 + 8, Q.Y);
    g_st8(sc.q + lane * 24 + 16, Q.Z);
  }
}
static bool ladder_regs() {  // development switch (A/B measurements only)
  static const bool v = [] { const char* e = getenv("FORGE_EC_GLV_LADDER"); return e != nullptr && e[0] == 'r'; }();
  return v;
}

// Warp-independent: every warp walks tiles of 32 lanes, applies the corrections, keeps the Jacobian results of
// GLV_FIN_DEFER tiles in thread-local memory with the running product of their Z and then inverts once per warp
// (Montgomery's trick along the tile axis; no __syncthreads, no idle warps while one lane set is being inverted).
__global__ void __launch_bounds__(GLV_FIN_THREADS, GLV_FIN_BLOCKS)
k_glv_finish(size_t n, GlvScratch sc, int out_fmt, uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  typedef CurveK1 C;
  typedef FeK1 F;
  extern __shared__ __align__(16) u32 smem[];
  const unsigned lane_in_warp = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint8_t* io = reinterpret_cast<uint8_t*>(smem) + warp * (65 * 32);
  const int osz = out_size(out_fmt);
  const size_t ntiles = (n + 31) / 32;
  const size_t gwarp = (size_t)blockIdx.x * (GLV_FIN_THREADS / 32) + warp, nwarps = (size_t)gridDim.x * (GLV_FIN_THREADS / 32);
  u32 slots[GLV_FIN_DEFER][36];  // X, Y, Z, prefix product, status
  u32 acc[8];
  int cnt = 0;
  size_t first_tile_of_batch = 0;
  set_small(acc, 1u);
  for (size_t tile = gwarp;; tile += nwarps) {
    const bool have = tile < ntiles;  // warp-uniform
    if (have) {
      size_t lane = tile * 32 + lane_in_warp;
      const bool live = lane < n;
      if (!live) lane = n - 1;
      Jac Q;
      u32 zs[8];
      g_ld8(Q.X, sc.q + lane * 24); g_ld8(Q.Y, sc.q + lane * 24 + 8); g_ld8(Q.Z, sc.q + lane * 24 + 16);
      g_ld8(zs, sc.zs + lane * 8);
      const u32 fl = sc.flags[lane];
      u32 st = (fl >> 8) & 0xFFu;
      glv_finish(Q, sc.tab + lane * GLV_TAB_WORDS, zs, fl);
      const u32 zz = F::is_zero(Q.Z);
      st |= (st == 0u ? zz & LANE_INFINITY : 0u);
      if (status != nullptr && live) status[lane] = (uint8_t)st;
      u32 one[8], z[8];
      set_small(one, 1u);
      select8(z, (st != 0u) ? 0xFFFFFFFFu : 0u, one, Q.Z);
      if (cnt == 0) first_tile_of_batch = tile;
      u32* s = slots[cnt];
#pragma unroll
      for (int i = 0; i < 8; i++) { s[i] = Q.X[i]; s[8 + i] = Q.Y[i]; s[16 + i] = z[i]; s[24 + i] = acc[i]; }
      s[32] = st;
      if (cnt == 0) copy8(acc, z); else F::mul(acc, acc, z);
      cnt++;
    }
    if (cnt == GLV_FIN_DEFER || (!have && cnt > 0)) {
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
      set_small(acc, 1u);
    }
    if (!have) break;
  }
}

static int g_sms = 0;
cudaError_t wmul_glv_setup_k1() {
  cudaError_t e = cudaFuncSetAttribute(k_glv_ladder, cudaFuncAttributeMaxDynamicSharedMemorySize, GLV_LAD_SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_glv_ladder, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
  return cudaFuncSetAttribute(k_glv_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, GLV_FIN_SMEM);
}

// persistent grid of the finish kernel: as many warps as tiles until the device is full (small batches finish in one
// tile time), then each warp walks several tiles and defers up to GLV_FIN_DEFER of them per inversion
static unsigned fin_grid(size_t m) {
  const size_t want = (m + (size_t)GLV_FIN_THREADS - 1) / (size_t)GLV_FIN_THREADS;  // one tile per warp until the device is full
  const size_t cap = (size_t)(g_sms > 0 ? g_sms : 148) * GLV_FIN_BLOCKS;
  const size_t g = want < cap ? want : cap;
  return (unsigned)(g < 1 ? 1 : g);
}
// n lanes in chunks of at most `chunk` lanes over one stream-ordered scratch allocation
cudaError_t wmul_glv_launch_k1(cudaStream_t st, const uint8_t* scalars, const uint8_t* points, int points_fmt, size_t n, int out_fmt,
                               uint8_t* out, uint8_t* status, size_t chunk, int* launches, StageEvents* se) {
  if (n == 0) return cudaSuccess;
  const size_t cap = n < chunk ? n : chunk;
  uint8_t* base = nullptr;
  cudaError_t e = scratch_alloc(reinterpret_cast<void**>(&base), cap * GLV_SCRATCH_BYTES_PER_LANE + 256, st);
  if (e != cudaSuccess) return e;
  GlvScratch sc;
  sc.tab = reinterpret_cast<u32*>(base);
  sc.q = sc.tab + cap * 192;
  sc.zs = sc.q + cap * 24;
  sc.kk = sc.zs + cap * 8;
  sc.flags = sc.kk + cap * 8;
  const int osz = out_size(out_fmt);
  const int psz = points_fmt == FEC_PT_SEC1_COMPRESSED ? 33 : 64;
  auto mark = [&]() {
    if (se != nullptr && se->used < se->cap) cudaEventRecord(se->ev[se->used++], st);
  };
  for (size_t off = 0; off < n; off += cap) {
    const size_t m = (n - off) < cap ? (n - off) : cap;
    mark();
    k_glv_prep<<<(unsigned)((m + GLV_PREP_THREADS - 1) / GLV_PREP_THREADS), GLV_PREP_THREADS, 0, st>>>(
        scalars + off * 32, points + off * psz, points_fmt, m, sc);
    mark();
    if (ladder_regs())
      k_glv_ladder_regs<<<(unsigned)((m + GLV_LAD_THREADS - 1) / GLV_LAD_THREADS), GLV_LAD_THREADS, 0, st>>>(m, sc);
    else
      k_glv_ladder<<<(unsigned)((m + GLV_LAD_THREADS - 1) / GLV_LAD_THREADS), GLV_LAD_THREADS, GLV_LAD_SMEM, st>>>(m, sc);
    mark();
    k_glv_finish<<<fin_grid(m), GLV_FIN_THREADS, GLV_FIN_SMEM, st>>>(
        m, sc, out_fmt, out + off * osz, status ? status + off : nullptr);
    mark();
    if (launches) *launches += 3;
  }
  e = cudaGetLastError();
  cudaError_t e2 = cudaFreeAsync(base, st);
  return e != cudaSuccess ? e : e2;
}

}  // namespace fec
