// k_msm_pippenger.cu -- bucket method for sum_i k_i * P_i with many points.
//
// Replaces Curve::multi_scalar_multiply (forge-ec-core/src/lib.rs:934-951, p256.rs:2193-2211: n sequential
// multiplications and additions) for n >= MSM_PIPPENGER_MIN_* (about 10^5 points, where its fixed ~2 ms of bucket
// reduction is amortised; measured 294 M points/s at n = 2^22 against 59 M/s for products + reduction).  16-bit signed windows:
//   k = sum_w d_w 2^(16 w),  d_w in [-2^15, 2^15),  17 windows (the last one only absorbs a final carry; scalars above
//   2^255 are replaced by n - k with the point negated, so that this happens with probability 2^-16 only)
//   sum_i k_i P_i = sum_w 2^(16 w) * sum_b b * B[w][b],   B[w][b] = sum of +-P_i over the lanes with |d_w| = b.
// So every point costs 17 mixed additions instead of a 256-bit scalar multiplication.  Stages (all on one stream):
//   k_pip_digits   per lane: validate scalar and point (decompress SEC1), affine limbs to scratch, digits, bucket histogram
//   (exclusive scan of the 17 x 32769 counts -- cub::DeviceScan, plumbing)
//   k_pip_scatter  counting sort: lane ids (+ sign bit) into their bucket's segment (order inside a bucket is irrelevant)
//   k_pip_buckets  one thread per bucket: gathers its points from HBM/L2 and adds them (warp-uniform trip count)
//   k_pip_segments one thread per 64 buckets: running-sum trick  sum_b b B_b  inside the segment, plus the segment offset
//   k_pip_windows  one block per window: sum of its 512 segments
//   k_pip_final    Horner over the windows (16 doublings between them), one inversion, encoding
// The bucket a lane falls into depends on the scalar (like the table index of the fixed-base kernels); control flow
// within a warp does not.
#define FEC_NOINLINE_FIELD 1
// (P-256 stays canonical in this translation unit: its kernels are dominated by general additions with masks, where the weak form's
// rare-branch bodies cost more than the shorter reduction saves -- 288 against 277 M points/s; every value crossing to other kernels is
// a byte record)
#define FEC_P256_CANONICAL
#include <cub/device/device_scan.cuh>

#include "glv3.cuh"
#include "kernels.cuh"

namespace fec {

constexpr int PIP_W = 16;
constexpr int PIP_WINDOWS = 17;
constexpr int PIP_BUCKETS = (1 << (PIP_W - 1)) + 1;  // bucket ids 0 (unused) .. 32768
constexpr int PIP_KEYS = PIP_WINDOWS * PIP_BUCKETS;
constexpr int PIP_SEG = 64;                           // buckets per segment
constexpr int PIP_SEGS = (PIP_BUCKETS - 1) / PIP_SEG; // 512 segments cover buckets 1..32768
constexpr int PIP_THREADS = 128;

struct PipScratch {
  u32* aff;      // [n][16]   affine X, Y (little-endian limbs)
  int* dig;      // [17][n]   signed digits
  u32* ent;      // [17 n]    lane | sign << 31, grouped by (window, bucket)
  u32* cnt;      // [PIP_KEYS + 1]
  u32* off;      // [PIP_KEYS + 1]
  u32* cur;      // [PIP_KEYS]
  u32* bsum;     // [PIP_KEYS][24]  bucket sums (Jacobian)
  u32* ssum;     // [17][512][24]   segment sums
  u32* wsum;     // [17][24]        window sums
  u32* bad;      // [2]  bad[0]: OR of lane rejections, bad[1]: largest bucket
};

template <class C>
__global__ void __launch_bounds__(PIP_THREADS, 3)
k_pip_digits(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ points, int points_fmt, size_t n, PipScratch sc) {
  typedef typename C::F F;
  const size_t lane = (size_t)blockIdx.x * PIP_THREADS + threadIdx.x;
  const bool live = lane < n;
  const size_t ld = live ? lane : n - 1;
  u32 k[8], px[8], py[8], ok, idn = 0u;
  load_rec32_be(k, scalars + ld * 32);
  if (points_fmt == FEC_PT_SEC1_COMPRESSED) {
    uint8_t rec[33];
    for (int i = 0; i < 33; i++) rec[i] = points[ld * 33 + i];
    ok = sec1_decompress_id<C>(px, py, rec, idn);
  } else {
    load_rec32_be(px, points + ld * 64);
    load_rec32_be(py, points + ld * 64 + 32);
    ok = affine_from_plain_id<C>(px, py, idn);
  }
  u32 nn[8];
  C::N(nn);
  const u32 range_bad = geq8(k, nn);
  if (!live) return;  // no warp-collective operation follows
  if (range_bad) atomicOr(sc.bad, (u32)LANE_SCALAR_RANGE);
  else if (!ok && !idn) atomicOr(sc.bad, (u32)LANE_POINT_INVALID);
  if (idn) set_small(k, 0u);  // the identity as an input point contributes nothing: all digits zero (lib.rs:934-951 simply adds it)
  g_st8(sc.aff + lane * 16, px);
  g_st8(sc.aff + lane * 16 + 8, py);
  // k >= 2^255: use (n - k) (-P) instead, so that the top chunk stays below 2^15 and the 17th window (which has a single
  // bucket) almost never receives anything -- otherwise half of all lanes would pile up in it
  const u32 flip = 0u - (k[7] >> 31);
  {
    u32 nk[8];
    sub8(nk, nn, k);
    select8(k, flip, nk, k);
  }
  u32 carry = 0u;
#pragma unroll 1
  for (int w = 0; w < PIP_WINDOWS; w++) {
    const u32 chunk = w < 16 ? ((k[w >> 1] >> (16 * (w & 1))) & 0xFFFFu) : 0u;
    const u32 v = chunk + carry;
    carry = v >> 15 ? 1u : 0u;                        // v >= 2^15: digit v - 2^16, carry 1
    int d = (int)v - (int)(carry << 16);
    if (flip) d = -d;
    sc.dig[(size_t)w * n + lane] = d;
    if (d != 0) atomicAdd(sc.cnt + w * PIP_BUCKETS + (d < 0 ? -d : d), 1u);
  }
}

__global__ void __launch_bounds__(256) k_pip_scatter(size_t n, PipScratch sc) {
  const size_t lane = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (lane >= n) return;
#pragma unroll 1
  for (int w = 0; w < PIP_WINDOWS; w++) {
    const int d = sc.dig[(size_t)w * n + lane];
    if (d == 0) continue;
    const int key = w * PIP_BUCKETS + (d < 0 ? -d : d);
    const u32 pos = atomicAdd(sc.cur + key, 1u);
    sc.ent[sc.off[key] + pos] = (u32)lane | (d < 0 ? 0x80000000u : 0u);
  }
}

__global__ void __launch_bounds__(256) k_pip_maxcount(PipScratch sc) {
  u32 m = 0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < PIP_KEYS; i += gridDim.x * 256) m = max(m, sc.cnt[i]);
  atomicMax(sc.bad + 1, m);
}

template <class C>
__global__ void __launch_bounds__(PIP_THREADS, 3) k_pip_buckets(PipScratch sc) {
  typedef typename C::F F;
  const int key = blockIdx.x * PIP_THREADS + threadIdx.x;  // grid covers PIP_KEYS rounded up to whole warps
  const bool live = key < PIP_KEYS;
  const u32 o = live ? sc.off[key] : 0u, c = live ? sc.cnt[key] : 0u;
  u32 cmax = c;
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) cmax = max(cmax, __shfl_xor_sync(0xFFFFFFFFu, cmax, s));
  Jac acc;
  jac_set_inf<C>(acc);
#pragma unroll 1
  for (u32 i = 0; i < cmax; i++) {
    const bool have = i < c;
    const u32 e = have ? sc.ent[o + i] : 0u;
    const u32* p = sc.aff + (size_t)(e & 0x7FFFFFFFu) * 16;
    u32 x[8], y[8], ny[8];
    g_ld8(x, p);
    g_ld8(y, p + 8);
    F::neg(ny, y);
    select8(y, 0u - (e >> 31), ny, y);
    jac_madd_safe_skip<C>(acc, x, y, have ? 0u : 0xFFFFFFFFu);
  }
  if (live) {
    u32* b = sc.bsum + (size_t)key * 24;
    g_st8(b, acc.X); g_st8(b + 8, acc.Y); g_st8(b + 16, acc.Z);
  }
}

// segment s of window w: buckets lo = 1 + 64 s .. lo + 63:  sum_b b B_b = sum_b (b - lo + 1) B_b + (lo - 1) sum_b B_b
template <class C>
__global__ void __launch_bounds__(PIP_THREADS, 3) k_pip_segments(PipScratch sc) {
  const int t = blockIdx.x * PIP_THREADS + threadIdx.x;  // PIP_WINDOWS * PIP_SEGS threads exactly (a multiple of 128)
  const int w = t / PIP_SEGS, s = t % PIP_SEGS;
  const int lo = 1 + s * PIP_SEG;
  Jac run, tot;
  jac_set_inf<C>(run);
  jac_set_inf<C>(tot);
#pragma unroll 1
  for (int b = lo + PIP_SEG - 1; b >= lo; b--) {
    const u32* p = sc.bsum + ((size_t)w * PIP_BUCKETS + b) * 24;
    Jac q;
    g_ld8(q.X, p); g_ld8(q.Y, p + 8); g_ld8(q.Z, p + 16);
    jac_add_safe<C>(run, q);
    jac_add_safe<C>(tot, run);
  }
  // (lo - 1) * run, lo - 1 = 64 s < 2^15: double-and-add with masks (s is public, but lanes of a warp differ)
  Jac r;
  jac_set_inf<C>(r);
  const u32 m = (u32)(lo - 1);
#pragma unroll 1
  for (int bit = 14; bit >= 0; bit--) {
    jac_dbl<C>(r);
    Jac a = r;
    jac_add_safe<C>(a, run);
    const u32 take = 0u - ((m >> bit) & 1u);
    select8(r.X, take, a.X, r.X); select8(r.Y, take, a.Y, r.Y); select8(r.Z, take, a.Z, r.Z);
  }
  jac_add_safe<C>(tot, r);
  u32* o = sc.ssum + (size_t)t * 24;
  g_st8(o, tot.X); g_st8(o + 8, tot.Y); g_st8(o + 16, tot.Z);
}

template <class C>
__device__ __forceinline__ void pip_warp_sum(Jac& acc) {
#pragma unroll 1
  for (int off = 16; off >= 1; off >>= 1) {
    Jac o;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      o.X[i] = __shfl_xor_sync(0xFFFFFFFFu, acc.X[i], off);
      o.Y[i] = __shfl_xor_sync(0xFFFFFFFFu, acc.Y[i], off);
      o.Z[i] = __shfl_xor_sync(0xFFFFFFFFu, acc.Z[i], off);
    }
    jac_add_safe<C>(acc, o);
  }
}
// one block per window: sum of its segments
template <class C>
__global__ void __launch_bounds__(PIP_THREADS, 1) k_pip_windows(PipScratch sc) {
  __shared__ u32 sm[(PIP_THREADS / 32) * 24];
  const int w = blockIdx.x;
  Jac acc;
  jac_set_inf<C>(acc);
#pragma unroll 1
  for (int s = threadIdx.x; s < PIP_SEGS; s += PIP_THREADS) {  // PIP_SEGS is a multiple of PIP_THREADS
    const u32* p = sc.ssum + ((size_t)w * PIP_SEGS + s) * 24;
    Jac q;
    g_ld8(q.X, p); g_ld8(q.Y, p + 8); g_ld8(q.Z, p + 16);
    jac_add_safe<C>(acc, q);
  }
  pip_warp_sum<C>(acc);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 8; i++) { sm[warp * 24 + i] = acc.X[i]; sm[warp * 24 + 8 + i] = acc.Y[i]; sm[warp * 24 + 16 + i] = acc.Z[i]; }
  }
  __syncthreads();
  if (warp == 0) {
    jac_set_inf<C>(acc);
    if (lane < PIP_THREADS / 32) {
#pragma unroll
      for (int i = 0; i < 8; i++) { acc.X[i] = sm[lane * 24 + i]; acc.Y[i] = sm[lane * 24 + 8 + i]; acc.Z[i] = sm[lane * 24 + 16 + i]; }
    }
    pip_warp_sum<C>(acc);
    if (lane == 0) {
      u32* o = sc.wsum + (size_t)w * 24;
#pragma unroll
      for (int i = 0; i < 8; i++) { o[i] = acc.X[i]; o[8 + i] = acc.Y[i]; o[16 + i] = acc.Z[i]; }
    }
  }
}
// one warp (all lanes compute the same thing, so the warp-uniform helpers stay valid): Horner over the windows
template <class C>
__global__ void __launch_bounds__(32, 1) k_pip_final(PipScratch sc, int out_fmt, uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  typedef typename C::F F;
  Jac acc;
  jac_set_inf<C>(acc);
#pragma unroll 1
  for (int w = PIP_WINDOWS - 1; w >= 0; w--) {
#pragma unroll 1
    for (int d = 0; d < PIP_W; d++) jac_dbl<C>(acc);
    const u32* p = sc.wsum + (size_t)w * 24;
    Jac q;
#pragma unroll
    for (int i = 0; i < 8; i++) { q.X[i] = p[i]; q.Y[i] = p[8 + i]; q.Z[i] = p[16 + i]; }
    jac_add_safe<C>(acc, q);
  }
  const u32 b = sc.bad[0];
  u32 st = (b & LANE_SCALAR_RANGE) ? (u32)LANE_SCALAR_RANGE : ((b & LANE_POINT_INVALID) ? (u32)LANE_POINT_INVALID : 0u);
  const u32 zz = F::is_zero(acc.Z);
  if (st == 0u && zz) st = LANE_INFINITY;
  u32 z[8], one[8], zi[8], x[8], y[8];
  F::one(one);
  select8(z, st != 0u ? 0xFFFFFFFFu : 0u, one, acc.Z);
  F::inv(zi, z);
  jac_to_affine_plain<C>(x, y, acc, zi, 0u);
  if (threadIdx.x == 0) {
    uint8_t rec[65];
    sec1_encode(rec, out_fmt, x, y, st ? 0xFFFFFFFFu : 0u);
    const int osz = out_size(out_fmt);
    for (int i = 0; i < osz; i++) out[i] = rec[i];
    if (status != nullptr) status[0] = (uint8_t)st;
  }
}

size_t msm_pippenger_scratch_bytes(size_t n) {
  size_t scan_tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (u32*)nullptr, (u32*)nullptr, PIP_KEYS + 1);
  return n * 64 + (size_t)PIP_WINDOWS * n * 8 + (size_t)(3 * (PIP_KEYS + 1) + 8) * 4 + (size_t)PIP_KEYS * 96 +
         (size_t)PIP_WINDOWS * PIP_SEGS * 96 + (size_t)PIP_WINDOWS * 96 + scan_tmp + 4096;
}

// All stages on `st`.  *max_bucket (host, optional) receives the largest bucket population after a stream synchronisation
// when `check_skew` is set: the caller falls back to products + reduction for pathologically repeated digits.
template <class C>
static cudaError_t msm_pippenger_t(cudaStream_t st, const uint8_t* scalars, const uint8_t* points, int points_fmt, size_t n, uint8_t* scratch,
                                   int out_fmt, uint8_t* out, uint8_t* status, u32 skew_limit, int* skewed) {
  PipScratch sc;
  uint8_t* p = scratch;
  auto take = [&](size_t bytes) { uint8_t* r = p; p += (bytes + 255) / 256 * 256; return r; };
  sc.aff = (u32*)take(n * 64);
  sc.dig = (int*)take((size_t)PIP_WINDOWS * n * 4);
  sc.ent = (u32*)take((size_t)PIP_WINDOWS * n * 4);
  sc.cnt = (u32*)take((size_t)(PIP_KEYS + 1) * 4);
  sc.off = (u32*)take((size_t)(PIP_KEYS + 1) * 4);
  sc.cur = (u32*)take((size_t)(PIP_KEYS + 1) * 4);
  sc.bsum = (u32*)take((size_t)PIP_KEYS * 96);
  sc.ssum = (u32*)take((size_t)PIP_WINDOWS * PIP_SEGS * 96);
  sc.wsum = (u32*)take((size_t)PIP_WINDOWS * 96);
  sc.bad = (u32*)take(16);
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, sc.cnt, sc.off, PIP_KEYS + 1);
  void* scan_tmp = take(scan_bytes);
  cudaError_t e;
  if ((e = cudaMemsetAsync(sc.cnt, 0, (size_t)(PIP_KEYS + 1) * 4, st)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(sc.cur, 0, (size_t)(PIP_KEYS + 1) * 4, st)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(sc.bad, 0, 16, st)) != cudaSuccess) return e;
  k_pip_digits<C><<<(unsigned)((n + PIP_THREADS - 1) / PIP_THREADS), PIP_THREADS, 0, st>>>(scalars, points, points_fmt, n, sc);
  if ((e = cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, sc.cnt, sc.off, PIP_KEYS + 1, st)) != cudaSuccess) return e;
  if (skewed != nullptr) {
    k_pip_maxcount<<<64, 256, 0, st>>>(sc);
    u32 h[2] = {0, 0};
    if ((e = cudaMemcpyAsync(h, sc.bad, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    *skewed = h[1] > skew_limit ? 1 : 0;
    if (*skewed) return cudaSuccess;
  }
  k_pip_scatter<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, sc);
  k_pip_buckets<C><<<(PIP_KEYS + PIP_THREADS - 1) / PIP_THREADS, PIP_THREADS, 0, st>>>(sc);
  k_pip_segments<C><<<PIP_WINDOWS * PIP_SEGS / PIP_THREADS, PIP_THREADS, 0, st>>>(sc);
  k_pip_windows<C><<<PIP_WINDOWS, PIP_THREADS, 0, st>>>(sc);
  k_pip_final<C><<<1, 32, 0, st>>>(sc, out_fmt, out, status);
  return cudaGetLastError();
}
cudaError_t msm_pippenger_launch(int curve, cudaStream_t st, const uint8_t* scalars, const uint8_t* points, int points_fmt, size_t n,
                                 uint8_t* scratch, int out_fmt, uint8_t* out, uint8_t* status, unsigned skew_limit, int* skewed) {
  return curve == FEC_SECP256K1
             ? msm_pippenger_t<CurveK1>(st, scalars, points, points_fmt, n, scratch, out_fmt, out, status, skew_limit, skewed)
             : msm_pippenger_t<CurveP256>(st, scalars, points, points_fmt, n, scratch, out_fmt, out, status, skew_limit, skewed);
}

}  // namespace fec
