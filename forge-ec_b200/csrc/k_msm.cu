// k_msm.cu -- multi-scalar multiplication sum_i k_i * P_i: the reduction that follows the batched
// k*P kernels.  Replaces Curve::multi_scalar_multiply (forge-ec-core/src/lib.rs:934-951, p256.rs:2193-2211), which
// adds the products one after the other on one core.
//
// The products come from the batched path as affine X || Y records with their lane status (fec_batch_mul on the same
// stream); here every thread folds a strided subset of them into a Jacobian accumulator (mixed additions), warps
// combine by shuffles and blocks through shared memory (full Jacobian additions), and a single-block second pass
// adds the per-block partial sums, normalises once and encodes.  Lanes whose product is the identity (k = 0)
// contribute nothing; a lane with an out-of-range scalar or an invalid point makes the whole call fail
// (status byte), exactly as a caller of the reference could not have built such a Scalar / PointAffine.
#define FEC_NOINLINE_FIELD 1
// (P-256 stays canonical in this translation unit: its kernels are dominated by general additions with masks, where the weak form's
// rare-branch bodies cost more than the shorter reduction saves -- 288 against 277 M points/s; every value crossing to other kernels is
// a byte record)
#define FEC_P256_CANONICAL
#include "kernels.cuh"

namespace fec {

constexpr int MSM_THREADS = 128;

// all 32 lanes of the warp end up with the warp's sum
template <class C>
__device__ __forceinline__ void warp_sum(Jac& acc) {
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
// thread 0 of the block ends up with the block's sum; sm: (MSM_THREADS / 32) * 24 words
template <class C>
__device__ __forceinline__ void block_sum(Jac& acc, u32* sm) {
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  warp_sum<C>(acc);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 8; i++) { sm[warp * 24 + i] = acc.X[i]; sm[warp * 24 + 8 + i] = acc.Y[i]; sm[warp * 24 + 16 + i] = acc.Z[i]; }
  }
  __syncthreads();
  if (warp == 0) {
    jac_set_inf<C>(acc);
    if (lane < MSM_THREADS / 32) {
#pragma unroll
      for (int i = 0; i < 8; i++) { acc.X[i] = sm[lane * 24 + i]; acc.Y[i] = sm[lane * 24 + 8 + i]; acc.Z[i] = sm[lane * 24 + 16 + i]; }
    }
    warp_sum<C>(acc);
  }
}

// partial sums of the products (xy: n x 64 B affine big-endian, st: n status bytes) -> partials[gridDim.x][24], *bad |= status
template <class C>
__global__ void __launch_bounds__(MSM_THREADS, 3)
k_msm_partial(const uint8_t* __restrict__ xy, const uint8_t* __restrict__ st, size_t n, u32* __restrict__ partials, u32* __restrict__ bad) {
  __shared__ u32 sm[(MSM_THREADS / 32) * 24];
  Jac acc;
  jac_set_inf<C>(acc);
  const size_t stride = (size_t)gridDim.x * MSM_THREADS;
  const size_t rounds = (n + stride - 1) / stride;  // uniform trip count: no lane leaves the loop early
  u32 worst = 0u;
#pragma unroll 1
  for (size_t r = 0; r < rounds; r++) {
    const size_t lane = r * stride + (size_t)blockIdx.x * MSM_THREADS + threadIdx.x;
    const bool live = lane < n;
    const size_t ld = live ? lane : n - 1;
    u32 x[8], y[8];
    load_rec32_be(x, xy + ld * 64);
    load_rec32_be(y, xy + ld * 64 + 32);
    C::F::from_plain(x, x);  // the products are byte records (integers); no-op except for P-256
    C::F::from_plain(y, y);
    const u32 s = live ? (u32)st[ld] : (u32)LANE_INFINITY;
    if (s == LANE_SCALAR_RANGE || s == LANE_POINT_INVALID) worst |= s;
    jac_madd_safe_skip<C>(acc, x, y, s != 0u ? 0xFFFFFFFFu : 0u);
  }
  if (worst) atomicOr(bad, worst);
  block_sum<C>(acc, sm);
  if (threadIdx.x == 0) {
    u32* o = partials + (size_t)blockIdx.x * 24;
#pragma unroll
    for (int i = 0; i < 8; i++) { o[i] = acc.X[i]; o[8 + i] = acc.Y[i]; o[16 + i] = acc.Z[i]; }
  }
}
// one block: sum of m partial sums, one inversion, encoding.  status: 0 ok, 1/2 a lane was rejected, 3 the sum is the identity
template <class C>
__global__ void __launch_bounds__(MSM_THREADS, 1)
k_msm_final(const u32* __restrict__ partials, int m, const u32* __restrict__ bad, int out_fmt, uint8_t* __restrict__ out,
            uint8_t* __restrict__ status) {
  typedef typename C::F F;
  __shared__ u32 sm[(MSM_THREADS / 32) * 24];
  Jac acc;
  jac_set_inf<C>(acc);
  const int rounds = (m + MSM_THREADS - 1) / MSM_THREADS;
#pragma unroll 1
  for (int r = 0; r < rounds; r++) {
    const int j = r * MSM_THREADS + (int)threadIdx.x;
    Jac q;
    jac_set_inf<C>(q);
    if (j < m) {
      const u32* p = partials + (size_t)j * 24;
#pragma unroll
      for (int i = 0; i < 8; i++) { q.X[i] = p[i]; q.Y[i] = p[8 + i]; q.Z[i] = p[16 + i]; }
    }
    jac_add_safe<C>(acc, q);
  }
  block_sum<C>(acc, sm);
  if (threadIdx.x == 0) {  // one lane; the branch depends on public data only
    const u32 b = *bad;
    u32 st = (b & LANE_SCALAR_RANGE) ? (u32)LANE_SCALAR_RANGE : ((b & LANE_POINT_INVALID) ? (u32)LANE_POINT_INVALID : 0u);
    const u32 zz = F::is_zero(acc.Z);
    if (st == 0u && zz) st = LANE_INFINITY;
    u32 z[8], one[8], zi[8], x[8], y[8];
    F::one(one);
    select8(z, st != 0u ? 0xFFFFFFFFu : 0u, one, acc.Z);
    F::inv(zi, z);
    jac_to_affine_plain<C>(x, y, acc, zi, 0u);
    uint8_t rec[65];
    sec1_encode(rec, out_fmt, x, y, st ? 0xFFFFFFFFu : 0u);
    const int osz = out_size(out_fmt);
    for (int i = 0; i < osz; i++) out[i] = rec[i];
    if (status != nullptr) status[0] = (uint8_t)st;
  }
}

// scratch: partials (blocks x 24 words) followed by one flag word; blocks <= msm_blocks(sms, n)
int msm_blocks(int sms, size_t n) {
  size_t b = (n + MSM_THREADS - 1) / MSM_THREADS, cap = (size_t)sms * 3;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}
template <class C>
static cudaError_t msm_reduce_t(int sms, cudaStream_t st, const uint8_t* xy, const uint8_t* lane_status, size_t n, u32* scratch,
                                int out_fmt, uint8_t* out, uint8_t* status) {
  const int blocks = msm_blocks(sms, n);
  u32* bad = scratch + (size_t)blocks * 24;
  cudaError_t e = cudaMemsetAsync(bad, 0, 4, st);
  if (e != cudaSuccess) return e;
  k_msm_partial<C><<<blocks, MSM_THREADS, 0, st>>>(xy, lane_status, n, scratch, bad);
  k_msm_final<C><<<1, MSM_THREADS, 0, st>>>(scratch, blocks, bad, out_fmt, out, status);
  return cudaGetLastError();
}
cudaError_t msm_reduce_launch(int curve, int sms, cudaStream_t st, const uint8_t* xy, const uint8_t* lane_status, size_t n, u32* scratch,
                              int out_fmt, uint8_t* out, uint8_t* status) {
  return curve == FEC_SECP256K1 ? msm_reduce_t<CurveK1>(sms, st, xy, lane_status, n, scratch, out_fmt, out, status)
                                : msm_reduce_t<CurveP256>(sms, st, xy, lane_status, n, scratch, out_fmt, out, status);
}

}  // namespace fec
