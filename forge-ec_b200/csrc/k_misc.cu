// k_misc.cu -- field microbenchmark (config 2), multiply-pipe calibration, synthetic inputs
#include "kernels.cuh"
#include "kdf.cuh"

#include <mutex>

namespace fec {

// ======================================================================================
// The library's own stream-ordered memory pool, one per device (r2: the process-wide default pool is left as the application
// configured it).  Scratch of earlier calls stays in the pool (release threshold = max) instead of going back to the driver at
// every synchronisation; scratch_trim gives it back (fec_ctx_destroy).  Pools live as long as the process.
// ======================================================================================
static std::mutex g_pool_mu;
static cudaMemPool_t g_pool[64] = {};
static cudaError_t pool_of(int dev, cudaMemPool_t* out) {
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (g_pool[dev] == nullptr) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool;
    cudaError_t e = cudaMemPoolCreate(&pool, &props);
    if (e != cudaSuccess) return e;
    unsigned long long keep = ~0ull;
    e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    if (e != cudaSuccess) { cudaMemPoolDestroy(pool); return e; }
    g_pool[dev] = pool;
  }
  *out = g_pool[dev];
  return cudaSuccess;
}
cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  cudaMemPool_t pool;
  if ((e = pool_of(dev, &pool)) != cudaSuccess) return e;
  return cudaMallocFromPoolAsync(p, bytes, pool, st);
}
void scratch_trim(int dev) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (dev >= 0 && dev < 64 && g_pool[dev] != nullptr) cudaMemPoolTrimTo(g_pool[dev], 0);
}

// ======================================================================================
// kernel: field microbenchmark (config 2) -- iters dependent operations per lane, in registers
// ======================================================================================
template <class F, int OP>
__global__ void __launch_bounds__(FB_THREADS)
k_field_bench(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out, size_t lanes, int iters) {
  for (size_t lane = (size_t)blockIdx.x * FB_THREADS + threadIdx.x; lane < lanes; lane += (size_t)gridDim.x * FB_THREADS) {
    u32 x[8], y[8], r[8];
    load_rec32_le(x, a + lane * 32);
    load_rec32_le(y, b + lane * 32);
    F::from_plain(x, x);  // P-256 computes in the Montgomery domain; the records are integers
    F::from_plain(y, y);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
      if (OP == FEC_FIELD_MUL) {
        F::mul(x, x, y);
      } else if (OP == FEC_FIELD_SQR) {
        F::sqr(x, x);
      } else {
        F::mul(x, x, y);
        F::sqr(x, x);
      }
    }
    F::to_plain(r, x);
    store_rec32_le(out + lane * 32, r);
  }
}

// ======================================================================================
// kernel: multiply-pipe calibration.  8 independent chains per thread, 64 steps per loop trip.
//   variant 0: IMAD.WIDE.U32      acc64 += a * b
//   variant 1: IMAD + IMAD.HI     lo += a*b ; hi += mulhi(a,b)     (two instructions per IMUL32)
//   variant 2: IMAD.WIDE.U32.X    carry-chained pairs, exactly the field code's instruction
// Each thread reports its accumulators so nothing can be removed.
// ======================================================================================
template <int VARIANT>
__global__ void __launch_bounds__(PEAK_THREADS)
k_imad_peak(u32* __restrict__ sink, u32 seed, int trips) {
  u32 a[PEAK_CHAINS], lo[PEAK_CHAINS], hi[PEAK_CHAINS];
  const u32 t = blockIdx.x * PEAK_THREADS + threadIdx.x;
#pragma unroll
  for (int c = 0; c < PEAK_CHAINS; c++) { a[c] = seed * (t + 1u) + 0x9E3779B9u * (c + 1u); lo[c] = t + c; hi[c] = seed + c; }
  const u32 b = seed | 0x80000001u;
#pragma unroll 1
  for (int it = 0; it < trips; it++) {
#pragma unroll
    for (int s = 0; s < PEAK_INNER; s++) {
      if (VARIANT == 0) {
        // independent (lo,hi) pairs, no carry between pairs: ptxas fuses each into one IMAD.WIDE.U32.
        // multiplicand = the neighbouring chain's high word, so nothing is loop-invariant
#pragma unroll
        for (int c = 0; c < PEAK_CHAINS; c++) {
          const u32 m = hi[(c + 1) % PEAK_CHAINS];
          lo[c] = mad_lo_cc(m, b, lo[c]);
          hi[c] = madc_hi(m, b, hi[c]);
        }
      } else if (VARIANT == 1) {
#pragma unroll
        for (int c = 0; c < PEAK_CHAINS; c++) {
          asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(hi[c]), "r"(b));
          asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[c]) : "r"(lo[c]), "r"(b));
        }
      } else {
        // one carry chain across the 8 pairs, as a row of the schoolbook product
        lo[0] = mad_lo_cc(a[0], b, lo[0]);
        hi[0] = madc_hi_cc(a[0], b, hi[0]);
#pragma unroll
        for (int c = 1; c < PEAK_CHAINS; c++) {
          lo[c] = madc_lo_cc(a[c], b, lo[c]);
          hi[c] = (c == PEAK_CHAINS - 1) ? madc_hi(a[c], b, hi[c]) : madc_hi_cc(a[c], b, hi[c]);
        }
      }
    }
  }
  u32 x = 0;
#pragma unroll
  for (int c = 0; c < PEAK_CHAINS; c++) x ^= lo[c] ^ hi[c];
  sink[t] = x;
}

// ======================================================================================
// kernel: synthetic inputs: Philox-4x32-10, key = seed, counter = (lane lo, lane hi,
// block, stream).  32 bytes per lane = blocks 2*sub and 2*sub+1.  mode 0: raw bytes;
// mode 1/2: big-endian scalar below the order of secp256k1 / P-256 (one conditional subtract),
// global lanes 0..7 replaced by the edge scalars when edge != 0.
// ======================================================================================
__device__ __forceinline__ void philox4x32_10(u32 c[4], u32 k0, u32 k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const u64 p0 = (u64)0xD2511F53u * c[0], p1 = (u64)0xCD9E8D57u * c[2];
    const u32 n0 = (u32)(p1 >> 32) ^ c[1] ^ k0, n1 = (u32)p1, n2 = (u32)(p0 >> 32) ^ c[3] ^ k1, n3 = (u32)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
template <class C>
__device__ __forceinline__ void synth_scalar(u32 w[8], size_t glane, int edge) {
  // w holds the 32 raw bytes as little-endian words of the byte string; value = big-endian integer
  u32 v[8], n[8], d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __byte_perm(w[7 - i], 0, 0x0123);
  C::N(n);
  const u32 bw = sub8(d, v, n);
  select8(v, bw - 1u, d, v);
  if (edge && glane < 8) {
    u32 e[8];
    set_small(e, (u32)glane);                                   // 0,1,2,3
    if (glane == 4) { set_small(e, 1u); sub8(e, n, e); }         // n-1
    if (glane == 5) { set_small(e, 2u); sub8(e, n, e); }         // n-2
    if (glane == 6) {                                            // (n+1)/2
      u32 one[8]; set_small(one, 1u);
      u32 c = add8(e, n, one);
#pragma unroll
      for (int i = 0; i < 7; i++) e[i] = (e[i] >> 1) | (e[i + 1] << 31);
      e[7] = (e[7] >> 1) | (c << 31);
    }
    if (glane == 7) {                                            // 2^255 mod n = 2^255 (n > 2^255)
      set_small(e, 0u); e[7] = 0x80000000u;
    }
    copy8(v, e);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) w[7 - i] = __byte_perm(v[i], 0, 0x0123);
}
__global__ void __launch_bounds__(256)
k_synth(uint8_t* __restrict__ out, size_t n, u32 seed_lo, u32 seed_hi, u32 stream, u32 sub, unsigned long long first_lane,
        int mode, int edge) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const unsigned long long g = first_lane + i;
    u32 w[8];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      u32 c[4] = {(u32)g, (u32)(g >> 32), 2u * sub + h, stream};
      philox4x32_10(c, seed_lo, seed_hi);
      w[4 * h] = c[0]; w[4 * h + 1] = c[1]; w[4 * h + 2] = c[2]; w[4 * h + 3] = c[3];
    }
    if (mode == 1) synth_scalar<CurveK1>(w, (size_t)g, edge);
    if (mode == 2) synth_scalar<CurveP256>(w, (size_t)g, edge);
    store_rec32_le(out + i * 32, w);
  }
}


template <class F>
static void field_bench_t(int op, int g, cudaStream_t st, const uint8_t* a, const uint8_t* b, uint8_t* o, size_t lanes, int iters) {
  if (op == FEC_FIELD_MUL) k_field_bench<F, FEC_FIELD_MUL><<<g, FB_THREADS, 0, st>>>(a, b, o, lanes, iters);
  else if (op == FEC_FIELD_SQR) k_field_bench<F, FEC_FIELD_SQR><<<g, FB_THREADS, 0, st>>>(a, b, o, lanes, iters);
  else k_field_bench<F, FEC_FIELD_MULSQR><<<g, FB_THREADS, 0, st>>>(a, b, o, lanes, iters);
}
cudaError_t field_bench_launch(int sms, cudaStream_t st, int prime, int op, const uint8_t* a, const uint8_t* b, uint8_t* out,
                               size_t lanes, int iters) {
  const int g = persistent_grid(sms, 8, (lanes + FB_THREADS - 1) / FB_THREADS);
  if (prime == FEC_PRIME_SECP256K1) field_bench_t<FeK1>(op, g, st, a, b, out, lanes, iters);
  else if (prime == FEC_PRIME_P256) field_bench_t<FeP256>(op, g, st, a, b, out, lanes, iters);
  else field_bench_t<Fe25519>(op, g, st, a, b, out, lanes, iters);
  return cudaGetLastError();
}
cudaError_t imad_peak_launch(int variant, int blocks, cudaStream_t st, u32* sink, u32 seed, int trips) {
  if (variant == 0) k_imad_peak<0><<<blocks, PEAK_THREADS, 0, st>>>(sink, seed, trips);
  else if (variant == 1) k_imad_peak<1><<<blocks, PEAK_THREADS, 0, st>>>(sink, seed, trips);
  else k_imad_peak<2><<<blocks, PEAK_THREADS, 0, st>>>(sink, seed, trips);
  return cudaGetLastError();
}
cudaError_t synth_launch(int sms, cudaStream_t st, uint8_t* out, size_t n, u32 seed_lo, u32 seed_hi, u32 stream_id, u32 sub,
                         unsigned long long first_lane, int mode, int edge) {
  const int g = persistent_grid(sms, 8, (n + 255) / 256);
  k_synth<<<g, 256, 0, st>>>(out, n, seed_lo, seed_hi, stream_id, sub, first_lane, mode, edge);
  return cudaGetLastError();
}

// ======================================================================================
// ECDH tail: key derivation from the 32-byte shared secrets (KeyExchange::derive_key)
// ======================================================================================
// One lane per thread, everything in registers (kdf.cuh); the lane-independent message blocks arrive as a kernel
// parameter.  Output records of a multiple of 16 bytes are written with 128-bit stores.
__global__ void __launch_bounds__(128)
k_kdf(int kdf, const uint8_t* __restrict__ x_be, const uint8_t* __restrict__ status, size_t n, const __grid_constant__ KdfPlan plan,
      uint8_t* __restrict__ out) {
  const size_t lane = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (lane >= n) return;
  u32 w[8], sec[8], okm[16];
  load_rec32_le(w, x_be + lane * 32);
#pragma unroll
  for (int i = 0; i < 8; i++) sec[i] = __byte_perm(w[i], 0, 0x0123);  // big-endian words of the byte string
#pragma unroll
  for (int i = 8; i < 16; i++) okm[i] = 0u;
  kdf_lane(okm, kdf, sec, plan);
  const u32 dead = (status != nullptr && status[lane] != 0) ? 0u : 0xFFFFFFFFu;
#pragma unroll
  for (int i = 0; i < 16; i++) okm[i] = __byte_perm(okm[i], 0, 0x0123) & dead;  // memory order
  const int out_len = plan.out_len;
  uint8_t* o = out + lane * (size_t)out_len;
  if ((out_len & 15) == 0 && (reinterpret_cast<size_t>(out) & 15) == 0) {
    uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (16 * q < out_len) o4[q] = make_uint4(okm[4 * q], okm[4 * q + 1], okm[4 * q + 2], okm[4 * q + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 64; i++)
      if (i < out_len) o[i] = (uint8_t)(okm[i >> 2] >> (8 * (i & 3)));
  }
}
cudaError_t kdf_launch(cudaStream_t st, int kdf, const uint8_t* x_be, const uint8_t* status, size_t n, const uint8_t* info, int info_len,
                       int out_len, uint8_t* out) {
  if (n == 0) return cudaSuccess;
  KdfPlan plan;
  kdf_plan_build(plan, kdf, info, info_len, out_len);
  k_kdf<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(kdf, x_be, status, n, plan, out);
  return cudaGetLastError();
}

}  // namespace fec
