// k_wmul_var.cuh -- variable-base kernel template; instantiated per curve in k_wmul_var_{k1,p256}.cu
#pragma once
#include "kernels.cuh"

namespace fec {

// ======================================================================================
// kernel: variable-base k*P  (config 3)
// ======================================================================================

template <class C>
__global__ void __launch_bounds__(VAR_THREADS, VAR_BLOCKS_PER_SM)
k_wmul_var(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ points, int points_fmt, size_t n, int out_fmt,
           uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  typedef typename C::F F;
  extern __shared__ __align__(16) u32 smem[];
  uint8_t* io = reinterpret_cast<uint8_t*>(smem);  // aliases the table outside its lifetime
  SmemTab<VAR_THREADS> tab{smem};
  BlockInv<F, VAR_THREADS / 32> inv{smem + VAR_TAB_BYTES / 4};
  const int osz = out_size(out_fmt);
  const size_t ntiles = (n + VAR_THREADS - 1) / VAR_THREADS;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t first = tile * VAR_THREADS;
    const size_t lane = first + threadIdx.x;
    const int cnt = (int)((n - first) < (size_t)VAR_THREADS ? (n - first) : (size_t)VAR_THREADS);
    const bool live = threadIdx.x < cnt;
    u32 k[8], px[8], py[8], K[8], negm, st, ok;
    set_small(k, 1u);
    C::G(px, py);
    ok = 0xFFFFFFFFu;
    if (points_fmt == FEC_PT_SEC1_COMPRESSED) {
      __syncthreads();
      block_copy(io, points + first * 33, (size_t)cnt * 33);
      __syncthreads();
      if (live) ok = sec1_decompress<C>(px, py, io + threadIdx.x * 33);
      __syncthreads();
    } else if (live) {
      load_rec32_be(px, points + lane * 64);
      load_rec32_be(py, points + lane * 64 + 32);
      ok = F::in_range(px) & F::in_range(py) & on_curve<C>(px, py);
    }
    if (live) load_rec32_be(k, scalars + lane * 32);
    {
      u32 gx[8], gy[8];
      C::G(gx, gy);
      select8(px, ok, px, gx);
      select8(py, ok, py, gy);
    }
    scalar_prepare<C>(K, negm, st, k);
    if (st != LANE_SCALAR_RANGE) st = (~ok & LANE_POINT_INVALID) | (ok & st);  // range > invalid point > k = 0
    Jac Q;
    wmul_var_core<C>(Q, K, px, py, tab, inv);
    u32 z[8], zi[8], one[8], x[8], y[8];
    set_small(one, 1u);
    const u32 zz = F::is_zero(Q.Z);
    select8(z, zz, one, Q.Z);
    inv(zi, z);
    jac_to_affine<C>(x, y, Q, zi, negm);
    st |= (st == 0u ? zz & LANE_INFINITY : 0u);
    __syncthreads();  // table is dead: reuse it as the output staging area
    if (live) sec1_encode(io + threadIdx.x * osz, out_fmt, x, y, st ? 0xFFFFFFFFu : 0u);
    __syncthreads();
    block_copy(out + first * osz, io, (size_t)cnt * osz);
    if (status != nullptr && live) status[lane] = (uint8_t)st;
    __syncthreads();
  }
}


template <class C>
cudaError_t wmul_var_setup_t() {
  return cudaFuncSetAttribute(k_wmul_var<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, VAR_SMEM_BYTES);
}
template <class C>
cudaError_t wmul_var_launch_t(int sms, cudaStream_t st, const uint8_t* scalars, const uint8_t* points, int points_fmt, size_t n,
                              int out_fmt, uint8_t* out, uint8_t* status) {
  const int g = persistent_grid(sms, VAR_BLOCKS_PER_SM, (n + VAR_THREADS - 1) / VAR_THREADS);
  k_wmul_var<C><<<g, VAR_THREADS, VAR_SMEM_BYTES, st>>>(scalars, points, points_fmt, n, out_fmt, out, status);
  return cudaGetLastError();
}
cudaError_t wmul_var_setup_k1();
cudaError_t wmul_var_setup_p256();
cudaError_t wmul_var_launch_k1(int, cudaStream_t, const uint8_t*, const uint8_t*, int, size_t, int, uint8_t*, uint8_t*);
cudaError_t wmul_var_launch_p256(int, cudaStream_t, const uint8_t*, const uint8_t*, int, size_t, int, uint8_t*, uint8_t*);

}  // namespace fec
