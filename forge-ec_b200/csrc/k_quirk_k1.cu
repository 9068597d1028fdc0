// k_quirk_k1.cu -- FEC_SEM_FORGE_QUIRK kernels: the reference's literal secp256k1, P-256, x25519() and Ed25519 key generation code
// paths, one lane per thread (see quirk_k1.cuh, quirk_p256.cuh, quirk_x25519.cuh, quirk_ed25519.cuh).  A parity artefact for BASELINE.json configs[0], not a fast path.
#include "kernels.cuh"
#include "quirk_k1.cuh"
#include "quirk_p256.cuh"
#include "quirk_x25519.cuh"
#include "quirk_ed25519.cuh"

namespace fec {

__global__ void __launch_bounds__(128)
k_quirk_k1(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ points, size_t n, int out_fmt,
           uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  const size_t lane = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (lane >= n) return;
  uint8_t k[32], xy[64], o[65];
  for (int i = 0; i < 32; i++) k[i] = scalars[lane * 32 + i];
  if (points != nullptr)
    for (int i = 0; i < 64; i++) xy[i] = points[lane * 64 + i];
  const int st = quirk::q_lane(k, points != nullptr ? xy : nullptr, out_fmt, o);
  const int osz = out_size(out_fmt);
  for (int i = 0; i < osz; i++) out[lane * osz + i] = o[i];
  if (status != nullptr) status[lane] = (uint8_t)st;
}

cudaError_t quirk_k1_launch(cudaStream_t st, const uint8_t* scalars, const uint8_t* points, size_t n, int out_fmt, uint8_t* out,
                            uint8_t* status) {
  k_quirk_k1<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(scalars, points, n, out_fmt, out, status);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(128)
k_quirk_p256(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ points, size_t n, int out_fmt,
             uint8_t* __restrict__ out, uint8_t* __restrict__ status) {
  const size_t lane = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (lane >= n) return;
  uint8_t k[32], xy[64], o[65];
  for (int i = 0; i < 32; i++) k[i] = scalars[lane * 32 + i];
  if (points != nullptr)
    for (int i = 0; i < 64; i++) xy[i] = points[lane * 64 + i];
  const int st = quirk256::p_lane(k, points != nullptr ? xy : nullptr, out_fmt, o);
  const int osz = out_size(out_fmt);
  for (int i = 0; i < osz; i++) out[lane * osz + i] = o[i];
  if (status != nullptr) status[lane] = (uint8_t)st;
}

cudaError_t quirk_p256_launch(cudaStream_t st, const uint8_t* scalars, const uint8_t* points, size_t n, int out_fmt, uint8_t* out,
                              uint8_t* status) {
  k_quirk_p256<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(scalars, points, n, out_fmt, out, status);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(128)
k_quirk_x25519(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ us, size_t n, uint8_t* __restrict__ out) {
  const size_t lane = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (lane >= n) return;
  uint8_t k[32], u[32], o[32];
  for (int i = 0; i < 32; i++) k[i] = scalars[lane * 32 + i];
  if (us != nullptr)
    for (int i = 0; i < 32; i++) u[i] = us[lane * 32 + i];
  quirk25519::x_lane(k, us != nullptr ? u : nullptr, o);
  for (int i = 0; i < 32; i++) out[lane * 32 + i] = o[i];
}

cudaError_t quirk_x25519_launch(cudaStream_t st, const uint8_t* scalars, const uint8_t* us, size_t n, uint8_t* out) {
  k_quirk_x25519<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(scalars, us, n, out);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(128)
k_quirk_ed25519(const uint8_t* __restrict__ seeds, size_t n, uint8_t* __restrict__ out) {
  const size_t lane = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (lane >= n) return;
  uint8_t s[32], o[32];
  for (int i = 0; i < 32; i++) s[i] = seeds[lane * 32 + i];
  quirked::e_lane(s, o);
  for (int i = 0; i < 32; i++) out[lane * 32 + i] = o[i];
}

cudaError_t quirk_ed25519_launch(cudaStream_t st, const uint8_t* seeds, size_t n, uint8_t* out) {
  k_quirk_ed25519<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(seeds, n, out);
  return cudaGetLastError();
}

}  // namespace fec
