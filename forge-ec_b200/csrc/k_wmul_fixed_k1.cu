#include "k_wmul_fixed.cuh"
namespace fec {
cudaError_t wmul_fixed_setup_k1() { return wmul_fixed_setup_t<CurveK1>(); }
cudaError_t wmul_fixed_launch_k1(int sms, cudaStream_t st, const uint8_t* s, const u32* comb, size_t n, int of, uint8_t* o,
                                 uint8_t* status) {
  return wmul_fixed_launch_t<CurveK1>(sms, st, s, comb, n, of, o, status);
}
cudaError_t build_comb_launch_k1(cudaStream_t st, u32* comb, u32* bases) { return build_comb_launch_t<CurveK1>(st, comb, bases); }
cudaError_t wmul_fixed_setup() {
  cudaError_t e = wmul_fixed_setup_k1();
  return e != cudaSuccess ? e : wmul_fixed_setup_p256();
}
cudaError_t wmul_fixed_launch(int curve, int sms, cudaStream_t st, const uint8_t* s, const u32* comb, size_t n, int of, uint8_t* o,
                              uint8_t* status) {
  return curve == FEC_SECP256K1 ? wmul_fixed_launch_k1(sms, st, s, comb, n, of, o, status)
                                : wmul_fixed_launch_p256(sms, st, s, comb, n, of, o, status);
}
cudaError_t build_comb_launch(int curve, cudaStream_t st, u32* comb, u32* bases) {
  return curve == FEC_SECP256K1 ? build_comb_launch_k1(st, comb, bases) : build_comb_launch_p256(st, comb, bases);
}
}  // namespace fec
