#include "k_wmul_fixed.cuh"
namespace fec {
cudaError_t wmul_fixed_setup_p256() { return wmul_fixed_setup_t<CurveP256>(); }
cudaError_t wmul_fixed_launch_p256(int sms, cudaStream_t st, const uint8_t* s, const u32* comb, size_t n, int of, uint8_t* o,
                                   uint8_t* status) {
  return wmul_fixed_launch_t<CurveP256>(sms, st, s, comb, n, of, o, status);
}
cudaError_t build_comb_launch_p256(cudaStream_t st, u32* comb, u32* bases) { return build_comb_launch_t<CurveP256>(st, comb, bases); }
cudaError_t wmul_fixed_ct_launch_p256(int sms, cudaStream_t st, const uint8_t* s, const u32* comb4, size_t n, int of, uint8_t* o,
                                      uint8_t* status) {
  return wmul_fixed_ct_launch_t<CurveP256>(sms, st, s, comb4, n, of, o, status);
}
cudaError_t build_comb_ct_launch_p256(cudaStream_t st, u32* comb4) { return build_comb_ct_launch_t<CurveP256>(st, comb4); }
}  // namespace fec
