#define FEC_NOINLINE_FIELD 1
#include "k_wmul_var.cuh"
namespace fec {
cudaError_t wmul_var_setup_p256() { return wmul_var_setup_t<CurveP256>(); }
cudaError_t wmul_var_launch_p256(int sms, cudaStream_t st, const uint8_t* s, const uint8_t* p, int pf, size_t n, int of, uint8_t* o,
                                 uint8_t* status) {
  return wmul_var_launch_t<CurveP256>(sms, st, s, p, pf, n, of, o, status);
}
}  // namespace fec
