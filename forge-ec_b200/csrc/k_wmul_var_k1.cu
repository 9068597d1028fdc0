#define FEC_NOINLINE_FIELD 1
#include <cstdlib>
#include "k_wmul_var.cuh"
namespace fec {
cudaError_t wmul_var_setup_k1() { return wmul_var_setup_t<CurveK1>(); }
cudaError_t wmul_var_launch_k1(int sms, cudaStream_t st, const uint8_t* s, const uint8_t* p, int pf, size_t n, int of, uint8_t* o,
                               uint8_t* status) {
  return wmul_var_launch_t<CurveK1>(sms, st, s, p, pf, n, of, o, status);
}
cudaError_t wmul_var_setup() {
  cudaError_t e = wmul_var_setup_k1();
  if (e == cudaSuccess) e = wmul_glv_setup_k1();
  if (e == cudaSuccess) e = wmul_vm_setup_p256();
  return e != cudaSuccess ? e : wmul_var_setup_p256();
}
// Development switch (A/B measurements only): FORGE_EC_VAR_IMPL=mono selects the single-kernel paths.
static bool var_mono() {
  static const bool v = [] { const char* e = getenv("FORGE_EC_VAR_IMPL"); return e != nullptr && e[0] == 'm'; }();
  return v;
}
cudaError_t wmul_var_launch(int curve, int sms, cudaStream_t st, const uint8_t* s, const uint8_t* p, int pf, size_t n, int of,
                            uint8_t* o, uint8_t* status, int* launches, StageEvents* se) {
  if (curve == FEC_SECP256K1 && !var_mono()) return wmul_glv_launch_k1(st, s, p, pf, n, of, o, status, GLV_CHUNK_LANES, launches, se);
  if (curve == FEC_P256 && !var_mono()) return wmul_vm_launch_p256(st, s, p, pf, n, of, o, status, GLV_CHUNK_LANES, launches, se);
  if (launches) *launches += 1;
  return curve == FEC_SECP256K1 ? wmul_var_launch_k1(sms, st, s, p, pf, n, of, o, status)
                                : wmul_var_launch_p256(sms, st, s, p, pf, n, of, o, status);
}
}  // namespace fec
