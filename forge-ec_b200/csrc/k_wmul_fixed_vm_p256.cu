#define FEC_NOINLINE_FIELD 1
#include "k_wmul_fixed_vm.cuh"
namespace fec {
cudaError_t wmul_fixed_vm_setup_p256() { return wmul_fixed_vm_setup_t<CurveP256>(); }
cudaError_t wmul_fixed_vm_launch_p256(int sms, cudaStream_t st, const uint8_t* s, const u32* comb, size_t n, int of, uint8_t* o,
                                      uint8_t* status) {
  return wmul_fixed_vm_launch_t<CurveP256>(sms, st, s, comb, n, of, o, status);
}
}  // namespace fec
