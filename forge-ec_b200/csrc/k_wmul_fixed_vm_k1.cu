#define FEC_NOINLINE_FIELD 1
#include "k_wmul_fixed_vm.cuh"
namespace fec {
cudaError_t wmul_fixed_vm_setup_k1() { return wmul_fixed_vm_setup_t<CurveK1>(); }
cudaError_t wmul_fixed_vm_launch_k1(int sms, cudaStream_t st, const uint8_t* s, const u32* comb, size_t n, int of, uint8_t* o,
                                    uint8_t* status) {
  return wmul_fixed_vm_launch_t<CurveK1>(sms, st, s, comb, n, of, o, status);
}
}  // namespace fec
