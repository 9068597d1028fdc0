#include "k_wmul_fixed.cuh"
namespace fec {
cudaError_t wmul_fixed_setup_k1() { return wmul_fixed_setup_t<CurveK1>(); }
cudaError_t wmul_fixed_launch_k1(int sms, cudaStream_t st, const uint8_t* s, const u32* comb, size_t n, int of, uint8_t* o,
                                 uint8_t* status) {
  return wmul_fixed_launch_t<CurveK1>(sms, st, s, comb, n, of, o, status);
}
cudaError_t build_comb_launch_k1(cudaStream_t st, u32* comb, u32* bases) { return build_comb_launch_t<CurveK1>(st, comb, bases); }
cudaError_t wmul_fixed_ct_launch_k1(int sms, cudaStream_t st, const uint8_t* s, const u32* comb4, size_t n, int of, uint8_t* o,
                                    uint8_t* status) {
  return wmul_fixed_ct_launch_t<CurveK1>(sms, st, s, comb4, n, of, o, status);
}
cudaError_t build_comb_ct_launch_k1(cudaStream_t st, u32* comb4) { return build_comb_ct_launch_t<CurveK1>(st, comb4); }
// FORGE_EC_CT_TABLE=1: the constant-time table form of k*G (4-bit comb in shared memory, masked scans): see lanes.cuh, DESIGN.md
bool fixed_ct_table() {
  static const bool v = [] { const char* e = getenv("FORGE_EC_CT_TABLE"); return e != nullptr && e[0] == '1'; }();
  return v;
}
cudaError_t wmul_fixed_ct_launch(int curve, int sms, cudaStream_t st, const uint8_t* s, const u32* comb4, size_t n, int of, uint8_t* o,
                                 uint8_t* status) {
  return curve == FEC_SECP256K1 ? wmul_fixed_ct_launch_k1(sms, st, s, comb4, n, of, o, status)
                                : wmul_fixed_ct_launch_p256(sms, st, s, comb4, n, of, o, status);
}
cudaError_t build_comb_ct_launch(int curve, cudaStream_t st, u32* comb4) {
  return curve == FEC_SECP256K1 ? build_comb_ct_launch_k1(st, comb4) : build_comb_ct_launch_p256(st, comb4);
}
cudaError_t wmul_fixed_setup() {
  cudaError_t e = wmul_fixed_setup_k1();
  if (e == cudaSuccess) e = wmul_fixed_vm_setup_k1();
  if (e == cudaSuccess) e = wmul_fixed_vm_setup_p256();
  return e != cudaSuccess ? e : wmul_fixed_setup_p256();
}
// Which kernel serves a curve was measured (2^22 lanes): the accumulator-machine kernel wins where the field is bound by
// the ALU pipe (P-256: 351 -> 441 M/s), the register-resident kernel with the field routines in line where it is bound
// by the multiply pipe (secp256k1: 493 vs 483 M/s).  FORGE_EC_FIXED_IMPL=regs|vm forces one of them (A/B measurements).
static int fixed_impl() {
  static const int v = [] { const char* e = getenv("FORGE_EC_FIXED_IMPL"); return e == nullptr ? 0 : (e[0] == 'r' ? 1 : (e[0] == 'v' ? 2 : 0)); }();
  return v;
}
cudaError_t wmul_fixed_launch(int curve, int sms, cudaStream_t st, const uint8_t* s, const u32* comb, size_t n, int of, uint8_t* o,
                              uint8_t* status) {
  const int impl = fixed_impl();
  // (r2) below 2^17 lanes the register-resident kernel also wins for P-256: a batch that does not fill the device is bound by
  // the latency of a lane, not by a pipe (2^14 lanes: 135 against 113 M/s, 2^16: 266 against 246, 2^18: 367 against 378)
  const bool vm = impl == 2 || (impl == 0 && curve == FEC_P256 && n >= (size_t(1) << 17));
  if (vm)
    return curve == FEC_SECP256K1 ? wmul_fixed_vm_launch_k1(sms, st, s, comb, n, of, o, status)
                                  : wmul_fixed_vm_launch_p256(sms, st, s, comb, n, of, o, status);
  return curve == FEC_SECP256K1 ? wmul_fixed_launch_k1(sms, st, s, comb, n, of, o, status)
                                : wmul_fixed_launch_p256(sms, st, s, comb, n, of, o, status);
}
cudaError_t build_comb_launch(int curve, cudaStream_t st, u32* comb, u32* bases) {
  return curve == FEC_SECP256K1 ? build_comb_launch_k1(st, comb, bases) : build_comb_launch_p256(st, comb, bases);
}
}  // namespace fec
