// hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the device headers of forge-ec_b200/csrc with g++ (the PTX carry flag is emulated in
// u256.cuh under #ifndef __CUDACC__) so that the per-lane limb arithmetic, recoding and encoders
// can be compared with the oracle on a machine that has no GPU.  Nothing here is linked into
// libforge_ec_cuda.so and nothing in the product calls it.
#include <cstring>
#include <vector>
#include "../../forge-ec_b200/csrc/lanes.cuh"
#include "../../forge-ec_b200/csrc/glv3.cuh"
#include "../../forge-ec_b200/csrc/p256v3.cuh"
#include "../../forge-ec_b200/csrc/quirk_k1.cuh"
#include "../../forge-ec_b200/csrc/quirk_p256.cuh"
#include "../../forge-ec_b200/csrc/quirk_x25519.cuh"
#include "../../forge-ec_b200/csrc/x25519_vm.cuh"
#include "../../forge-ec_b200/csrc/quirk_ed25519.cuh"
#include "../../forge-ec_b200/csrc/kdf.cuh"

using namespace fec;

namespace {
bool g_use_vm = true;  // variable base: three-stage accumulator-machine path (default) or the single-kernel path
struct HostTab {
  u32 e[8][16];
  void store(int j, const u32 x[8], const u32 y[8]) { copy8(e[j], x); copy8(e[j] + 8, y); }
  void load(u32 j, u32 x[8], u32 y[8]) const { copy8(x, e[j]); copy8(y, e[j] + 8); }
};
template <class F> struct HostInv {
  void operator()(u32 zi[8], const u32 z[8]) { F::inv(zi, z); }
};
// the 16 x 32768 tables are far too large to build on the host for a test: entries are computed on demand with the
// same routines the device table builders use (comb_base / comb_entry_jac, ed_comb_base / ed_comb_entry)
template <class C> struct HostComb {
  u32 bases[COMB_WINDOWS][16];
  HostComb() {
    for (int i = 0; i < COMB_WINDOWS; i++) comb_base<C>(bases[i], bases[i] + 8, i);
  }
  void operator()(int i, u32 idx, u32 x[8], u32 y[8]) const {
    typedef typename C::F F;
    Jac R;
    comb_entry_jac<C>(R, bases[i], bases[i] + 8, (int)idx);
    u32 zi[8];
    F::inv(zi, R.Z);
    jac_to_affine<C>(x, y, R, zi, 0u);
  }
};
// the same entries behind the address interface of wmul_fixed_vm_core (a small ring of buffers: an address stays valid
// while the next entry is being fetched)
template <class C> struct HostCombEntry {
  const HostComb<C>& comb;
  mutable u32 ring[4][16];
  mutable int next = 0;
  const u32* operator()(int i, u32 idx) const {
    u32* e = ring[next++ & 3];
    comb(i, idx, e, e + 8);
    return e;
  }
};
struct HostNiels {
  EdExt bases[COMB_WINDOWS];
  HostNiels() {
    for (int i = 0; i < COMB_WINDOWS; i++) ed_comb_base(bases[i], i);
  }
  void operator()(int i, u32 idx, u32 a[8], u32 b[8], u32 c[8]) const { ed_comb_entry(a, b, c, bases[i], (int)idx); }
};

// the 4-bit Niels comb of the constant-time form, entries built as k_build_niels_ct does and read through ScanNiels
struct HostNiels4 {
  u32* tab;
  ScanNiels scan;
  HostNiels4() {
    tab = new u32[CT_NIELS_WORDS];
    for (int i = 0; i < CT_WINDOWS; i++) {
      EdExt Bi;
      ed_comb_base<CT_W>(Bi, i);
      for (int j = 0; j < CT_ENTRIES; j++) {
        u32* o = tab + ((size_t)i * CT_ENTRIES + j) * 24;
        ed_comb_entry<CT_W>(o, o + 8, o + 16, Bi, j);
      }
    }
    scan.base = tab;
  }
  void operator()(int i, u32 idx, u32 a[8], u32 b[8], u32 c[8]) const { scan(i, idx, a, b, c); }
};

template <class F> void field_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* o) {
  u32 x[8], y[8], r[8], t[8];
  load_le32(x, a); load_le32(y, b);
  F::from_plain(x, x); F::from_plain(y, y);  // byte records are integers; FeP256 computes in the Montgomery domain
  switch (op) {
    case 0: F::mul(t, x, y); break;
    case 1: F::sqr(t, x); break;
    case 2: F::add(t, x, y); break;
    case 3: F::sub(t, x, y); break;
    case 4: F::neg(t, x); break;
    case 5: F::inv(t, x); break;
    case 7: F::mul2(t, x); break;
    case 8: F::mul3(t, x); break;
    case 9: F::mul4(t, x); break;
    case 10: F::mul8(t, x); break;
    case 11: { u32 m = F::is_zero(x); set_small(t, 0u); if (m & 1u) F::one(t); } break;
    default: F::dbl(t, x); break;
  }
  F::to_plain(r, t);
  store_le32(o, r);
}

template <class C>
int wmul(const uint8_t* k_be, const uint8_t* pt, int pfmt, int fixed, int ofmt, uint8_t* out) {
  typedef typename C::F F;
  static HostComb<C> comb;
  u32 k[8], K[8], negm, st, px[8], py[8];
  load_be32(k, k_be);
  scalar_prepare<C>(K, negm, st, k);
  Jac Q;
  if (fixed && g_use_vm) {
    std::vector<u32> slots(VS_COUNT * 8);
    HostCombEntry<C> entry{comb};
    wmul_fixed_vm_core<C, 1>(slots.data(), K, entry);
    copy8(Q.X, &slots[VS_X * 8]); copy8(Q.Y, &slots[VS_Y * 8]); copy8(Q.Z, &slots[VS_Z * 8]);
  } else if (fixed) {
    wmul_fixed_core<C>(Q, K, comb);
  } else {
    u32 ok;
    if (pfmt == 1) {
      ok = sec1_decompress<C>(px, py, pt);
    } else {
      load_be32(px, pt); load_be32(py, pt + 32);
      ok = affine_from_plain<C>(px, py);
    }
    u32 gx[8], gy[8];
    C::G(gx, gy);
    select8(px, ok, px, gx); select8(py, ok, py, gy);
    if (st != LANE_SCALAR_RANGE) st = (~ok & LANE_POINT_INVALID) | (ok & st);
    HostTab tab;
    HostInv<F> inv;
    if (C::CURVE_ID == 0) {  // secp256k1: GLV + effective-affine table, exactly as k_wmul_var does
      u32 nn[8], one1[8], zs[8];
      C::N(nn);
      set_small(one1, 1u);
      const u32 bad = geq8(k, nn), zero = is_zero8(k);
      select8(k, bad | zero, one1, k);
      negm = 0u;
      if (g_use_vm) {
        // the three stages of k_wmul_glv_k1.cu with T = 1: table in "global" memory, slots in "shared" memory
        std::vector<u32> gtab(GLV_TAB_WORDS), slots(VS_COUNT * 8);
        u32 KK[8], flags;
        glv_prepare(gtab.data(), zs, KK, flags, k, px, py);
        glv_ladder<1>(slots.data(), gtab.data(), KK, flags);
        copy8(Q.X, &slots[VS_X * 8]); copy8(Q.Y, &slots[VS_Y * 8]); copy8(Q.Z, &slots[VS_Z * 8]);
        glv_finish(Q, gtab.data(), zs, flags);
      } else {
        wmul_var_glv_core(Q, zs, k, px, py, tab);
        F::mul(Q.Z, Q.Z, zs);
      }
    } else if (g_use_vm) {
      // the stages of k_wmul_vm_p256.cu with T = 1 (the sign is applied by jac_to_affine_plain below, as k_p3_finish does)
      std::vector<u32> gtab(P3_TAB_WORDS), slots(VS_COUNT * 8);
      GlobalLaneTab gt{gtab.data()};
      wmul_var_table<C>(gt, px, py, inv);
      p3_ladder<1>(slots.data(), gtab.data(), K);
      copy8(Q.X, &slots[VS_X * 8]); copy8(Q.Y, &slots[VS_Y * 8]); copy8(Q.Z, &slots[VS_Z * 8]);
    } else {
      wmul_var_core<C>(Q, K, px, py, tab, inv);
    }
  }
  u32 z[8], zi[8], one[8], x[8], y[8];
  F::one(one);
  u32 zz = F::is_zero(Q.Z);
  select8(z, zz, one, Q.Z);
  F::inv(zi, z);
  jac_to_affine_plain<C>(x, y, Q, zi, negm);
  if (zz && st == 0) st = LANE_INFINITY;
  sec1_encode(out, ofmt, x, y, st ? 0xFFFFFFFFu : 0u);
  return (int)st;
}
// the fused chains of linear operations of FeK1 (fields.cuh: sub2_dbl, sub_twice, sub3, sub_mul8, cneg_sub; mul_pow2<4>): operands are taken as they
// are (weak representatives included), the result is canonicalised
void k1_fused(int op, const uint8_t* t, const uint8_t* a, const uint8_t* b, uint8_t* o) {
  u32 x[8], y[8], z[8], r[8], c[8];
  load_le32(x, t); load_le32(y, a); load_le32(z, b);
  switch (op) {
    case 0: FeK1::sub2_dbl(r, x, y, z); break;
    case 1: FeK1::sub_twice(r, x, y); break;
    case 2: FeK1::sub3(r, x, y, z); break;
    case 3: FeK1::sub_mul8(r, x, y); break;
    case 4: FeK1::cneg_sub(r, x, 0u, y); break;
    case 5: FeK1::cneg_sub(r, x, 0xFFFFFFFFu, y); break;
    case 6: FeK1::template mul_pow2<4>(r, x); break;
    default: {  // D = 2 (t - a - b) kept unfolded, then (t - 2D) + (D - a) through the raw consumers = D + t - a ... checked as 2(t-a-b) + t - a
      u32 d[8], v, x3[8], dm[8];
      FeK1::sub2_dbl_raw(d, v, x, y, z);
      FeK1::sub_twice_raw(x3, x, d, v);   // t - 2D
      FeK1::rsub_raw(dm, d, v, y);        // D - a
      FeK1::add(r, x3, dm);
      break;
    }
  }
  FeK1::canon(c, r);
  store_le32(o, c);
}
}  // namespace

extern "C" {
void hs_k1_fused(int op, const uint8_t* t, const uint8_t* a, const uint8_t* b, uint8_t* o) { k1_fused(op, t, a, b, o); }
// FeP256 on RAW Montgomery-domain limbs (weak representatives included), raw result: op 0 mul, 1 sqr, 2 add, 3 sub, 4 neg, 5 half,
// 6 canon, 7 is_zero (0 / 1), 8 to_plain, 9 mul3, 10 mul8
void hs_p256_raw(int op, const uint8_t* a, const uint8_t* b, uint8_t* o) {
  u32 x[8], y[8], r[8];
  load_le32(x, a); load_le32(y, b);
  switch (op) {
    case 0: FeP256::mul_inl(r, x, y); break;
    case 1: FeP256::sqr_inl(r, x); break;
    case 2: FeP256::add(r, x, y); break;
    case 3: FeP256::sub(r, x, y); break;
    case 4: FeP256::neg(r, x); break;
    case 5: FeP256::half(r, x); break;
    case 6: FeP256::canon(r, x); break;
    case 7: set_small(r, FeP256::is_zero(x) & 1u); break;
    case 8: FeP256::to_plain(r, x); break;
    case 9: FeP256::mul3(r, x); break;
    default: FeP256::mul8(r, x); break;
  }
  store_le32(o, r);
}
// Fe25519 under FEC_25519_FOLD255 (the X25519 ladder kernel): op 0 / 1 = add_r1 / sub_r1 (canonicalised result), op 2 / 3 =
// mul / sqr with the RAW result (the caller checks its range: below 2^255 + 2^11 in that build)
void hs_f25519_r1(int op, const uint8_t* a, const uint8_t* b, uint8_t* o) {
  u32 x[8], y[8], r[8], c[8];
  load_le32(x, a); load_le32(y, b);
  if (op == 0) { Fe25519::add_r1(r, x, y); Fe25519::canon(c, r); }
  else if (op == 1) { Fe25519::sub_r1(r, x, y); Fe25519::canon(c, r); }
  else if (op == 2) { Fe25519::mul_inl(c, x, y); }
  else { Fe25519::sqr_inl(c, x); }
  store_le32(o, c);
}
void hs_field_op(int prime, int op, const uint8_t* a, const uint8_t* b, uint8_t* o) {
  if (prime == 0) field_op<FeK1>(op, a, b, o);
  else if (prime == 1) field_op<FeP256>(op, a, b, o);
  else field_op<Fe25519>(op, a, b, o);
}
int hs_quirk_k1(const uint8_t* k_be, const uint8_t* xy_be, int ofmt, uint8_t* out) { return quirk::q_lane(k_be, xy_be, ofmt, out); }
void hs_quirk_ed25519(const uint8_t* seed, uint8_t* out) { quirked::e_lane(seed, out); }
void hs_quirk_x25519(const uint8_t* k, const uint8_t* u, uint8_t* out) { quirk25519::x_lane(k, u, out); }
int hs_quirk_p256(const uint8_t* k_be, const uint8_t* xy_be, int ofmt, uint8_t* out) { return quirk256::p_lane(k_be, xy_be, ofmt, out); }
void hs_kdf(int kdf, const uint8_t* secret32, const uint8_t* info, int info_len, int out_len, uint8_t* okm) {
  u32 sec[8];
  for (int i = 0; i < 8; i++)
    sec[i] = ((u32)secret32[4 * i] << 24) | ((u32)secret32[4 * i + 1] << 16) | ((u32)secret32[4 * i + 2] << 8) | secret32[4 * i + 3];
  KdfPlan plan;  // the per-call part the host prepares (kdf_launch does the same), then one lane of k_kdf
  kdf_plan_build(plan, kdf, info, info_len, out_len);
  u32 o[16] = {0};
  kdf_lane(o, kdf, sec, plan);
  for (int i = 0; i < out_len; i++) okm[i] = (uint8_t)(o[i >> 2] >> (24 - 8 * (i & 3)));
}
void hs_set_var_impl(int use_vm) { g_use_vm = use_vm != 0; }
int hs_wmul(int curve, const uint8_t* k_be, const uint8_t* pt, int pfmt, int fixed, int ofmt, uint8_t* out) {
  return curve == 0 ? wmul<CurveK1>(k_be, pt, pfmt, fixed, ofmt, out) : wmul<CurveP256>(k_be, pt, pfmt, fixed, ofmt, out);
}
void hs_x25519(const uint8_t* k, const uint8_t* u, uint8_t* out) {
  u32 kk[8], uu[8], x2[8], z2[8], zi[8], t[8], r[8], one[8];
  load_le32(kk, k); load_le32(uu, u);
  x25519_core(x2, z2, kk, uu);
  set_small(one, 1u);
  u32 zz = Fe25519::is_zero(z2);
  select8(t, zz, one, z2);
  Fe25519::inv(zi, t);
  Fe25519::mul(t, x2, zi);
  Fe25519::canon(r, t);
  for (int i = 0; i < 8; i++) r[i] &= ~zz;
  store_le32(out, r);
}
// the accumulator-machine form of the ladder (x25519_vm.cuh; k_x25519_vm on the device), one lane, T = 1
void hs_x25519_vm(const uint8_t* k, const uint8_t* u, uint8_t* out) {
  u32 kk[8], uu[8], x2[8], z2[8], zi[8], t[8], r[8], one[8];
  u32 slots[XS_COUNT * 8];
  load_le32(kk, k); load_le32(uu, u);
  x25519_vm_core<1>(slots, x2, z2, kk, uu);
  set_small(one, 1u);
  u32 zz = Fe25519::is_zero(z2);
  select8(t, zz, one, z2);
  Fe25519::inv(zi, t);
  Fe25519::mul(t, x2, zi);
  Fe25519::canon(r, t);
  for (int i = 0; i < 8; i++) r[i] &= ~zz;
  store_le32(out, r);
}
void hs_ed25519_pubkey(const uint8_t* in, int is_seed, uint8_t* out) {
  static HostNiels niels;
  u32 s[8], K[8], zi[8], o[8];
  if (is_seed) ed25519_expand_seed(s, in); else load_le32(s, in);
  ed_scalar_prepare(K, s);
  EdExt Q;
  ed_comb_core(Q, K, niels);
  Fe25519::inv(zi, Q.Z);
  ed_encode(o, Q, zi);
  store_le32(out, o);
}
void hs_ed25519_pubkey_ct(const uint8_t* in, int is_seed, uint8_t* out) {  // constant-time table form (FORGE_EC_CT_TABLE=1)
  static HostNiels4 niels;
  u32 s[8], K[8], zi[8], o[8];
  if (is_seed) ed25519_expand_seed(s, in); else load_le32(s, in);
  ed_scalar_prepare(K, s);
  EdExt Q;
  ed_comb_core<HostNiels4, CT_W>(Q, K, niels);
  Fe25519::inv(zi, Q.Z);
  ed_encode(o, Q, zi);
  store_le32(out, o);
}
void hs_sha512_32(const uint8_t* msg, uint8_t* digest) {
  u64 h[8];
  sha512_32(h, msg);
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) digest[8 * i + j] = (uint8_t)(h[i] >> (56 - 8 * j));
}
}
