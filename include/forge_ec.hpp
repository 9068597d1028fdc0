// forge_ec.hpp -- C++ host side above the C ABI (forge_ec_cuda.h): the reference's trait surface for the batched path,
// with the reference's names, argument meaning and error behaviour, so that code written against forge-ec reads the
// same.  Header-only; link with -lforge_ec_cuda.  There is no CPU fallback: Engine's constructor throws when no sm_100
// device is usable.
//
//   forge-ec (Rust)                                                         here
//   forge_ec_core::Error                        lib.rs:70-104               forge_ec::Error (an exception carrying the variant)
//   Scalar::from_bytes / to_bytes               secp256k1.rs:2271-2297      Scalar::from_bytes / to_bytes (32 B big-endian)
//   PointAffine::{x,y,is_identity,to_bytes,to_bytes_with_format}
//                                               secp256k1.rs:875-894, 1010-1067     AffinePoint
//   Curve::{generator,multiply,to_affine,multi_scalar_multiply}
//                                               lib.rs:784-952              Secp256k1 / P256 static members
//   KeyExchange::{derive_shared_secret,derive_key}  secp256k1.rs:1846-1905, p256.rs:2281-2344
//   x25519(scalar, u)                           curve25519.rs:1624-1716     forge_ec::x25519 / batch_x25519
//   Ed25519Signature::derive_public_key         forge-ec-signature/src/eddsa.rs:450-509
//   (new) BatchMul::batch_mul                   SURVEY.md section 8b        Curve::batch_mul
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "forge_ec_cuda.h"

namespace forge_ec {

// forge_ec_core::Error (lib.rs:70-104), the variants this path can produce
enum class ErrorKind { InvalidEncoding, InvalidScalar, PointNotOnCurve, UnsupportedOperation, GenericError };
class Error : public std::runtime_error {
 public:
  Error(ErrorKind k, const std::string& what) : std::runtime_error(what), kind_(k) {}
  ErrorKind kind() const { return kind_; }

 private:
  ErrorKind kind_;
};
inline void check(int rc, const char* where) {
  if (rc == FEC_OK) return;
  const ErrorKind k = rc == FEC_ERR_INVALID_ARG ? ErrorKind::InvalidEncoding
                      : (rc == FEC_ERR_NO_DEVICE || rc == FEC_ERR_UNSUPPORTED) ? ErrorKind::UnsupportedOperation
                                                                              : ErrorKind::GenericError;
  throw Error(k, std::string(where) + ": " + fec_strerror(rc));
}
inline void check_lane(uint8_t status, const char* where) {
  if (status == FEC_LANE_SCALAR_RANGE) throw Error(ErrorKind::InvalidScalar, std::string(where) + ": scalar out of range");
  if (status == FEC_LANE_POINT_INVALID) throw Error(ErrorKind::PointNotOnCurve, std::string(where) + ": point not on curve");
}

// owns the devices, streams, tables and scratch; one per device set (the reference is stateless)
class Engine {
 public:
  Engine() { check(fec_ctx_create(nullptr, 0, &ctx_), "fec_ctx_create"); }
  explicit Engine(const std::vector<int>& devices) {
    check(fec_ctx_create(devices.data(), (int)devices.size(), &ctx_), "fec_ctx_create");
  }
  ~Engine() { fec_ctx_destroy(ctx_); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  fec_ctx* ctx() const { return ctx_; }
  int device_count() const { return fec_ctx_device_count(ctx_); }
  uint64_t launch_count() const { return fec_ctx_launch_count(ctx_); }

 private:
  fec_ctx* ctx_ = nullptr;
};

struct Scalar {
  std::array<uint8_t, 32> be{};  // big-endian, as Scalar::to_bytes (trait) writes it
  static Scalar from_bytes(const uint8_t b[32]) { Scalar s; std::memcpy(s.be.data(), b, 32); return s; }
  static Scalar from_u64(uint64_t v) {
    Scalar s;
    for (int i = 0; i < 8; i++) s.be[31 - i] = (uint8_t)(v >> (8 * i));
    return s;
  }
  const std::array<uint8_t, 32>& to_bytes() const { return be; }
};

enum class PointFormat { Compressed, Uncompressed, Hybrid };  // forge-ec-core/src/lib.rs:536-543
struct AffinePoint {
  std::array<uint8_t, 64> xy{};  // X || Y big-endian (FieldElement::to_bytes twice); all zero = the identity
  bool infinity = true;
  static AffinePoint from_xy(const uint8_t b[64]) {
    AffinePoint p;
    std::memcpy(p.xy.data(), b, 64);
    p.infinity = false;
    for (int i = 0; i < 64 && !p.infinity; i++) if (b[i]) return p;
    p.infinity = true;  // the all-zero record is the identity (forge-ec-encoding/src/point.rs:41-47)
    return p;
  }
  bool is_identity() const { return infinity; }
  std::array<uint8_t, 32> x() const { std::array<uint8_t, 32> r; std::memcpy(r.data(), xy.data(), 32); return r; }
  std::array<uint8_t, 32> y() const { std::array<uint8_t, 32> r; std::memcpy(r.data(), xy.data() + 32, 32); return r; }
  // SEC1: 02|03 || X (33 B), 04 || X || Y (65 B) or the hybrid 06|07 || X || Y (65 B, secp256k1.rs:1048-1066); the identity
  // is all zero (secp256k1.rs:875-894, point.rs:41-47)
  std::vector<uint8_t> to_bytes_with_format(PointFormat f) const {
    std::vector<uint8_t> o(f == PointFormat::Compressed ? 33 : 65, 0);
    if (infinity) return o;
    if (f == PointFormat::Compressed) {
      o[0] = (uint8_t)(2 | (xy[63] & 1));
      std::memcpy(o.data() + 1, xy.data(), 32);
    } else {
      o[0] = f == PointFormat::Hybrid ? (uint8_t)(6 | (xy[63] & 1)) : (uint8_t)4;
      std::memcpy(o.data() + 1, xy.data(), 64);
    }
    return o;
  }
  std::vector<uint8_t> to_bytes() const { return to_bytes_with_format(PointFormat::Compressed); }
};

template <fec_curve C>
struct WeierstrassCurve {
  // BatchMul::batch_mul: k_i * G when points == nullptr, else k_i * P_i.  A rejected lane raises the error
  // the reference's per-call API would (InvalidScalar / PointNotOnCurve); k = 0 gives the identity.
  static std::vector<AffinePoint> batch_mul(Engine& e, const std::vector<AffinePoint>* points, const std::vector<Scalar>& scalars) {
    const size_t n = scalars.size();
    if (points != nullptr && points->size() != n) throw Error(ErrorKind::InvalidEncoding, "batch_mul: points and scalars differ in length");
    std::vector<uint8_t> k(n * 32), p, out(n * 64), st(n);
    for (size_t i = 0; i < n; i++) std::memcpy(&k[i * 32], scalars[i].be.data(), 32);
    if (points != nullptr) {
      p.resize(n * 64);
      for (size_t i = 0; i < n; i++) {
        // the identity travels as the all-zero record and k * identity is the identity (Curve::multiply, lib.rs:820-832)
        if (!(*points)[i].infinity) std::memcpy(&p[i * 64], (*points)[i].xy.data(), 64);
      }
    }
    if (n == 0) return {};
    check(fec_batch_mul(e.ctx(), C, FEC_SEM_STD, k.data(), points ? p.data() : nullptr, FEC_PT_AFFINE_XY_BE, n, FEC_OUT_AFFINE_XY_BE,
                        out.data(), st.data()), "fec_batch_mul");
    std::vector<AffinePoint> r(n);
    for (size_t i = 0; i < n; i++) {
      check_lane(st[i], "batch_mul");
      if (st[i] == FEC_LANE_OK) r[i] = AffinePoint::from_xy(&out[i * 64]);
    }
    return r;
  }
  static AffinePoint generator(Engine& e) { return multiply(e, nullptr, Scalar::from_u64(1)); }
  // CompressedPoint::to_affine + Curve::validate_point for a batch of 33-byte SEC1 records (point.rs:70-144, lib.rs:905-923)
  static std::vector<AffinePoint> batch_decode(Engine& e, const std::vector<std::array<uint8_t, 33>>& records) {
    const size_t n = records.size();
    std::vector<uint8_t> out(n * 64), st(n);
    std::vector<AffinePoint> r(n);
    if (n == 0) return r;
    check(fec_batch_point_decode(e.ctx(), C, records[0].data(), FEC_PT_SEC1_COMPRESSED, n, FEC_OUT_AFFINE_XY_BE, out.data(), st.data()),
          "fec_batch_point_decode");
    for (size_t i = 0; i < n; i++) {
      if (st[i] == FEC_LANE_POINT_INVALID) throw Error(ErrorKind::InvalidEncoding, "batch_decode: not a point of the curve");
      if (st[i] == FEC_LANE_OK) r[i] = AffinePoint::from_xy(&out[i * 64]);
    }
    return r;
  }
  // Curve::multiply + to_affine (lib.rs:820-832), one lane
  static AffinePoint multiply(Engine& e, const AffinePoint* point, const Scalar& k) {
    if (point == nullptr) return batch_mul(e, nullptr, {k})[0];
    const std::vector<AffinePoint> pts{*point};
    return batch_mul(e, &pts, {k})[0];
  }
  // Curve::multi_scalar_multiply (lib.rs:934-951)
  static AffinePoint multi_scalar_multiply(Engine& e, const std::vector<AffinePoint>& points, const std::vector<Scalar>& scalars) {
    const size_t n = scalars.size();
    if (points.size() != n || n == 0) throw Error(ErrorKind::InvalidEncoding, "multi_scalar_multiply: empty or mismatched input");
    std::vector<uint8_t> k(n * 32), p(n * 64);
    for (size_t i = 0; i < n; i++) {
      std::memcpy(&k[i * 32], scalars[i].be.data(), 32);
      if (!points[i].infinity) std::memcpy(&p[i * 64], points[i].xy.data(), 64);  // identity: all-zero record, adds nothing
    }
    uint8_t out[64], st = 0;
    check(fec_msm(e.ctx(), C, k.data(), p.data(), FEC_PT_AFFINE_XY_BE, n, FEC_OUT_AFFINE_XY_BE, out, &st), "fec_msm");
    check_lane(st, "multi_scalar_multiply");
    return st == FEC_LANE_OK ? AffinePoint::from_xy(out) : AffinePoint{};
  }
  // KeyExchange::derive_shared_secret: the x-coordinate of k * P; the identity is an error (secp256k1.rs:1885-1905)
  static std::array<uint8_t, 32> derive_shared_secret(Engine& e, const Scalar& private_key, const AffinePoint& public_key) {
    std::array<uint8_t, 32> out{};
    uint8_t st = 0;
    check(fec_batch_ecdh(e.ctx(), C, private_key.be.data(), public_key.xy.data(), FEC_PT_AFFINE_XY_BE, 1, FEC_KDF_NONE, nullptr, 0, 32,
                         out.data(), &st), "fec_batch_ecdh");
    check_lane(st, "derive_shared_secret");
    if (st == FEC_LANE_INFINITY) throw Error(ErrorKind::GenericError, "derive_shared_secret: the shared point is the identity");
    return out;
  }
  // n exchanges followed by the curve's derive_key (HKDF-SHA-256 for secp256k1, the XOR placeholder for P-256)
  static std::vector<uint8_t> batch_ecdh(Engine& e, const std::vector<Scalar>& scalars, const std::vector<AffinePoint>& points,
                                         const std::vector<uint8_t>& info, size_t out_len, std::vector<uint8_t>* status = nullptr) {
    const size_t n = scalars.size();
    if (points.size() != n) throw Error(ErrorKind::InvalidEncoding, "batch_ecdh: points and scalars differ in length");
    std::vector<uint8_t> k(n * 32), p(n * 64), out(n * out_len), st(n);
    for (size_t i = 0; i < n; i++) {
      std::memcpy(&k[i * 32], scalars[i].be.data(), 32);
      std::memcpy(&p[i * 64], points[i].xy.data(), 64);
    }
    if (n == 0) return out;
    check(fec_batch_ecdh(e.ctx(), C, k.data(), p.data(), FEC_PT_AFFINE_XY_BE, n, C == FEC_SECP256K1 ? FEC_KDF_HKDF_SHA256 : FEC_KDF_XOR_INFO,
                         info.data(), info.size(), out_len, out.data(), st.data()), "fec_batch_ecdh");
    if (status != nullptr) *status = st;
    return out;
  }
};
using Secp256k1 = WeierstrassCurve<FEC_SECP256K1>;
using P256 = WeierstrassCurve<FEC_P256>;

// x25519(scalar, u) (curve25519.rs:1624-1716); us == nullptr: the base point u = 9
inline std::vector<std::array<uint8_t, 32>> batch_x25519(Engine& e, const std::vector<std::array<uint8_t, 32>>& scalars,
                                                         const std::vector<std::array<uint8_t, 32>>* us) {
  const size_t n = scalars.size();
  if (us != nullptr && us->size() != n) throw Error(ErrorKind::InvalidEncoding, "batch_x25519: scalars and u differ in length");
  std::vector<std::array<uint8_t, 32>> out(n);
  if (n == 0) return out;
  check(fec_batch_x25519(e.ctx(), scalars[0].data(), us ? (*us)[0].data() : nullptr, n, out[0].data()), "fec_batch_x25519");
  return out;
}
inline std::array<uint8_t, 32> x25519(Engine& e, const std::array<uint8_t, 32>& scalar, const std::array<uint8_t, 32>& u) {
  const std::vector<std::array<uint8_t, 32>> us{u};
  return batch_x25519(e, {scalar}, &us)[0];
}

// The fixed-base step of `sign` for a batch of nonces: ECDSA r = x(k G) (forge-ec-signature/src/ecdsa.rs:108-131),
// BIP-340 x-only R and the even-y rule (schnorr.rs:372-385), EdDSA R = [r mod L]B from 64-byte hash values (eddsa.rs:120-131).
struct NonceCommit {
  std::vector<uint8_t> R, r, status;  // n x {33 | 32} B, n x 32 B, n lane status bytes
};
inline NonceCommit batch_nonce_commit(Engine& e, fec_curve curve, fec_nonce_mode mode, const std::vector<uint8_t>& nonces) {
  const size_t rec_in = mode == FEC_NONCE_EDDSA ? 64 : 32;
  if (nonces.size() % rec_in != 0) throw Error(ErrorKind::InvalidEncoding, "batch_nonce_commit: ragged nonce buffer");
  const size_t n = nonces.size() / rec_in;
  NonceCommit o;
  o.R.resize(n * (mode == FEC_NONCE_ECDSA ? 33 : 32));
  o.r.resize(n * 32);
  o.status.resize(n);
  if (n == 0) return o;
  check(fec_batch_nonce_commit(e.ctx(), curve, mode, nonces.data(), n, o.R.data(), o.r.data(), o.status.data()), "fec_batch_nonce_commit");
  return o;
}

struct Ed25519Signature {
  // seed -> SHA-512 -> clamp -> [s]B -> 32 bytes (forge-ec-signature/src/eddsa.rs:450-509, RFC 8032 5.1.5)
  static std::vector<std::array<uint8_t, 32>> batch_derive_public_key(Engine& e, const std::vector<std::array<uint8_t, 32>>& seeds) {
    std::vector<std::array<uint8_t, 32>> out(seeds.size());
    if (seeds.empty()) return out;
    check(fec_batch_ed25519_pubkey(e.ctx(), seeds[0].data(), 1, seeds.size(), out[0].data()), "fec_batch_ed25519_pubkey");
    return out;
  }
  static std::array<uint8_t, 32> derive_public_key(Engine& e, const std::array<uint8_t, 32>& seed) {
    return batch_derive_public_key(e, {seed})[0];
  }
};

}  // namespace forge_ec
