// test_traits.cpp -- the C++ host side (include/forge_ec.hpp) exercised the way the reference's own unit tests exercise
// its traits: generator and small multiples against SEC2 / FIPS 186-4 values (secp256k1.rs:2769-2786, p256.rs:2498-2556),
// ECDH commutes, n-1 and out-of-range scalars, an off-curve point, multi_scalar_multiply, the RFC 7748 6.1 exchange
// (curve25519.rs:2061-2089) and RFC 8032 7.1 TEST 1 (eddsa.rs:572-585).
// Exit code 0: all checks passed on the GPU; 77: no usable device and the library said so loudly (UnsupportedOperation).
#include <algorithm>
#include <cstdio>
#include <string>

#include "forge_ec.hpp"

using namespace forge_ec;

static int failures = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); failures++; } \
  } while (0)

template <size_t N>
static std::array<uint8_t, N> hex(const char* s) {
  std::array<uint8_t, N> r{};
  for (size_t i = 0; i < N; i++) { unsigned v; std::sscanf(s + 2 * i, "%2x", &v); r[i] = (uint8_t)v; }
  return r;
}
template <class V>
static std::string to_hex(const V& v) {
  std::string s;
  char b[3];
  for (uint8_t c : v) { std::snprintf(b, sizeof b, "%02x", c); s += b; }
  return s;
}

template <class Curve>
static void weierstrass(Engine& e, const char* g1, const char* g2, const char* g3, const char* gdead, const char* order_hex) {
  const AffinePoint G = Curve::generator(e);
  CHECK(!G.is_identity());
  CHECK(to_hex(G.to_bytes()) == g1);
  CHECK(G.to_bytes_with_format(PointFormat::Uncompressed)[0] == 4 && G.to_bytes_with_format(PointFormat::Uncompressed).size() == 65);
  CHECK(to_hex(Curve::multiply(e, nullptr, Scalar::from_u64(2)).to_bytes()) == g2);
  CHECK(to_hex(Curve::multiply(e, &G, Scalar::from_u64(3)).to_bytes()) == g3);          // variable-base path on P = G
  CHECK(to_hex(Curve::multiply(e, nullptr, Scalar::from_u64(0xdeadbeefULL)).to_bytes()) == gdead);
  // k = 0 is the identity and encodes as zeros (point.rs:41-47)
  const AffinePoint O = Curve::multiply(e, nullptr, Scalar::from_u64(0));
  CHECK(O.is_identity() && to_hex(O.to_bytes()) == std::string(66, '0'));
  // (n-1) G = -G: same x, other y parity; n is out of range -> InvalidScalar
  auto n = hex<32>(order_hex);
  Scalar nm1 = Scalar::from_bytes(n.data());
  nm1.be[31] -= 1;
  const AffinePoint mG = Curve::multiply(e, nullptr, nm1);
  CHECK(mG.x() == G.x() && mG.y() != G.y());
  bool threw = false;
  try { Curve::multiply(e, nullptr, Scalar::from_bytes(n.data())); } catch (const Error& err) { threw = err.kind() == ErrorKind::InvalidScalar; }
  CHECK(threw);
  // a point that is not on the curve -> PointNotOnCurve
  uint8_t bad[64] = {0};
  bad[31] = 2; bad[63] = 2;
  const AffinePoint B = AffinePoint::from_xy(bad);
  threw = false;
  try { Curve::multiply(e, &B, Scalar::from_u64(5)); } catch (const Error& err) { threw = err.kind() == ErrorKind::PointNotOnCurve; }
  CHECK(threw);
  // ECDH commutes (p256.rs:2540-2556): a (b G) == b (a G), as points and as shared secrets
  const Scalar a = Scalar::from_u64(0x1234567ULL), b = Scalar::from_u64(0xdeadbeefULL);
  const AffinePoint A = Curve::multiply(e, nullptr, a), Bp = Curve::multiply(e, nullptr, b);
  CHECK(Curve::multiply(e, &Bp, a).xy == Curve::multiply(e, &A, b).xy);
  CHECK(Curve::derive_shared_secret(e, a, Bp) == Curve::derive_shared_secret(e, b, A));
  CHECK(Curve::derive_shared_secret(e, a, Bp) == Curve::multiply(e, &Bp, a).x());
  // multi_scalar_multiply: 2 G + 3 (2G) + 5 (3G) = 23 G, and a sum that cancels
  const AffinePoint G2 = Curve::multiply(e, nullptr, Scalar::from_u64(2)), G3 = Curve::multiply(e, nullptr, Scalar::from_u64(3));
  const AffinePoint S = Curve::multi_scalar_multiply(e, {G, G2, G3}, {Scalar::from_u64(2), Scalar::from_u64(3), Scalar::from_u64(5)});
  CHECK(S.xy == Curve::multiply(e, nullptr, Scalar::from_u64(23)).xy);
  CHECK(Curve::multi_scalar_multiply(e, {G, G}, {Scalar::from_u64(1), nm1}).is_identity());
  // batch form: lanes are independent and keep their order
  std::vector<Scalar> ks;
  for (uint64_t i = 1; i <= 70; i++) ks.push_back(Scalar::from_u64(i * 0x9e3779b97f4a7c15ULL));
  const std::vector<AffinePoint> pub = Curve::batch_mul(e, nullptr, ks);
  std::vector<AffinePoint> pts(ks.size(), Bp);
  const std::vector<AffinePoint> sh = Curve::batch_mul(e, &pts, ks);
  for (size_t i = 0; i < ks.size(); i += 23) {
    CHECK(pub[i].xy == Curve::multiply(e, nullptr, ks[i]).xy);
    CHECK(sh[i].xy == Curve::multiply(e, &pub[i], b).xy);   // k_i (b G) == b (k_i G)
  }
  // batched ECDH with the curve's derive_key: both sides agree
  const std::vector<uint8_t> info{'c', 'p', 'p'};
  std::vector<AffinePoint> pubs2(ks.size(), A);
  std::vector<uint8_t> st1, st2;
  const auto k1 = Curve::batch_ecdh(e, ks, pubs2, info, 32, &st1);
  const auto k2 = Curve::batch_ecdh(e, std::vector<Scalar>(ks.size(), a), pub, info, 32, &st2);
  CHECK(k1 == k2);
  for (uint8_t s : st1) CHECK(s == 0);
  // r2: the hybrid encoding (PointFormat::Hybrid, secp256k1.rs:1048-1066): 06|07 || X || Y with the parity of y
  const auto hyb = G.to_bytes_with_format(PointFormat::Hybrid);
  CHECK(hyb.size() == 65 && hyb[0] == (6 | (G.xy[63] & 1)) && std::equal(hyb.begin() + 1, hyb.end(), G.xy.begin()));
  // r2: the identity as an input point: k * O = O (Curve::multiply, lib.rs:820-832), and a sum simply skips it
  CHECK(Curve::multiply(e, &O, Scalar::from_u64(5)).is_identity());
  CHECK(Curve::multi_scalar_multiply(e, {G, O, G2}, {Scalar::from_u64(2), Scalar::from_u64(7), Scalar::from_u64(3)}).xy ==
        Curve::multiply(e, nullptr, Scalar::from_u64(8)).xy);
  // r2: CompressedPoint::to_affine for a batch (point.rs:70-144): compressed -> affine round trip, a bad prefix is refused
  std::array<uint8_t, 33> c1{}, c2{};
  const auto gb = G.to_bytes(), g2b = G2.to_bytes();
  std::copy(gb.begin(), gb.end(), c1.begin());
  std::copy(g2b.begin(), g2b.end(), c2.begin());
  const auto dec = Curve::batch_decode(e, {c1, c2});
  CHECK(dec[0].xy == G.xy && dec[1].xy == G2.xy);
  c2[0] = 5;
  threw = false;
  try { Curve::batch_decode(e, {c1, c2}); } catch (const Error& err) { threw = err.kind() == ErrorKind::InvalidEncoding; }
  CHECK(threw);
}

int main() {
  try {
    Engine e;
    weierstrass<Secp256k1>(e, "0279be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798",
                           "02c6047f9441ed7d6d3045406e95c07cd85c778e4b8cef3ca7abac09b95c709ee5",
                           "02f9308a019258c31049344f85f89d5229b531c845836f99b08601f113bce036f9",
                           "0276d2fdf1302d1fa9556f4df94ec84cefba6d482e54f47c6c2a238c1baa560f0e",
                           "fffffffffffffffffffffffffffffffebaaedce6af48a03bbfd25e8cd0364141");
    weierstrass<P256>(e, "036b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296",
                      "037cf27b188d034f7e8a52380304b51ac3c08969e277f21b35a60b48fc47669978",
                      "025ecbe4d1a6330a44c8f7ef951d4bf165e6c6b721efada985fb41661bc6e7fd6c",
                      "02b487d183dc4806058eb31a29bedefd7bcca987b77a381a3684871d8449c18394",
                      "ffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551");
    // RFC 7748 6.1 (the vectors the reference quotes at curve25519.rs:2061-2089)
    const auto alice = hex<32>("77076d0a7318a57d3c16c17251b26645df4c2f87ebc0992ab177fba51db92c2a");
    const auto bob = hex<32>("5dab087e624a8a4b79e17f8b83800ee66f3bb1292618b6fd1c2f8b27ff88e0eb");
    std::array<uint8_t, 32> nine{};
    nine[0] = 9;
    const auto alice_pub = x25519(e, alice, nine), bob_pub = batch_x25519(e, {bob}, nullptr)[0];
    CHECK(to_hex(alice_pub) == "8520f0098930a754748b7ddcb43ef75a0dbf3a0d26381af4eba4a98eaa9b4e6a");
    CHECK(to_hex(bob_pub) == "de9edb7d7b7dc1b4d35b61c2ece435373f8343c85b78674dadfc7e146f882b4f");
    CHECK(to_hex(x25519(e, alice, bob_pub)) == "4a5d9d5ba4ce2de1728e3bf480350f25e07e21c947d19e3376f09b3c1e161742");
    CHECK(x25519(e, bob, alice_pub) == x25519(e, alice, bob_pub));
    // RFC 8032 7.1 TEST 1 (eddsa.rs:572-585)
    const auto seed = hex<32>("9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60");
    CHECK(to_hex(Ed25519Signature::derive_public_key(e, seed)) == "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a");
    // r2: the nonce commitment of sign(): RFC 6979 A.2.5 (P-256, SHA-256, "sample"): k -> r = x(k G) (ecdsa.rs:108-131)
    {
      const auto k = hex<32>("a6e3c57dd01abe90086538398355dd4c3b17aa873382b0f24d6129493d8aad60");
      const NonceCommit nc = batch_nonce_commit(e, FEC_P256, FEC_NONCE_ECDSA, std::vector<uint8_t>(k.begin(), k.end()));
      CHECK(nc.status[0] == 0 && to_hex(nc.r) == "efd48b2aacb6a8fd1140dd9cd45e81d69d2c877b56aaf991c34d0ea84eaf3716");
      CHECK(nc.R.size() == 33 && std::equal(nc.r.begin(), nc.r.end(), nc.R.begin() + 1));
    }
    CHECK(e.launch_count() > 0);
  } catch (const Error& err) {
    if (err.kind() == ErrorKind::UnsupportedOperation) {
      std::printf("no usable device: %s\n", err.what());
      return 77;
    }
    std::printf("unexpected error: %s\n", err.what());
    return 1;
  }
  if (failures) return 1;
  std::printf("test_traits: all checks passed\n");
  return 0;
}
