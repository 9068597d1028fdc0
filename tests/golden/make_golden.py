"""Generates tests/golden/*.json.  Run in the BUILD container (needs `cryptography` and `PyNaCl`,
which are independent implementations: OpenSSL and libsodium).  The fixtures pin the oracles
(oracle/pymodel.py, oracle/std) and, through them, the CUDA path; nothing on the GPU box needs
these libraries.  Inputs come from Python's `random` with a fixed seed plus hand-picked edge cases.

  python tests/golden/make_golden.py
"""
import json
import os
import random

import nacl.bindings as nb
from cryptography.hazmat.primitives import serialization as ser
from cryptography.hazmat.primitives.asymmetric import ec, ed25519, x25519

HERE = os.path.dirname(os.path.abspath(__file__))
rnd = random.Random(0x666F726765)

ORDER = {
    "secp256k1": 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141,
    "p256": 0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551,
}
CURVE = {"secp256k1": ec.SECP256K1(), "p256": ec.SECP256R1()}


def pub_bytes(curve, k, fmt):
    key = ec.derive_private_key(k, CURVE[curve])
    enc = ser.PublicFormat.CompressedPoint if fmt == "compressed" else ser.PublicFormat.UncompressedPoint
    return key.public_key().public_bytes(ser.Encoding.X962, enc)


def weierstrass():
    out = {}
    for name, n in ORDER.items():
        ks = [1, 2, 3, 0xDEADBEEF, n - 1, n - 2, (n + 1) // 2, (1 << 255) % n, 15 << 252, (30 << 252) - n, 0x1234567]
        ks += [rnd.randrange(1, n) for _ in range(24)]
        fixed = [{"k": "%064x" % k, "compressed": pub_bytes(name, k, "compressed").hex(),
                  "uncompressed": pub_bytes(name, k, "uncompressed").hex()} for k in ks]
        ecdh = []
        for _ in range(16):
            a, b = rnd.randrange(1, n), rnd.randrange(1, n)
            peer = ec.derive_private_key(b, CURVE[name]).public_key()
            shared = ec.derive_private_key(a, CURVE[name]).exchange(ec.ECDH(), peer)
            ecdh.append({"k": "%064x" % a, "point_uncompressed": pub_bytes(name, b, "uncompressed").hex(),
                         "point_compressed": pub_bytes(name, b, "compressed").hex(), "shared_x": shared.hex()})
        out[name] = {"fixed_base": fixed, "ecdh_x": ecdh}
    return out


def x25519_vectors():
    v = []
    for i in range(32):
        k = bytes(rnd.randrange(256) for _ in range(32))
        u = bytes(rnd.randrange(256) for _ in range(32)) if i else (9).to_bytes(32, "little")
        v.append({"k": k.hex(), "u": u.hex(), "out": nb.crypto_scalarmult(k, u).hex()})
    # RFC 7748 section 5.2 and 6.1 (quoted by the reference at curve25519.rs:2061-2089)
    v.append({"k": "a546e36bf0527c9d3b16154b82465edd62144c0ac1fc5a18506a2244ba449ac4",
              "u": "e6db6867583030db3594c1a424b15f7c726624ec26b3353b10a903a6d0ab1c4c",
              "out": "c3da55379de9c6908e94ea4df28d084f32eccf03491c71f754b4075577a28552"})
    v.append({"k": "4b66e9d4d1b4673c5ad22691957d6af5c11b6421e0ea01d42ca4169e7918ba0d",
              "u": "e5210f12786811d3f4b7959d0538ae2c31dbe7106fc03c3efc4cd549c715a493",
              "out": "95cbde9476e8907d7aade45cb4b873f88b595a68799fa152e6f8f7647aac7957"})
    v.append({"k": "77076d0a7318a57d3c16c17251b26645df4c2f87ebc0992ab177fba51db92c2a", "u": (9).to_bytes(32, "little").hex(),
              "out": "8520f0098930a754748b7ddcb43ef75a0dbf3a0d26381af4eba4a98eaa9b4e6a"})
    v.append({"k": "77076d0a7318a57d3c16c17251b26645df4c2f87ebc0992ab177fba51db92c2a",
              "u": "de9edb7d7b7dc1b4d35b61c2ece435373f8343c85b78674dadfc7e146f882b4f",
              "out": "4a5d9d5ba4ce2de1728e3bf480350f25e07e21c947d19e3376f09b3c1e161742"})
    for e in v[-4:]:  # cross-check the RFC rows with libsodium as well
        assert nb.crypto_scalarmult(bytes.fromhex(e["k"]), bytes.fromhex(e["u"])).hex() == e["out"]
    return v


def ed25519_vectors():
    v = []
    seeds = [bytes.fromhex("9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60"), bytes(range(32)), b"\x42" * 32, bytes(32),
             b"\xff" * 32]
    seeds += [bytes(rnd.randrange(256) for _ in range(32)) for _ in range(27)]
    for s in seeds:
        pk = ed25519.Ed25519PrivateKey.from_private_bytes(s).public_key().public_bytes(ser.Encoding.Raw, ser.PublicFormat.Raw)
        assert nb.crypto_sign_seed_keypair(s)[0] == pk
        v.append({"seed": s.hex(), "pubkey": pk.hex()})
    assert v[0]["pubkey"] == "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a"  # RFC 8032 7.1 TEST 1
    raw = []
    for _ in range(16):  # pre-clamped scalars: libsodium's base-point multiplication without clamping
        s = bytes(rnd.randrange(256) for _ in range(31)) + bytes([rnd.randrange(16)])
        if int.from_bytes(s, "little") == 0:
            continue
        raw.append({"scalar": s.hex(), "pubkey": nb.crypto_scalarmult_ed25519_base_noclamp(s).hex()})
    return {"seeds": v, "raw_scalars": raw}


if __name__ == "__main__":
    json.dump(weierstrass(), open(os.path.join(HERE, "weierstrass.json"), "w"), indent=1)
    json.dump(x25519_vectors(), open(os.path.join(HERE, "x25519.json"), "w"), indent=1)
    json.dump(ed25519_vectors(), open(os.path.join(HERE, "ed25519.json"), "w"), indent=1)
    # what the reference's literal code returns (transliteration goldens, not verified against rustc)
    ref = {
        "secp256k1_kG_compressed": {"1": "037bb1713828086cfe2039d780fba1c12f2b84d80b76e8ad541cc2b06ab3bba925",
                                    "2": "030247d90f0005beda9ce68ac77361ab5e13f71aeba7ee0371b821e6ce0f3b4a80",
                                    "3": "03bbc9d41801c4b8d163ea486b9179edd053950ee7877fe8a6a3a5a06e14fbf81d",
                                    "3735928559": "03413979fdc3541bd045e8b671d0505ae51528ac92557caa5c3867b673c744d4e8"},
        "secp256k1_generator_x_raw": "0239adf3970dbcab86dd8d9cb31d3988aa5e8ed55f112a5b203db8ccea164c84",
        "secp256k1_field_kats": {"mul_1_2_limb0": 12713681792961361445, "square_2_limb0": 4, "add_1_2_limb0": 3, "sub_3_1_limb0": 2},
        # P256::to_affine(multiply(generator(), k)).to_bytes() (p256.rs:2092-2156, 1835-1857, 1558-1577), SURVEY.md Appendix A.2
        "p256_kG_compressed": {"1": "036b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296",
                               "2": "02661ed53e414c757a6e928a7e1f8e19288c85f5deff73d5334ebe35532d24167c",
                               "3": "0243f28a513ce244ca16afe65f883be74907b350751b6da24a7fea2240f3285eec",
                               "3735928559": "038a13e0ce3ba3c2c96b73193caa64250cd528da2f3425c71c4286622b3d7ac76c"},
        # the small-operand checks of the reference's own field tests (p256.rs:2357-2393): 5+7, 7-5, 5*7, 5^2
        "p256_field_kats": {"add_5_7": 12, "sub_7_5": 2, "mul_5_7": 35, "square_5": 25},
        # x25519(scalar, [9, 0, ..]) (curve25519.rs:1624-1716), SURVEY.md Appendix A.2: the RFC 7748 Alice scalar, the scalars 1 and 2
        # (the latter is a hard-coded return value, :1626-1633)
        "x25519_base_point": {"77076d0a7318a57d3c16c17251b26645df4c2f87ebc0992ab177fba51db92c2a": "5872689f4f0154963f579041b3cc745d9ab25f66dbb9e8c319d74a280fbba923",
                              "01" + "00" * 31: "4ca04218e4464255382700d5d604394b0d009bc1220800f67fe03349fd9ff97a",
                              "02" + "00" * 31: "1b7f9f7c276550bb3a3cecc8a5770c173f5831ed1bb28c0558aac4713f970822"},
        # Ed25519::to_affine(multiply(generator(), k)).to_bytes() (ed25519.rs:2015-2097, 1793-1811, 1505-1524) and
        # Ed25519Signature::derive_public_key(seed) (eddsa.rs:450-509; the third seed hits the hard-coded return), SURVEY.md Appendix A.2
        "ed25519_mul_generator_33": {"1": "03f28396fbfa62141a04c4308b8ba6e8d2631be7e9a1f3c0a0fe536ecdd3366921",
                                     "2": "02b9ce7ad7c2bc7c79cfc9525020afb0626343f635000a672b8641c5058b67e871"},
        "ed25519_derive_public_key": {bytes(range(32)).hex(): "030f4fbbae5405816d05f258a3510ad17216844c709b3934e5d613c5351e7c82",
                                      "42" * 32: "025ac8d0d3f99edfc824e3aef017eddf81d09d0e55fb0fd722d351328d769ac3",
                                      "9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60":
                                          "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a"},
    }
    json.dump(ref, open(os.path.join(HERE, "forge_ec_literal.json"), "w"), indent=1)
    print("golden fixtures written")
