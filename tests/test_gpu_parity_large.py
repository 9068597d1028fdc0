"""GPU parity at configuration size: every entry point of the C ABI on >= 2^16 lanes against oracle/std/oracle_std.c (the
threaded C restatement of the specified behaviour -- not the GPU, not a few hundred lanes of the Python model), byte for byte.
BASELINE.json configs[0] is exactly the first case: 2^16 Philox scalars with the 8 edge lanes.  Expected values
never come from the GPU; the MSM expectations are the oracle's own products added one by one."""
import hashlib
import hmac

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 1 << 16


def _be(v):
    return v.to_bytes(32, "big")


def _sprinkle(pm, curve, sc, pts=None):
    """edge scalars (next to the group order, out of range, the exceptional cases of the window recodings) and rejected /
    identity points on a few lanes in the middle of the batch"""
    c = pm.CURVES[curve]
    edge = [c.n, c.n + 1, 2**256 - 1, (30 << 252) - c.n, 15 << 252, c.n - 15, c.n - 30, 16, 17, (1 << 128) - 1, 1 << 128, (1 << 255) - 1]
    for i, v in enumerate(edge):
        sc[1000 + 37 * i] = np.frombuffer(_be(v % 2**256), dtype=np.uint8)
    if pts is not None:
        bad = [_be(1) + _be(1), _be(c.p) + _be(c.gy), b"\xff" * 64, bytes(64), _be(c.gx) + _be(c.gy)]
        for j, b in enumerate(bad):
            pts[2000 + 41 * j] = np.frombuffer(b, dtype=np.uint8)


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("out_fmt", [0, 4])
def test_fixed_base_2p16_matches_c_oracle(engine, ostd, pm, curve, out_fmt):
    """BASELINE.json configs[0]: k*G over 2^16 Philox scalars, compressed (and hybrid) encoding"""
    sc = pm.random_scalars_be(curve, N, stream=1)
    if out_fmt == 4:
        _sprinkle(pm, curve, sc)
    out, st = engine.batch_mul(curve, sc, None, out_fmt=out_fmt)
    exp, est = ostd.mul(curve, sc.tobytes(), None, 0, out_fmt)
    assert st.tobytes() == est
    assert out.tobytes() == exp
    assert est[:8] == bytes([3, 0, 0, 0, 0, 0, 0, 0])  # lane 0 is k = 0


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("points_fmt", [0, 1])
def test_variable_base_2p16_matches_c_oracle(engine, fec, ostd, pm, curve, points_fmt):
    """configs[2] / the P-256 k*P path on 2^16 (k, P) pairs, X || Y and SEC1 compressed input"""
    sc = pm.random_scalars_be(curve, N, stream=3)
    r = pm.random_scalars_be(curve, N, stream=3, sub=1, edge=False)
    xy, _ = ostd.mul(curve, r.tobytes(), None, 0, 2)                       # the points come from the oracle, not from the GPU
    pts = np.frombuffer(xy, dtype=np.uint8).reshape(N, 64).copy()
    _sprinkle(pm, curve, sc, pts if points_fmt == 0 else None)
    if points_fmt == 1:
        comp, _ = ostd.mul(curve, r.tobytes(), None, 0, 0)
        pts = np.frombuffer(comp, dtype=np.uint8).reshape(N, 33).copy()
        pts[2000, 0] = 4          # bad prefix
        pts[2041, 0] = 0          # the identity
        pts[2082, 1:] = 0xFF      # x >= p
    out, st = engine.batch_mul(curve, sc, pts, points_fmt=points_fmt, out_fmt=1)
    exp, est = ostd.mul(curve, sc.tobytes(), pts.tobytes(), points_fmt, 1)
    assert st.tobytes() == est
    assert out.tobytes() == exp
    assert set(est) == {0, 1, 2, 3}


def test_x25519_2p16_matches_c_oracle(engine, ostd, pm):
    """configs[3]: X25519 on 2^16 (scalar, u) pairs incl. low-order and non-canonical u, and on the base point"""
    ks = pm.philox_bytes32(N, stream=4)
    us = pm.philox_bytes32(N, stream=4, sub=1).copy()
    low = [0, 1, 2**255 - 19, 2**255 - 18, 2**255 - 1, 2**256 - 1, 325606250916557431795983626356110631294008115727848805560023387167927233504,
           39382357235489614581723060781553021112529911719440698176882885853963445705823, 2**255 - 20]
    for i, v in enumerate(low):
        us[500 + i] = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8)
    assert engine.batch_x25519(ks, us).tobytes() == ostd.x25519(ks.tobytes(), us.tobytes())
    assert engine.batch_x25519(ks, None).tobytes() == ostd.x25519(ks.tobytes(), None)


def test_ed25519_2p16_matches_c_oracle(engine, ostd, pm):
    """configs[4]: Ed25519 key generation from 2^16 seeds (SHA-512 on the device) and from raw scalars"""
    seeds = pm.philox_bytes32(N, stream=5)
    assert engine.batch_ed25519_pubkey(seeds, True).tobytes() == ostd.ed(seeds.tobytes(), True)
    sc = pm.philox_bytes32(N, stream=5, sub=1).copy()
    L = 2**252 + 27742317777372353535851937790883648493
    for i, v in enumerate([0, 1, 2, L, L - 1, L + 1, 2**255, 2**256 - 1, 8 * L]):
        sc[300 + i] = np.frombuffer((v % 2**256).to_bytes(32, "little"), dtype=np.uint8)
    assert engine.batch_ed25519_pubkey(sc, False).tobytes() == ostd.ed(sc.tobytes(), False)


def _hkdf(secret, info, out_len):
    prk = hmac.new(bytes(32), secret, hashlib.sha256).digest()
    okm, t, i = b"", b"", 1
    while len(okm) < out_len:
        t = hmac.new(prk, t + info + bytes([i]), hashlib.sha256).digest()
        okm += t
        i += 1
    return okm[:out_len]


@pytest.mark.parametrize("curve", [0, 1])
def test_ecdh_kdf_2p16_matches_c_oracle(engine, fec, ostd, pm, curve):
    """8f.2 on 2^16 pairs: x(k*P) from the C oracle, HKDF-SHA-256 (secp256k1.rs:1846-1883) from hashlib / the P-256 XOR
    placeholder (p256.rs:2314-2344) restated here; rejected lanes give zero records"""
    a = pm.random_scalars_be(curve, N, stream=6)
    b = pm.random_scalars_be(curve, N, stream=6, sub=1, edge=False)
    xy, _ = ostd.mul(curve, b.tobytes(), None, 0, 2)
    B = np.frombuffer(xy, dtype=np.uint8).reshape(N, 64).copy()
    _sprinkle(pm, curve, a, B)
    xs, est = ostd.mul(curve, a.tobytes(), B.tobytes(), 0, 3)
    out, st = engine.batch_ecdh(curve, a, B)
    assert st.tobytes() == est and out.tobytes() == xs
    info, out_len = b"forge-ec ecdh 2^16", 40
    kdf = fec.KDF_HKDF_SHA256 if curve == 0 else fec.KDF_XOR_INFO
    got, st = engine.batch_ecdh(curve, a, B, kdf=kdf, info=info, out_len=out_len)
    exp = bytearray(N * out_len)
    for i in range(N):
        if est[i] == 0:
            sec = xs[32 * i : 32 * i + 32]
            exp[out_len * i : out_len * (i + 1)] = _hkdf(sec, info, out_len) if curve == 0 else pm.kdf_xor_info(sec, info, out_len)
    assert st.tobytes() == est and got.tobytes() == bytes(exp)


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("n", [5000, 1 << 17])
def test_msm_matches_c_oracle(engine, fec, ostd, pm, curve, n):
    """8f.3: sum k_i P_i through the products + reduction path (n = 5000) and the bucket method (n = 2^17) against the
    oracle's products added one by one (ostd_msm); nothing in the expectation comes from the GPU"""
    c = pm.CURVES[curve]
    k = pm.random_scalars_be(curve, n, stream=9)
    r = pm.random_scalars_be(curve, n, stream=9, sub=1, edge=False)
    xy, _ = ostd.mul(curve, r.tobytes(), None, 0, 2)
    pts = np.frombuffer(xy, dtype=np.uint8).reshape(n, 64).copy()
    pts[77] = 0                                                            # the identity as an input point adds nothing
    prods = ostd.mul(curve, k.tobytes(), pts.tobytes(), 0, 2)              # the oracle's products, computed once
    assert prods[1][77] == 3
    exp = ostd.msm_products(curve, *prods, 0)
    assert exp[1] == 0 and engine.msm(curve, k, pts) == exp
    assert engine.msm(curve, k, pts, out_fmt=1) == ostd.msm_products(curve, *prods, 1)
    # SEC1 compressed input of the same points (prefix 0x00 = the identity): the same sum
    comp = np.empty((n, 33), dtype=np.uint8)
    comp[:, 0] = 2 | (pts[:, 63] & 1)
    comp[:, 1:] = pts[:, :32]
    comp[77] = 0
    assert engine.msm(curve, k, comp, points_fmt=fec.PT_SEC1_COMPRESSED) == exp
    if n <= 5000:  # the oracle on the decompression and on the rejections themselves
        assert ostd.msm(curve, k.tobytes(), comp.tobytes(), 1, 0) == exp
    bad = k.copy()
    bad[n // 2] = np.frombuffer(_be(c.n + 5), dtype=np.uint8)
    off = pts.copy()
    off[n // 3, 63] ^= 1
    if n <= 5000:
        assert ostd.msm(curve, bad.tobytes(), pts.tobytes(), 0, 0) == (bytes(33), 1)
        assert ostd.msm(curve, k.tobytes(), off.tobytes(), 0, 0) == (bytes(33), 2)
    assert engine.msm(curve, bad, pts) == (bytes(33), 1)
    assert engine.msm(curve, k, off) == (bytes(33), 2)
    # repeated scalars (one bucket per window holds everything: the bucket method falls back) and one repeated point
    same = np.tile(np.frombuffer(_be(0x1234567 << 200 | 0xABCDEF), dtype=np.uint8), n).reshape(n, 32)
    assert engine.msm(curve, same, pts) == ostd.msm(curve, same.tobytes(), pts.tobytes(), 0, 0)
    onep = np.tile(pts[5], n).reshape(n, 64)
    ksum = sum(int.from_bytes(k[i].tobytes(), "big") for i in range(n)) % c.n
    assert engine.msm(curve, k, onep) == (ostd.mul(curve, _be(ksum), pts[5].tobytes(), 0, 0)[0], 0)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_nonce_commit_2p16_matches_c_oracle(engine, ostd, pm, mode):
    """8f.4 on 2^16 nonces: ECDSA (secp256k1), BIP-340, EdDSA (64-byte hash values)"""
    if mode == 2:
        nonces = pm.philox_bytes32(2 * N, stream=13).reshape(N, 64)
    else:
        nonces = pm.random_scalars_be(0, N, stream=12)
        _sprinkle(pm, 0, nonces)
    R, r, st = engine.batch_nonce_commit(0, mode, nonces)
    eR, er, est = ostd.nonce_commit(0, mode, nonces.tobytes())
    assert st.tobytes() == est and r.tobytes() == er and R.tobytes() == eR


@pytest.mark.parametrize("curve", [0, 1])
def test_point_decode_2p16_matches_c_oracle(engine, ostd, pm, curve):
    """8f.1 on its own: 2^16 SEC1 compressed keys -> validated affine X || Y"""
    r = pm.random_scalars_be(curve, N, stream=14, edge=False)
    comp = np.frombuffer(ostd.mul(curve, r.tobytes(), None, 0, 0)[0], dtype=np.uint8).reshape(N, 33).copy()
    comp[2000, 0] = 4; comp[2041, 0] = 0; comp[2082, 1:] = 0xFF; comp[3000:3100, 32] ^= 1
    out, st = engine.batch_point_decode(curve, comp)
    exp, est = ostd.point_decode(curve, comp.tobytes(), 1, 2)
    assert st.tobytes() == est and out.tobytes() == exp and {0, 2, 3} <= set(est)
