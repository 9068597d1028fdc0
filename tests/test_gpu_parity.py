"""GPU parity: the CUDA path, called through the C ABI, against the oracle on identical inputs.
Bit-exact for every byte (integer work; no tolerance)."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _be(v):
    return v.to_bytes(32, "big")


def _edge_scalars(c):
    return [0, 1, 2, 3, c.n - 1, c.n - 2, (c.n + 1) // 2, (1 << 255) % c.n, 0xDEADBEEF, c.n, c.n + 1, 2**256 - 1,
            (30 << 252) - c.n, 15 << 252, c.n - 15, c.n - 30, 16, 17, (1 << 128) - 1]


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("out_fmt", [0, 1, 2, 3, 4])  # compressed, uncompressed, X || Y, X, hybrid (PointFormat::Hybrid)
def test_fixed_base_matches_model(engine, pm, curve, out_fmt):
    c = pm.CURVES[curve]
    n = 300
    sc = pm.random_scalars_be(curve, n, stream=1)
    edge = _edge_scalars(c)
    for i, v in enumerate(edge):
        sc[8 + i] = np.frombuffer(_be(v), dtype=np.uint8)
    out, st = engine.batch_mul(curve, sc, None, out_fmt=out_fmt)
    exp, est = pm.batch_mul(curve, sc.tobytes(), None, n, out_fmt)
    assert st.tobytes() == est
    assert out.tobytes() == exp


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("points_fmt", [0, 1])
def test_variable_base_matches_model(engine, pm, curve, points_fmt):
    c = pm.CURVES[curve]
    n = 200
    sc = pm.random_scalars_be(curve, n, stream=3)
    edge = _edge_scalars(c)
    for i, v in enumerate(edge):
        sc[8 + i] = np.frombuffer(_be(v), dtype=np.uint8)
    r = pm.random_scalars_be(curve, n, stream=3, sub=1, edge=False)
    pts_xy, _ = pm.batch_mul(curve, r.tobytes(), None, n, pm.OUT_AFFINE_XY_BE)
    pts_xy = bytearray(pts_xy)
    # lanes 40..59: P = G (small multiples meet the table); lane 60..: invalid points
    g = _be(c.gx) + _be(c.gy)
    for i in range(40, 60):
        pts_xy[64 * i : 64 * i + 64] = g
    bad = [_be(1) + _be(1), _be(c.p) + _be(c.gy), _be(c.gx) + _be(c.p + 1 if curve == 0 else c.p), bytes(64), b"\xff" * 64]
    for j, b in enumerate(bad):
        pts_xy[64 * (60 + j) : 64 * (61 + j)] = b
    sc[64] = 0  # k = 0 on an invalid point: reported as an invalid point
    pts_xy[64 * 64 : 64 * 65] = _be(2) + _be(2)
    if points_fmt == 0:
        pts = bytes(pts_xy)
    else:
        recs = []
        for i in range(n):
            x = int.from_bytes(pts_xy[64 * i : 64 * i + 32], "big")
            y = int.from_bytes(pts_xy[64 * i + 32 : 64 * i + 64], "big")
            recs.append(bytes([2 | (y & 1)]) + _be(x))
        recs[60] = b"\x04" + recs[60][1:]          # bad prefix
        recs[61] = b"\x02" + _be(c.p)              # x out of range
        recs[62] = b"\x02" + _be(5 if curve == 0 else 1)  # no square root (checked below by the model)
        recs[63] = bytes(33)
        recs[64] = b"\x05" + recs[64][1:]
        pts = b"".join(recs)
    out, st = engine.batch_mul(curve, sc, np.frombuffer(pts, dtype=np.uint8), points_fmt=points_fmt)
    exp, est = pm.batch_mul(curve, sc.tobytes(), pts, n, 0, points_fmt)
    assert st.tobytes() == est
    assert out.tobytes() == exp
    # lane 63 is the identity as an input point (all-zero X || Y resp. prefix 0x00): k * O = O, status 3
    assert est[60] == 2 and est[63] == 3 and est[64] == 2


def test_golden_weierstrass(engine, fec, pm):
    """SURVEY Appendix A.1 (independently produced with OpenSSL via `cryptography`)."""
    k1 = {1: "0279be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798",
          2: "02c6047f9441ed7d6d3045406e95c07cd85c778e4b8cef3ca7abac09b95c709ee5",
          3: "02f9308a019258c31049344f85f89d5229b531c845836f99b08601f113bce036f9",
          0xDEADBEEF: "0276d2fdf1302d1fa9556f4df94ec84cefba6d482e54f47c6c2a238c1baa560f0e"}
    p256 = {1: "036b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296",
            2: "037cf27b188d034f7e8a52380304b51ac3c08969e277f21b35a60b48fc47669978",
            3: "025ecbe4d1a6330a44c8f7ef951d4bf165e6c6b721efada985fb41661bc6e7fd6c",
            0xDEADBEEF: "02b487d183dc4806058eb31a29bedefd7bcca987b77a381a3684871d8449c18394"}
    for curve, table in ((0, k1), (1, p256)):
        ks = list(table)
        sc = np.frombuffer(b"".join(_be(k) for k in ks), dtype=np.uint8)
        out, st = engine.batch_mul(curve, sc, None)
        assert [out[i].tobytes().hex() for i in range(len(ks))] == [table[k] for k in ks]
        assert not st.any()
    # ECDH x-coordinate: k = 0xdeadbeef, P = 0x1234567 G
    P = fec.Secp256k1.multiply(None, fec.Scalar.from_int(0x1234567), engine)
    S = fec.Secp256k1.multiply(P, fec.Scalar.from_int(0xDEADBEEF), engine)
    assert S.x().hex() == "9f1817c02d82908ac12b2427a13b07c9a21fb8b797296eb8af47f4f91f1dac0f"
    # identity
    Z = fec.P256.multiply(None, fec.Scalar.from_int(0), engine)
    assert Z.is_identity() and Z.to_bytes() == bytes(33)


def test_x25519_matches_model_and_rfc7748(engine, fec, pm):
    alice = bytes.fromhex("77076d0a7318a57d3c16c17251b26645df4c2f87ebc0992ab177fba51db92c2a")
    bob_pub = bytes.fromhex("de9edb7d7b7dc1b4d35b61c2ece435373f8343c85b78674dadfc7e146f882b4f")
    assert fec.x25519(alice, (9).to_bytes(32, "little"), engine).hex() == \
        "8520f0098930a754748b7ddcb43ef75a0dbf3a0d26381af4eba4a98eaa9b4e6a"
    assert fec.x25519(alice, bob_pub, engine).hex() == "4a5d9d5ba4ce2de1728e3bf480350f25e07e21c947d19e3376f09b3c1e161742"
    n = 600
    ks = pm.philox_bytes32(n, stream=4)
    us = pm.philox_bytes32(n, stream=4, sub=1).copy()
    special = [bytes(32), (1).to_bytes(32, "little"), (2**255 - 19).to_bytes(32, "little"), (2**255 - 1).to_bytes(32, "little"),
               b"\xff" * 32, (2**255 - 20).to_bytes(32, "little"), (9).to_bytes(32, "little"),
               # RFC 7748 section 5.2 second vector's u (bit 255 set: must be masked)
               bytes.fromhex("e5210f12786811d3f4b7959d0538ae2c31dbe7106fc03c3efc4cd549c715a493")]
    for i, s in enumerate(special):
        us[i] = np.frombuffer(s, dtype=np.uint8)
    out = engine.batch_x25519(ks, us)
    for i in range(n):
        assert out[i].tobytes() == pm.x25519(ks[i].tobytes(), us[i].tobytes()), i
    # u = NULL (the base point 9) takes the fixed-base route on the Edwards form of the curve: it must agree with the ladder
    # run on an explicit u = 9, with the model, and on the extreme scalars (all zero / all one bits before clamping)
    kb = ks.copy()
    kb[0] = 0
    kb[1] = 255
    base = engine.batch_x25519(kb, None)
    nine = np.tile(np.frombuffer(pm.X25519_BASE, dtype=np.uint8), (n, 1))
    assert base.tobytes() == engine.batch_x25519(kb, nine).tobytes()
    for i in range(50):
        assert base[i].tobytes() == pm.x25519(kb[i].tobytes(), pm.X25519_BASE)
    assert fec.batch_x25519([alice], None, engine)[0].hex() == "8520f0098930a754748b7ddcb43ef75a0dbf3a0d26381af4eba4a98eaa9b4e6a"
    # RFC 7748 5.2 vector 1
    k = bytes.fromhex("a546e36bf0527c9d3b16154b82465edd62144c0ac1fc5a18506a2244ba449ac4")
    u = bytes.fromhex("e6db6867583030db3594c1a424b15f7c726624ec26b3353b10a903a6d0ab1c4c")
    assert fec.x25519(k, u, engine).hex() == "c3da55379de9c6908e94ea4df28d084f32eccf03491c71f754b4075577a28552"


def test_x25519_rfc7748_iterated_vector_on_gpu(engine):
    """RFC 7748 section 5.2: k = u = 9, (k, u) <- (X25519(k, u), k); the RFC's values after 1 and 1000 iterations, each step
    one single-lane call of the ladder kernel."""
    k = u = (9).to_bytes(32, "little")
    for i in range(1, 1001):
        r = engine.batch_x25519(np.frombuffer(k, dtype=np.uint8), np.frombuffer(u, dtype=np.uint8), n=1)[0].tobytes()
        k, u = r, k
        if i == 1:
            assert k.hex() == "422c8e7a6227d7bca1350b3e2bb7279f7897b87bb6854b783c60e80311ae3079"
    assert k.hex() == "684cf59ba83309552800ef566f2f4d3c1c3887c49360e3875f2eb94d99532c51"


def test_ed25519_pubkey_matches_model_and_rfc8032(engine, fec, pm):
    seed = bytes.fromhex("9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60")
    assert fec.Ed25519Signature.derive_public_key(seed, engine).hex() == \
        "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a"
    assert fec.Ed25519Signature.derive_public_key(bytes(range(32)), engine).hex() == \
        "03a107bff3ce10be1d70dd18e74bc09967e4d6309ba50d5f1ddc8664125531b8"
    assert fec.Ed25519Signature.derive_public_key(b"\x42" * 32, engine).hex() == \
        "2152f8d19b791d24453242e15f2eab6cb7cffa7b6a5ed30097960e069881db12"
    n = 300
    seeds = pm.philox_bytes32(n, stream=5)
    out = engine.batch_ed25519_pubkey(seeds, True)
    for i in range(n):
        assert out[i].tobytes() == pm.ed25519_pubkey(seeds[i].tobytes()), i
    raw = pm.philox_bytes32(n, stream=5, sub=1).copy()
    special = [0, 1, 2, pm.ED_L, pm.ED_L - 1, pm.ED_L + 1, 2**256 - 1, 2**255, 2**255 - 2, 8]
    for i, v in enumerate(special):
        raw[i] = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8)
    out = engine.batch_ed25519_pubkey(raw, False)
    for i in range(n):
        assert out[i].tobytes() == pm.ed25519_pubkey(raw[i].tobytes(), False), i


@pytest.mark.parametrize("prime", [0, 1, 2])
@pytest.mark.parametrize("op", [0, 1, 2])
def test_field_bench_matches_model(engine, pm, prime, op):
    p = pm.PRIMES[prime]
    n = 2048
    a = pm.random_field_le(prime, n, stream=2).copy()
    b = pm.random_field_le(prime, n, stream=2, sub=1).copy()
    edge = [0, 1, 2, p - 1, p - 2, (p + 1) // 2, (1 << 255) % p, (1 << 224) % p, (2**256 - 1) % p]
    for i, v in enumerate(edge):
        a[i] = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8)
        b[(i * 3) % len(edge)] = np.frombuffer(edge[(i * 5) % len(edge)].to_bytes(32, "little"), dtype=np.uint8)
    for iters in (1, 7):
        out, sec = engine.field_bench(prime, op, a, b, iters)
        assert sec > 0
        for i in range(n):
            ai = int.from_bytes(a[i].tobytes(), "little")
            bi = int.from_bytes(b[i].tobytes(), "little")
            assert int.from_bytes(out[i].tobytes(), "little") == pm.field_bench_model(prime, op, iters, ai, bi), (i, iters)


def test_p256_large_operand_field_kats_on_gpu(engine, fec, pm):
    """the reference's own large-operand field values (p256.rs:2454-2480: Gx^2, Gx^3, 3 Gx, Gy^2 as u64 limbs) through the GPU
    field kernels -- the same constants tests/test_oracle.py pins both oracles to"""
    val = lambda l: sum(v << (64 * i) for i, v in enumerate(l))
    X2 = val([12074202155401100, 3726334282074508753, 9331909631644438744, 11022199779588240050])
    X3 = val([6985818112209442057, 5293983511093485517, 13285487596276262425, 4350650246863171228])
    TX = val([15988812018543642563, 7280764249650076386, 16876875322344915671, 4703857913423513302])
    Y2 = val([13753198298469232017, 5299206390010787296, 9373276401007028734, 6187767046927055789])
    c = pm.CURVES[1]
    le = lambda v: np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8)
    a = np.stack([le(c.gx), le(c.gx), le(3), le(c.gy), le(X2)])
    b = np.stack([le(c.gx), le(c.gx), le(c.gx), le(c.gy), le(c.gx)])
    mul1, _ = engine.field_bench(fec.PRIME_P256, fec.FIELD_MUL, a, b, 1)
    mul2, _ = engine.field_bench(fec.PRIME_P256, fec.FIELD_MUL, a, b, 2)
    sqr1, _ = engine.field_bench(fec.PRIME_P256, fec.FIELD_SQR, a, b, 1)
    g = lambda o, i: int.from_bytes(o[i].tobytes(), "little")
    assert g(mul1, 0) == X2 and g(sqr1, 0) == X2 and g(mul2, 1) == X3 and g(mul1, 4) == X3
    assert g(mul1, 2) == TX and g(mul1, 3) == Y2 and g(sqr1, 3) == Y2
    assert (X3 - TX + c.b) % c.p == Y2


def test_ragged_and_empty(engine, pm):
    """n = 0, n = 1, n not a multiple of any tile size; outputs beyond n untouched."""
    out, st = engine.batch_mul(0, np.empty(0, dtype=np.uint8), None, n=0)
    assert out.shape == (0, 33) and st.shape == (0,)
    assert engine.batch_x25519(np.empty(0, dtype=np.uint8), None, n=0).shape == (0, 32)
    for n in (1, 31, 129, 257, 1000):
        sc = pm.random_scalars_be(0, n, stream=6, edge=False)
        out, st = engine.batch_mul(0, sc, None)
        exp, est = pm.batch_mul(0, sc[: min(n, 40)].tobytes(), None, min(n, 40), 0)
        assert out[: min(n, 40)].tobytes() == exp and not st.any()


def test_synth_inputs_match_philox_model(engine, pm):
    import torch

    n = 5000
    buf = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    for mode, stream_id, sub, first, edge in ((0, 3, 0, 0, False), (0, 4, 1, 12345, False), (1, 3, 0, 0, True), (2, 5, 0, 0, True),
                                              (1, 3, 2, 4, True)):
        engine.synth_dev(mode, stream_id, sub, first, n, buf, edge=edge, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = buf.cpu().numpy().reshape(n, 32)
        if mode == 0:
            exp = pm.philox_bytes32(n, stream_id, first, sub)
        else:
            exp = pm.random_scalars_be(mode - 1, n, stream_id, first, edge, sub)
        assert np.array_equal(got, exp), (mode, stream_id, sub, first)


@pytest.mark.parametrize("n", [1 << 14, (1 << 22) + 4099])
def test_device_pointer_api_and_properties(engine, fec, pm, n):
    """Size-independent properties at small and at full launch size (the larger case spans two scratch chunks of the
    three-kernel k*P path and ends in a ragged tile): ECDH commutes on both curves, fixed-base and variable-base agree
    on P = G, the first lanes match the model, X25519 exchanges agree and its base-point route equals the ladder."""
    import torch

    st = torch.cuda.current_stream().cuda_stream
    a = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    b = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    engine.synth_dev(1, 3, 0, 0, n, a, edge=False, stream=st)
    engine.synth_dev(1, 3, 1, 0, n, b, edge=False, stream=st)
    A = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    B = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    engine.batch_mul_dev(0, a, None, n, A, out_fmt=2, stream=st)
    engine.batch_mul_dev(0, b, None, n, B, out_fmt=2, stream=st)
    s1 = torch.empty(n * 33, dtype=torch.uint8, device="cuda")
    s2 = torch.empty(n * 33, dtype=torch.uint8, device="cuda")
    st1 = torch.empty(n, dtype=torch.uint8, device="cuda")
    engine.batch_mul_dev(0, a, B, n, s1, st1, stream=st)
    engine.batch_mul_dev(0, b, A, n, s2, None, stream=st)
    torch.cuda.synchronize()
    assert torch.equal(s1, s2)
    assert int(st1.sum()) == 0
    # first lanes against the model
    exp, _ = pm.batch_mul(0, a[: 32 * 16].cpu().numpy().tobytes(), B[: 64 * 16].cpu().numpy().tobytes(), 16, 0)
    assert s1[: 33 * 16].cpu().numpy().tobytes() == exp
    tail = n - 16   # and the last lanes (ragged tile, second chunk)
    exp, _ = pm.batch_mul(0, a[32 * tail:].cpu().numpy().tobytes(), B[64 * tail:].cpu().numpy().tobytes(), 16, 0)
    assert s1[33 * tail:].cpu().numpy().tobytes() == exp
    # k*G through the fixed-base table == k*P with P = G through the variable-base kernels
    G = torch.from_numpy(np.frombuffer((_be(pm.CURVES[0].gx) + _be(pm.CURVES[0].gy)) * 4096, dtype=np.uint8).copy()).cuda()
    engine.batch_mul_dev(0, a, None, 4096, s1, stream=st)
    engine.batch_mul_dev(0, a, G, 4096, s2, stream=st)
    torch.cuda.synchronize()
    assert torch.equal(s1[: 33 * 4096], s2[: 33 * 4096])
    # P-256: the same exchange, at the full size too (two scratch chunks and a ragged tile of its three-kernel path)
    m = n
    engine.synth_dev(2, 5, 0, 0, m, a, edge=False, stream=st)
    engine.synth_dev(2, 5, 1, 0, m, b, edge=False, stream=st)
    engine.batch_mul_dev(1, a, None, m, A, out_fmt=2, stream=st)
    engine.batch_mul_dev(1, b, None, m, B, out_fmt=2, stream=st)
    engine.batch_mul_dev(1, a, B, m, s1, st1, stream=st)
    engine.batch_mul_dev(1, b, A, m, s2, None, stream=st)
    torch.cuda.synchronize()
    assert torch.equal(s1[: 33 * m], s2[: 33 * m]) and int(st1[:m].sum()) == 0
    exp, _ = pm.batch_mul(1, a[32 * tail : 32 * n].cpu().numpy().tobytes(), B[64 * tail : 64 * n].cpu().numpy().tobytes(), 16, 0)
    assert s1[33 * tail : 33 * n].cpu().numpy().tobytes() == exp
    G = torch.from_numpy(np.frombuffer((_be(pm.CURVES[1].gx) + _be(pm.CURVES[1].gy)) * 4096, dtype=np.uint8).copy()).cuda()
    engine.batch_mul_dev(1, a, None, 4096, s1, stream=st)
    engine.batch_mul_dev(1, a, G, 4096, s2, stream=st)
    torch.cuda.synchronize()
    assert torch.equal(s1[: 33 * 4096], s2[: 33 * 4096])
    # X25519: both sides derive the same secret
    ka = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    kb = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    engine.synth_dev(0, 4, 0, 0, n, ka, stream=st)
    engine.synth_dev(0, 4, 1, 0, n, kb, stream=st)
    pa, pb, sa, sb = (torch.empty(n * 32, dtype=torch.uint8, device="cuda") for _ in range(4))
    engine.batch_x25519_dev(ka, None, n, pa, stream=st)
    engine.batch_x25519_dev(kb, None, n, pb, stream=st)
    engine.batch_x25519_dev(ka, pb, n, sa, stream=st)
    engine.batch_x25519_dev(kb, pa, n, sb, stream=st)
    torch.cuda.synchronize()
    assert torch.equal(sa, sb)
    nine = torch.zeros(n * 32, dtype=torch.uint8, device="cuda")
    nine[::32] = 9
    engine.batch_x25519_dev(ka, nine, n, sb, stream=st)   # the ladder on an explicit u = 9 vs the fixed-base route
    torch.cuda.synchronize()
    assert torch.equal(pa, sb)
    assert engine.launch_count > 0


def test_host_buffer_pipeline_spans_chunks(engine, fec, pm):
    """The host-buffer entry points move data in 2^19-lane chunks over three streams: a call of 2^20 + 4133 lanes (three
    chunks, a ragged last one) must equal the device-pointer call on the same inputs, for k*G, k*P and Ed25519 keygen."""
    import torch

    n = (1 << 20) + 4133
    st = torch.cuda.current_stream().cuda_stream
    k = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    r = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    engine.synth_dev(1, 3, 0, 0, n, k, edge=True, stream=st)
    engine.synth_dev(1, 3, 1, 0, n, r, edge=False, stream=st)
    P = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    o_dev = torch.empty(n * 33, dtype=torch.uint8, device="cuda")
    s_dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    engine.batch_mul_dev(0, r, None, n, P, out_fmt=2, stream=st)
    engine.batch_mul_dev(0, k, P, n, o_dev, s_dev, stream=st)
    torch.cuda.synchronize()
    hk, hr, hP = k.cpu().numpy(), r.cpu().numpy(), P.cpu().numpy()
    out, stat = engine.batch_mul(0, hr, None, out_fmt=fec.OUT_AFFINE_XY_BE)
    assert out.tobytes() == hP.tobytes() and not stat.any()
    out, stat = engine.batch_mul(0, hk, hP)
    assert out.tobytes() == o_dev.cpu().numpy().tobytes() and stat.tobytes() == s_dev.cpu().numpy().tobytes()
    exp, est = pm.batch_mul(0, hk[-32 * 8:].tobytes(), hP[-64 * 8:].tobytes(), 8, 0)   # the ragged tail against the model
    assert out[-8:].tobytes() == exp and stat[-8:].tobytes() == est
    pk_dev = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    engine.batch_ed25519_pubkey_dev(k, True, n, pk_dev, stream=st)
    torch.cuda.synchronize()
    assert engine.batch_ed25519_pubkey(hk, True).tobytes() == pk_dev.cpu().numpy().tobytes()


def test_every_batch_size_regime_gives_the_same_lanes(engine, fec, pm):
    """The deferred-inversion kernels run in three regimes -- one tile per warp with a warp-level inversion (small grids), a
    block-level inversion shared by four warps (full grid, few tiles per warp), and 8-16 deferred tiles per warp-level
    inversion (large batches).  Lanes are independent, so a call on the first n lanes must return the first n records of the
    call on all of them, for every n; the head of the full result is checked against the model."""
    import torch

    st = torch.cuda.current_stream().cuda_stream
    nmax = 1_100_009
    sizes = [1, 37, 4101, 70_001, 300_007, 454_656, 454_688, 460_003, 757_760 + 33]
    k = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
    u = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
    k2 = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
    engine.synth_dev(1, 7, 0, 0, nmax, k, edge=True, stream=st)
    engine.synth_dev(2, 7, 0, 0, nmax, k2, edge=True, stream=st)
    engine.synth_dev(0, 7, 1, 0, nmax, u, stream=st)
    runs = {
        "k1_kG": (33, lambda n, o: engine.batch_mul_dev(0, k, None, n, o, stream=st)),
        "p256_kG": (33, lambda n, o: engine.batch_mul_dev(1, k2, None, n, o, stream=st)),
        "x25519": (32, lambda n, o: engine.batch_x25519_dev(k, u, n, o, stream=st)),
        "x25519_base": (32, lambda n, o: engine.batch_x25519_dev(k, None, n, o, stream=st)),
        "ed25519": (32, lambda n, o: engine.batch_ed25519_pubkey_dev(k, True, n, o, stream=st)),
    }
    for name, (rec, fn) in runs.items():
        full = torch.empty(nmax * rec, dtype=torch.uint8, device="cuda")
        fn(nmax, full)
        part = torch.empty(nmax * rec, dtype=torch.uint8, device="cuda")
        for n in sizes:
            if name == "x25519" and n > 310_000:
                continue   # the ladder has no table: three sizes cover its regimes (X_DEFER = 8, 4 resident blocks)
            part.zero_()
            fn(n, part)
            torch.cuda.synchronize()
            assert torch.equal(part[: n * rec], full[: n * rec]), (name, n)
            assert int(part[n * rec : n * rec + 64].sum()) == 0, (name, n)   # nothing written past the batch
    kh, k2h, uh = k[: 32 * 24].cpu().numpy().tobytes(), k2[: 32 * 24].cpu().numpy().tobytes(), u[: 32 * 24].cpu().numpy().tobytes()
    o = torch.empty(24 * 33, dtype=torch.uint8, device="cuda")
    engine.batch_mul_dev(0, k, None, 24, o, stream=st); torch.cuda.synchronize()
    assert o.cpu().numpy().tobytes() == pm.batch_mul(0, kh, None, 24, 0)[0]
    engine.batch_mul_dev(1, k2, None, 24, o, stream=st); torch.cuda.synchronize()
    assert o.cpu().numpy().tobytes() == pm.batch_mul(1, k2h, None, 24, 0)[0]
    engine.batch_x25519_dev(k, u, 24, o, stream=st); torch.cuda.synchronize()
    assert o[: 24 * 32].cpu().numpy().tobytes() == b"".join(pm.x25519(kh[32 * i : 32 * i + 32], uh[32 * i : 32 * i + 32]) for i in range(24))
    engine.batch_ed25519_pubkey_dev(k, True, 24, o, stream=st); torch.cuda.synchronize()
    assert o[: 24 * 32].cpu().numpy().tobytes() == b"".join(pm.ed25519_pubkey(kh[32 * i : 32 * i + 32]) for i in range(24))


def test_imad_peak_runs(engine):
    rates = [engine.imad_peak(v) for v in (0, 1, 2)]
    assert all(r > 1e11 for r in rates), rates


def test_forge_quirk_mode_replays_the_reference_literally(engine, fec, pm):
    """FEC_SEM_FORGE_QUIRK: BASELINE.json configs[0] (secp256k1 k*G over 2^16 synthetic scalars) and a k*P batch,
    byte for byte against oracle/ref (the literal restatement of forge-ec's own limb code) and the Appendix A.2
    goldens; the same for P-256; compressed input is refused."""
    import ctypes
    import json
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "oracle")], check=True, stdout=subprocess.DEVNULL)
    oref = ctypes.CDLL(os.path.join(root, "oracle", "_build", "liboracle_ref.so"))
    lit = json.load(open(os.path.join(root, "tests", "golden", "forge_ec_literal.json")))
    for ks, exp in lit["secp256k1_kG_compressed"].items():
        out, st = engine.batch_mul(0, _be(int(ks)), None, sem=fec.SEM_FORGE_QUIRK)
        assert out.tobytes().hex() == exp and st[0] == 0
    n = 1 << 16
    k = pm.random_scalars_be(0, n, stream=1)          # config 1 inputs, edge lanes included
    out, st = engine.batch_mul(0, k, None, sem=fec.SEM_FORGE_QUIRK)
    exp = np.zeros(n * 33, dtype=np.uint8)
    oref.oref_secp256k1_batch_mul(ctypes.c_void_p(k.ctypes.data), None, ctypes.c_size_t(n), ctypes.c_void_p(exp.ctypes.data),
                                  os.cpu_count() or 1)
    assert out.tobytes() == exp.tobytes()
    # the reference's N constant has its two top limbs swapped (secp256k1.rs:27-28): that is the bound its parser applies
    n_ref = (0xFFFFFFFFFFFFFFFE << 192) | (0xFFFFFFFFFFFFFFFF << 128) | (pm.CURVES[0].n & ((1 << 128) - 1))
    kv = [int.from_bytes(k[i].tobytes(), "big") for i in range(n)]
    assert st.tolist() == [1 if v >= n_ref else 3 if v == 0 else 0 for v in kv]
    m = 2048
    r = pm.random_scalars_be(0, m, stream=3, sub=1, edge=False)
    pts, _ = engine.batch_mul(0, r, None, out_fmt=fec.OUT_AFFINE_XY_BE)          # true points r_i G
    out, st = engine.batch_mul(0, k[:m], pts, sem=fec.SEM_FORGE_QUIRK)
    exp = np.zeros(m * 33, dtype=np.uint8)
    pc = np.ascontiguousarray(pts)
    oref.oref_secp256k1_batch_mul(ctypes.c_void_p(k.ctypes.data), ctypes.c_void_p(pc.ctypes.data), ctypes.c_size_t(m),
                                  ctypes.c_void_p(exp.ctypes.data), os.cpu_count() or 1)
    assert out.tobytes() == exp.tobytes()
    with pytest.raises(fec.FecError):   # compressed input is refused: the reference's own decoder rejects its own encodings
        engine.batch_mul(0, k[:4], pts[:4, :33].copy(), sem=fec.SEM_FORGE_QUIRK, points_fmt=fec.PT_SEC1_COMPRESSED)
    # P-256: the same literal parity against oracle/ref/oracle_ref_p256.c -- k*G over 2^14 synthetic scalars (edge lanes
    # included) and a k*P batch on true points and on arbitrary coordinates (the reference's from_affine validates nothing)
    for ks, exp in lit["p256_kG_compressed"].items():
        out, st = engine.batch_mul(1, _be(int(ks)), None, sem=fec.SEM_FORGE_QUIRK)
        assert out.tobytes().hex() == exp and st[0] == 0
    n = 1 << 14
    k = pm.random_scalars_be(1, n, stream=1)
    out, st = engine.batch_mul(1, k, None, sem=fec.SEM_FORGE_QUIRK)
    exp = np.zeros(n * 33, dtype=np.uint8)
    oref.oref_p256_batch_mul(ctypes.c_void_p(k.ctypes.data), None, ctypes.c_size_t(n), ctypes.c_void_p(exp.ctypes.data), os.cpu_count() or 1)
    assert out.tobytes() == exp.tobytes()
    kv = [int.from_bytes(k[i].tobytes(), "big") for i in range(n)]
    assert st.tolist() == [1 if v >= pm.CURVES[1].n else 3 if v == 0 else 0 for v in kv]
    m = 1024
    r = pm.random_scalars_be(1, m, stream=3, sub=1, edge=False)
    pts, _ = engine.batch_mul(1, r, None, out_fmt=fec.OUT_AFFINE_XY_BE)
    pc = np.ascontiguousarray(pts)
    pc[m // 2:] = np.frombuffer(pm.philox_bytes32(m, stream=9).tobytes(), dtype=np.uint8).reshape(m // 2, 64)   # arbitrary coordinates
    pc[m // 2:, 0] &= 0x7F
    pc[m // 2:, 32] &= 0x7F
    out, st = engine.batch_mul(1, k[:m], pc, sem=fec.SEM_FORGE_QUIRK)
    exp = np.zeros(m * 33, dtype=np.uint8)
    oref.oref_p256_batch_mul(ctypes.c_void_p(k.ctypes.data), ctypes.c_void_p(pc.ctypes.data), ctypes.c_size_t(m), ctypes.c_void_p(exp.ctypes.data),
                             os.cpu_count() or 1)
    assert out.tobytes() == exp.tobytes()


def test_forge_quirk_mode_replays_the_reference_x25519_literally(engine, fec, pm):
    """FEC_SEM_FORGE_QUIRK for X25519: 2^14 synthetic (scalar, u) lanes and 2^12 base-point lanes byte for byte against
    oracle/ref/oracle_ref_x25519.c, the Appendix A.2 goldens, and the device-pointer entry point."""
    import ctypes
    import json
    import os
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "oracle")], check=True, stdout=subprocess.DEVNULL)
    oref = ctypes.CDLL(os.path.join(root, "oracle", "_build", "liboracle_ref.so"))
    lit = json.load(open(os.path.join(root, "tests", "golden", "forge_ec_literal.json")))
    for ks, exp in lit["x25519_base_point"].items():
        out = engine.batch_x25519(np.frombuffer(bytes.fromhex(ks), dtype=np.uint8), None, n=1, sem=fec.SEM_FORGE_QUIRK)
        assert out.tobytes().hex() == exp
    n = 1 << 14
    ks = pm.philox_bytes32(n, stream=4).copy()
    us = pm.philox_bytes32(n, stream=4, sub=1).copy()
    ks[0] = 0
    ks[0, 0] = 2                      # the hard-coded scalar
    us[1] = 0                         # u = 0
    us[2] = 255                       # u = 2^256 - 1 (bit 255 masked, the rest is not below p)
    out = engine.batch_x25519(ks, us, sem=fec.SEM_FORGE_QUIRK)
    exp = np.zeros(n * 32, dtype=np.uint8)
    kc, uc = np.ascontiguousarray(ks), np.ascontiguousarray(us)
    oref.oref_x25519_batch(ctypes.c_void_p(kc.ctypes.data), ctypes.c_void_p(uc.ctypes.data), ctypes.c_size_t(n), ctypes.c_void_p(exp.ctypes.data),
                           os.cpu_count() or 1)
    assert out.tobytes() == exp.tobytes()
    m = 1 << 12
    out = engine.batch_x25519(ks[:m], None, sem=fec.SEM_FORGE_QUIRK)
    oref.oref_x25519_batch(ctypes.c_void_p(kc.ctypes.data), None, ctypes.c_size_t(m), ctypes.c_void_p(exp.ctypes.data), os.cpu_count() or 1)
    assert out.tobytes() == exp[: 32 * m].tobytes()
    dk, du = torch.from_numpy(kc[:m].copy()).cuda(), torch.from_numpy(uc[:m].copy()).cuda()
    do = torch.empty(m * 32, dtype=torch.uint8, device="cuda")
    engine.batch_x25519_dev(dk, du, m, do, sem=fec.SEM_FORGE_QUIRK)
    torch.cuda.synchronize()
    oref.oref_x25519_batch(ctypes.c_void_p(kc.ctypes.data), ctypes.c_void_p(uc.ctypes.data), ctypes.c_size_t(m), ctypes.c_void_p(exp.ctypes.data), 4)
    assert do.cpu().numpy().tobytes() == exp[: 32 * m].tobytes()
    # the standard semantics are untouched by the selector
    assert engine.batch_x25519(ks[:64], us[:64]).tobytes() == engine.batch_x25519(ks[:64], us[:64], sem=fec.SEM_STD).tobytes()


def test_forge_quirk_mode_replays_the_reference_ed25519_keygen_literally(engine, fec, pm):
    """FEC_SEM_FORGE_QUIRK for Ed25519 key generation: 2^13 synthetic seeds byte for byte against
    oracle/ref/oracle_ref_ed25519.c, the Appendix A.2 goldens and the device-pointer entry point; scalars are refused (the
    reference's derive_public_key only exists for seeds)."""
    import ctypes
    import json
    import os
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "oracle")], check=True, stdout=subprocess.DEVNULL)
    oref = ctypes.CDLL(os.path.join(root, "oracle", "_build", "liboracle_ref.so"))
    lit = json.load(open(os.path.join(root, "tests", "golden", "forge_ec_literal.json")))
    for seed, exp in lit["ed25519_derive_public_key"].items():
        out = engine.batch_ed25519_pubkey(np.frombuffer(bytes.fromhex(seed), dtype=np.uint8), True, n=1, sem=fec.SEM_FORGE_QUIRK)
        assert out.tobytes().hex() == exp
    n = 1 << 13
    seeds = np.ascontiguousarray(pm.philox_bytes32(n, stream=5))     # ~32 of them start with 0x9d: the hard-coded branch
    out = engine.batch_ed25519_pubkey(seeds, True, sem=fec.SEM_FORGE_QUIRK)
    exp = np.zeros(n * 32, dtype=np.uint8)
    oref.oref_ed25519_batch_pubkey(ctypes.c_void_p(seeds.ctypes.data), ctypes.c_size_t(n), ctypes.c_void_p(exp.ctypes.data), os.cpu_count() or 1)
    assert out.tobytes() == exp.tobytes()
    assert sum(1 for i in range(n) if seeds[i, 0] == 0x9D) > 0
    m = 1024
    ds = torch.from_numpy(seeds[:m].copy()).cuda()
    do = torch.empty(m * 32, dtype=torch.uint8, device="cuda")
    engine.batch_ed25519_pubkey_dev(ds, True, m, do, sem=fec.SEM_FORGE_QUIRK)
    torch.cuda.synchronize()
    assert do.cpu().numpy().tobytes() == exp[: 32 * m].tobytes()
    with pytest.raises(fec.FecError):
        engine.batch_ed25519_pubkey(seeds[:4], False, sem=fec.SEM_FORGE_QUIRK)


def test_multi_device_context_shards_lanes(fec, pm):
    """fec_ctx_create over every visible device: contiguous lane shards, one host thread and pipeline per device,
    results identical to the single-device run (skipped on a one-GPU box)."""
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 5000 + 1024 * g
    k = pm.random_scalars_be(0, n, stream=3)
    r = pm.random_scalars_be(0, n, stream=3, sub=1, edge=False)
    with fec.Engine() as one, fec.Engine(devices=list(range(g))) as many:
        assert many.device_count == g
        pts, _ = one.batch_mul(0, r, None, out_fmt=fec.OUT_AFFINE_XY_BE)
        a, sa = one.batch_mul(0, k, pts)
        b, sb = many.batch_mul(0, k, pts)
        assert a.tobytes() == b.tobytes() and sa.tobytes() == sb.tobytes()
        assert many.batch_x25519(k).tobytes() == one.batch_x25519(k).tobytes()
        assert many.batch_ed25519_pubkey(k).tobytes() == one.batch_ed25519_pubkey(k).tobytes()
        exp, est = pm.batch_mul(0, k[:64].tobytes(), pts[:64].tobytes(), 64, 0)
        assert b[:64].tobytes() == exp and sb[:64].tobytes() == est


@pytest.mark.parametrize("curve", [0, 1])
def test_batched_ecdh_matches_model(engine, fec, pm, curve):
    """x(k*P) + the reference's key derivations, against the model (hashlib HKDF); both parties of an
    exchange derive the same key; rejected lanes give zero records and their status."""
    c = pm.CURVES[curve]
    n = 300
    a = pm.random_scalars_be(curve, n, stream=6, edge=False)
    b = pm.random_scalars_be(curve, n, stream=6, sub=1, edge=False)
    A, _ = engine.batch_mul(curve, a, None, out_fmt=fec.OUT_AFFINE_XY_BE)
    B, _ = engine.batch_mul(curve, b, None, out_fmt=fec.OUT_AFFINE_XY_BE)
    s1, st1 = engine.batch_ecdh(curve, a, B)
    s2, st2 = engine.batch_ecdh(curve, b, A)
    assert s1.tobytes() == s2.tobytes() and not st1.any() and not st2.any()
    exp, est = pm.batch_ecdh(curve, a[:40].tobytes(), B[:40].tobytes(), 40)
    assert s1[:40].tobytes() == exp
    xs, _ = engine.batch_mul(curve, a, B, out_fmt=fec.OUT_X_BE)          # the 32-byte output format on its own
    assert xs.tobytes() == s1.tobytes()
    for kdf, info, out_len in ((fec.KDF_HKDF_SHA256, b"forge-ec ecdh test", 32), (fec.KDF_HKDF_SHA256, b"", 64),
                               (fec.KDF_HKDF_SHA256, bytes(range(200)), 48), (fec.KDF_HKDF_SHA256, bytes(range(256)), 33),
                               (fec.KDF_HKDF_SHA256, b"ab", 17), (fec.KDF_XOR_INFO, b"info", 40), (fec.KDF_XOR_INFO, bytes(range(70)), 64)):
        k1, _ = engine.batch_ecdh(curve, a, B, kdf=kdf, info=info, out_len=out_len)
        k2, _ = engine.batch_ecdh(curve, b, A, kdf=kdf, info=info, out_len=out_len)
        assert k1.tobytes() == k2.tobytes()
        exp, _ = pm.batch_ecdh(curve, a[:40].tobytes(), B[:40].tobytes(), 40, kdf, info, out_len)
        assert k1[:40].tobytes() == exp
    # rejected lanes: scalar >= n, scalar 0, point off the curve; compressed public keys
    bad_k = np.frombuffer(_be(c.n) + _be(0) + _be(5) + _be(7), dtype=np.uint8).reshape(4, 32)
    pts = np.concatenate([B[:2], np.frombuffer(_be(2) + _be(2), dtype=np.uint8).reshape(1, 64), B[3:4]])
    out, st = engine.batch_ecdh(curve, bad_k, pts, kdf=fec.KDF_HKDF_SHA256, info=b"x", out_len=32)
    exp, est = pm.batch_ecdh(curve, bad_k.tobytes(), pts.tobytes(), 4, fec.KDF_HKDF_SHA256, b"x", 32)
    assert out.tobytes() == exp and st.tobytes() == est and st.tolist() == [1, 3, 2, 0]
    comp, _ = engine.batch_mul(curve, b, None)
    k3, _ = engine.batch_ecdh(curve, a, comp, kdf=fec.KDF_HKDF_SHA256, info=b"c", out_len=32, points_fmt=fec.PT_SEC1_COMPRESSED)
    k4, _ = engine.batch_ecdh(curve, a, B, kdf=fec.KDF_HKDF_SHA256, info=b"c", out_len=32)
    assert k3.tobytes() == k4.tobytes()
    # the trait mirror
    cls = fec.Secp256k1 if curve == 0 else fec.P256
    sec = cls.derive_shared_secret(fec.Scalar(a[0].tobytes()), fec.AffinePoint(B[0].tobytes()), engine)
    assert sec == s1[0].tobytes()


@pytest.mark.parametrize("curve", [0, 1])
def test_multi_scalar_multiply_matches_model(engine, fec, ostd, pm, curve):
    """sum_i k_i*P_i = the batched products + a device-wide reduction, against the oracle that adds the
    products one by one (lib.rs:934-951; oracle/std `ostd_msm`, and the Python model on a small case); cancellations,
    repeated points (the doubling case of the reduction), identity lanes, rejected lanes, one to several thousand points.
    The expected values never come from the GPU (the bucket method and 2^17 points: test_gpu_parity_large.py)."""
    c = pm.CURVES[curve]
    for n in (1, 2, 33, 257, 5000):
        k = pm.random_scalars_be(curve, n, stream=7, edge=(n >= 33))
        r = pm.random_scalars_be(curve, n, stream=7, sub=1, edge=False)
        pts = np.frombuffer(ostd.mul(curve, r.tobytes(), None, 0, 2)[0], dtype=np.uint8).reshape(n, 64)
        assert engine.msm(curve, k, pts) == ostd.msm(curve, k.tobytes(), pts.tobytes(), 0, 0), (curve, n)
    # a small case entirely in the Python model
    n = 6
    k = pm.random_scalars_be(curve, n, stream=8, edge=False)
    r = pm.random_scalars_be(curve, n, stream=8, sub=1, edge=False)
    pts = np.frombuffer(pm.batch_mul(curve, r.tobytes(), None, n, pm.OUT_AFFINE_XY_BE)[0], dtype=np.uint8).reshape(n, 64)
    assert engine.msm(curve, k, pts, out_fmt=fec.OUT_SEC1_UNCOMPRESSED) == pm.msm(curve, k.tobytes(), pts.tobytes(), n, 1)
    # k P + (n - k) P = identity; P + P + P + P (every combine step is a doubling); identity lanes only
    kk = int.from_bytes(k[0].tobytes(), "big")
    two = np.frombuffer(_be(kk) + _be(c.n - kk), dtype=np.uint8)
    rec, st = engine.msm(curve, two, np.concatenate([pts[0], pts[0]]))
    assert rec == bytes(33) and st == 3
    four = np.frombuffer(_be(1) * 64, dtype=np.uint8)
    rec, st = engine.msm(curve, four, np.tile(pts[1], 64))
    exp, _ = pm.batch_mul(curve, _be(64), pts[1].tobytes(), 1, 0)
    assert rec == exp and st == 0
    rec, st = engine.msm(curve, np.zeros(96, dtype=np.uint8), pts[:3])
    assert rec == bytes(33) and st == 3
    # the identity as an input point is simply added (lib.rs:934-951): it changes nothing
    withid = np.concatenate([pts[:3].reshape(-1), np.zeros(64, dtype=np.uint8), pts[3:].reshape(-1)])
    kid = np.concatenate([k[:3].reshape(-1), k[0], k[3:].reshape(-1)])
    assert engine.msm(curve, kid, withid) == pm.msm(curve, k.tobytes(), pts.tobytes(), n, 0)
    # rejected lanes
    bad = np.frombuffer(_be(5) + _be(c.n), dtype=np.uint8)
    assert engine.msm(curve, bad, pts[:2]) == (bytes(33), 1)
    off = np.concatenate([pts[0], np.frombuffer(_be(2) + _be(2), dtype=np.uint8)])
    assert engine.msm(curve, np.frombuffer(_be(5) + _be(6), dtype=np.uint8), off) == (bytes(33), 2)
    # trait mirror
    cls = fec.Secp256k1 if curve == 0 else fec.P256
    P = cls.multi_scalar_multiply([fec.AffinePoint(pts[i].tobytes()) for i in range(n)], [fec.Scalar(k[i].tobytes()) for i in range(n)], engine)
    assert P.to_bytes() == pm.msm(curve, k.tobytes(), pts.tobytes(), n, 0)[0]
    assert cls.multi_scalar_multiply([], [], engine).is_identity()
    # Curve::multiply(identity, k) is the identity (lib.rs:820-832): a k = 0 result fed back into multiply
    O = cls.multiply(None, fec.Scalar.from_int(0), engine)
    assert O.is_identity() and cls.multiply(O, fec.Scalar.from_int(5), engine).is_identity()


@pytest.mark.parametrize("curve", [0, 1])
def test_msm_bucket_method_negation_property(engine, fec, ostd, pm, curve):
    """the bucket (Pippenger) path at n >= 10^5 (oracle comparison: test_gpu_parity_large.py): sum_i k_i P_i with every k_i
    negated is the negated point -- same x, other y parity"""
    c = pm.CURVES[curve]
    n = 140000
    k = pm.random_scalars_be(curve, n, stream=9)
    r = pm.random_scalars_be(curve, n, stream=9, sub=1, edge=False)
    pts, _ = engine.batch_mul(curve, r, None, out_fmt=fec.OUT_AFFINE_XY_BE)
    kneg = np.frombuffer(b"".join(_be((c.n - int.from_bytes(k[i].tobytes(), "big")) % c.n) for i in range(n)), dtype=np.uint8)
    a, sa = engine.msm(curve, k, pts)
    b, sb = engine.msm(curve, kneg, pts)
    assert sa == sb == 0 and a[1:] == b[1:] and a[0] == b[0] ^ 1


def test_cpp_host_side_on_gpu():
    """tests/cpp/test_traits.cpp: the reference's own trait-level checks (generator, small multiples, n-1, out-of-range
    scalar, off-curve point, ECDH commutes, multi_scalar_multiply, RFC 7748 6.1, RFC 8032 TEST 1) through the C++ host side
    and the host-buffer C ABI."""
    import subprocess
    from test_host import _build_cpp_traits_test
    r = subprocess.run([_build_cpp_traits_test()], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


_AB_CHILD = r"""
import hashlib, importlib, os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
fec = importlib.import_module("forge-ec_b200")
eng = fec.Engine()
n = 4096
h = hashlib.sha256()
rng = np.random.default_rng(20261018)
for curve in (0, 1):
    order = (0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141, 0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551)[curve]
    k = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    k[:, 0] &= 0x7F                                   # below both group orders
    for i, v in enumerate((0, 1, 2, order - 1, order - 2, (order + 1) // 2)):
        k[i] = np.frombuffer(v.to_bytes(32, "big"), dtype=np.uint8)
    r = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    r[:, 0] &= 0x7F
    r[:, 31] |= 1
    pts, st0 = eng.batch_mul(curve, r, None, out_fmt=fec.OUT_AFFINE_XY_BE)
    out1, st1 = eng.batch_mul(curve, k, None)
    out2, st2 = eng.batch_mul(curve, k, pts)
    for a in (pts, st0, out1, st1, out2, st2):
        h.update(np.ascontiguousarray(a).tobytes())
ks = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
us = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
us[0] = 0
us[1] = 255
h.update(np.ascontiguousarray(eng.batch_x25519(ks, us)).tobytes())
ks[2] = 0
ks[3] = 255
h.update(np.ascontiguousarray(eng.batch_x25519(ks, None)).tobytes())              # base point: the Ed25519 fixed-base route
h.update(np.ascontiguousarray(eng.batch_ed25519_pubkey(ks, True)).tobytes())      # seeds
h.update(np.ascontiguousarray(eng.batch_ed25519_pubkey(ks, False)).tobytes())     # scalars as they are
print("DIGEST", h.hexdigest())
"""


def test_every_kernel_form_returns_the_same_bytes(tmp_path):
    """The forms kept for A/B measurements (single-kernel k*P, register-resident / accumulator-machine fixed base, the
    accumulator-machine X25519 ladder) against the default kernels on 4096 seeded lanes per entry point, edge scalars included.
    The switches are read once per process, so each form runs in a child process and reports a digest of all outputs.
    FORGE_EC_CT_TABLE=1 covers k*G of both Weierstrass curves AND the Ed25519 / X25519-base-point comb (scanned 4-bit tables)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "ab_child.py"
    script.write_text(_AB_CHILD)
    digests = {}
    for name, env in (("default", {}), ("var_mono", {"FORGE_EC_VAR_IMPL": "mono"}), ("fixed_regs", {"FORGE_EC_FIXED_IMPL": "regs"}),
                      ("fixed_vm", {"FORGE_EC_FIXED_IMPL": "vm"}), ("x_vm", {"FORGE_EC_X_IMPL": "vm"}),
                      ("fixed_ct_table", {"FORGE_EC_CT_TABLE": "1"}), ("ladder_regs", {"FORGE_EC_GLV_LADDER": "regs"})):
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, str(script), root], capture_output=True, text=True, env=e)
        assert r.returncode == 0, (name, r.stdout[-2000:], r.stderr[-2000:])
        digests[name] = [l for l in r.stdout.splitlines() if l.startswith("DIGEST")][0]
    assert len(set(digests.values())) == 1, digests


def test_nonce_commit_matches_oracle_and_published_vectors(engine, fec, ostd, pm):
    """the signing-side fixed-base step through fec_batch_nonce_commit -- ECDSA r = x(k G) (ecdsa.rs:108-131),
    BIP-340 even-y rule (schnorr.rs:372-385), EdDSA [r mod L]B from the 64-byte hash value (RFC 8032 5.1.6) -- against RFC 6979 /
    BIP-340 / RFC 8032 vectors and the C oracle, rejected lanes included"""
    from test_oracle import nonce_commit_goldens

    def commit(curve, mode, nonces):
        R, r, st = engine.batch_nonce_commit(curve, mode, np.frombuffer(nonces, dtype=np.uint8))
        return R.tobytes(), r.tobytes(), st.tobytes()

    nonce_commit_goldens(pm, commit)
    n = 3000
    for curve in (0, 1):
        c = pm.CURVES[curve]
        k = pm.random_scalars_be(curve, n, stream=12)
        for i, v in enumerate(_edge_scalars(c)):
            k[8 + i] = np.frombuffer(_be(v % 2**256), dtype=np.uint8)
        for mode in ((fec.NONCE_ECDSA, fec.NONCE_BIP340) if curve == 0 else (fec.NONCE_ECDSA,)):
            assert commit(curve, mode, k.tobytes()) == ostd.nonce_commit(curve, mode, k.tobytes()), (curve, mode)
    h = pm.philox_bytes32(2 * n, stream=13).reshape(n, 64).copy()
    L = pm.ED_L
    for i, v in enumerate([0, 2**512 - 1, L, L - 1, 2**252, (L << 200) + 5, 2**511, (L << 259) - 1, 2**252 - 1, L + 1]):
        h[i] = np.frombuffer(v.to_bytes(64, "little"), dtype=np.uint8)
    assert commit(0, fec.NONCE_EDDSA, h.tobytes()) == ostd.nonce_commit(0, 2, h.tobytes())
    # the BIP-340 mode belongs to secp256k1
    with pytest.raises(fec.FecError):
        engine.batch_nonce_commit(1, fec.NONCE_BIP340, k)


@pytest.mark.parametrize("curve", [0, 1])
def test_point_decode_matches_oracle(engine, fec, ostd, pm, curve):
    """the missing standalone decode of round 1: fec_batch_point_decode = CompressedPoint::to_affine (point.rs:70-144) +
    Curve::validate_point (lib.rs:905-923) without a scalar multiplication; every input and output format"""
    c = pm.CURVES[curve]
    n = 1500
    r = pm.random_scalars_be(curve, n, stream=14, edge=False)
    comp = np.frombuffer(ostd.mul(curve, r.tobytes(), None, 0, 0)[0], dtype=np.uint8).reshape(n, 33).copy()
    comp[3, 0] = 4; comp[5, 0] = 0; comp[7, 1:] = 0xFF; comp[9, 32] ^= 1; comp[11] = 0
    for ofmt in (0, 1, 2, 3, 4):
        out, st = engine.batch_point_decode(curve, comp, out_fmt=ofmt)
        assert (out.tobytes(), st.tobytes()) == ostd.point_decode(curve, comp.tobytes(), 1, ofmt)
    assert {0, 2, 3} <= set(st.tolist())
    xy = np.frombuffer(ostd.mul(curve, r.tobytes(), None, 0, 2)[0], dtype=np.uint8).reshape(n, 64).copy()
    xy[3, 63] ^= 1; xy[5] = 0; xy[7] = 0xFF
    out, st = engine.batch_point_decode(curve, xy, points_fmt=fec.PT_AFFINE_XY_BE, out_fmt=fec.OUT_SEC1_COMPRESSED)
    assert (out.tobytes(), st.tobytes()) == ostd.point_decode(curve, xy.tobytes(), 0, 0)
    cls = fec.Secp256k1 if curve == 0 else fec.P256
    pts = cls.batch_decode([comp[0].tobytes(), comp[1].tobytes(), comp[5].tobytes()], engine)
    assert pts[0].to_bytes() == comp[0].tobytes() and pts[2].is_identity()
    with pytest.raises(ValueError):
        cls.batch_decode([comp[3].tobytes()], engine)
