"""oracle/pymodel.py -- Python big-integer cross-model of the hot path (TEST INFRASTRUCTURE).

This file is part of the *oracle*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
(``forge-ec_b200/``) never does.

It states, with Python integers, what the reference *specifies* for the path (the
``oracle_std`` half of SURVEY.md section 0's two-oracle contract):

* secp256k1 / P-256 ``Curve::multiply`` followed by ``to_affine`` and SEC1 encoding
  (forge-ec-curves/src/secp256k1.rs:2635-2692, 1342-1363, 875-894;
  forge-ec-curves/src/p256.rs:2120-2156, 1835-1857, 1558-1577;
  forge-ec-encoding/src/point.rs:39-67, 186-211),
* ``x25519(scalar, u)`` (forge-ec-curves/src/curve25519.rs:1624-1716) = RFC 7748,
* ``Ed25519Signature::derive_public_key`` (forge-ec-signature/src/eddsa.rs:450-509) = RFC 8032
  section 5.1.5,

plus the Philox-4x32-10 input generator of SURVEY.md section 8d.  It is slow; it exists to pin the C
oracle (``oracle/std``) and, through it, the CUDA path.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass

import numpy as np

# --------------------------------------------------------------------------------------
# curve parameters (SEC2 / FIPS 186-4 values the reference quotes:
# secp256k1.rs:22-28, 2610-2622; p256.rs:18-33, 2094-2106)
# --------------------------------------------------------------------------------------


@dataclass(frozen=True)
class WeierstrassCurve:
    name: str
    p: int
    a: int
    b: int
    n: int
    gx: int
    gy: int


SECP256K1 = WeierstrassCurve(
    "secp256k1",
    p=2**256 - 2**32 - 977,
    a=0,
    b=7,
    n=0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141,
    gx=0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
    gy=0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8,
)
P256 = WeierstrassCurve(
    "p256",
    p=2**256 - 2**224 + 2**192 + 2**96 - 1,
    a=-3,
    b=0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B,
    n=0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551,
    gx=0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
    gy=0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5,
)
CURVES = {0: SECP256K1, 1: P256, "secp256k1": SECP256K1, "p256": P256}

P25519 = 2**255 - 19
A24 = 121665
ED_D = (-121665 * pow(121666, -1, P25519)) % P25519
ED_L = 2**252 + 27742317777372353535851937790883648493
ED_BY = (4 * pow(5, -1, P25519)) % P25519


def _ed_recover_x(y: int, sign: int) -> int:
    x2 = (y * y - 1) * pow(ED_D * y * y + 1, -1, P25519) % P25519
    x = pow(x2, (P25519 + 3) // 8, P25519)
    if (x * x - x2) % P25519 != 0:
        x = x * pow(2, (P25519 - 1) // 4, P25519) % P25519
    assert (x * x - x2) % P25519 == 0
    if x & 1 != sign:
        x = P25519 - x
    return x


ED_BX = _ed_recover_x(ED_BY, 0)

# --------------------------------------------------------------------------------------
# short Weierstrass arithmetic (Jacobian, Python ints)
# --------------------------------------------------------------------------------------


def _jdbl(c: WeierstrassCurve, P):
    X, Y, Z = P
    if Z == 0 or Y == 0:
        return (1, 1, 0)
    p = c.p
    S = 4 * X * Y * Y % p
    M = (3 * X * X + c.a * pow(Z, 4, p)) % p
    X3 = (M * M - 2 * S) % p
    Y3 = (M * (S - X3) - 8 * pow(Y, 4, p)) % p
    Z3 = 2 * Y * Z % p
    return (X3, Y3, Z3)


def _jadd(c: WeierstrassCurve, P, Q):
    if P[2] == 0:
        return Q
    if Q[2] == 0:
        return P
    p = c.p
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    Z1Z1 = Z1 * Z1 % p
    Z2Z2 = Z2 * Z2 % p
    U1 = X1 * Z2Z2 % p
    U2 = X2 * Z1Z1 % p
    S1 = Y1 * Z2 * Z2Z2 % p
    S2 = Y2 * Z1 * Z1Z1 % p
    if U1 == U2:
        if S1 == S2:
            return _jdbl(c, P)
        return (1, 1, 0)
    H = (U2 - U1) % p
    R = (S2 - S1) % p
    HH = H * H % p
    HHH = H * HH % p
    V = U1 * HH % p
    X3 = (R * R - HHH - 2 * V) % p
    Y3 = (R * (V - X3) - S1 * HHH) % p
    Z3 = H * Z1 * Z2 % p
    return (X3, Y3, Z3)


def w_mul(c: WeierstrassCurve, k: int, P):
    """k*P for affine P=(x,y) or None (infinity); returns affine (x,y) or None."""
    if P is None or k % c.n == 0:
        return None
    k %= c.n
    Q = (1, 1, 0)
    B = (P[0], P[1], 1)
    for bit in bin(k)[2:]:
        Q = _jdbl(c, Q)
        if bit == "1":
            Q = _jadd(c, Q, B)
    if Q[2] == 0:
        return None
    zi = pow(Q[2], -1, c.p)
    return (Q[0] * zi * zi % c.p, Q[1] * zi * zi * zi % c.p)


def w_is_on_curve(c: WeierstrassCurve, P) -> bool:
    x, y = P
    return 0 <= x < c.p and 0 <= y < c.p and (y * y - (x * x * x + c.a * x + c.b)) % c.p == 0


# --------------------------------------------------------------------------------------
# encodings (forge-ec-encoding/src/point.rs:39-67 compressed, 186-211 uncompressed;
# identity -> 0x00 followed by zero padding)
# --------------------------------------------------------------------------------------

OUT_SEC1_COMPRESSED, OUT_SEC1_UNCOMPRESSED, OUT_AFFINE_XY_BE, OUT_X_BE, OUT_SEC1_HYBRID = 0, 1, 2, 3, 4
OUT_SIZE = {0: 33, 1: 65, 2: 64, 3: 32, 4: 65}


def sec1_encode(P, fmt: int = OUT_SEC1_COMPRESSED) -> bytes:
    if P is None:
        return bytes(OUT_SIZE[fmt])
    x, y = P
    if fmt == OUT_SEC1_COMPRESSED:
        return bytes([2 | (y & 1)]) + x.to_bytes(32, "big")
    if fmt == OUT_SEC1_UNCOMPRESSED:
        return b"\x04" + x.to_bytes(32, "big") + y.to_bytes(32, "big")
    if fmt == OUT_X_BE:
        return x.to_bytes(32, "big")
    if fmt == OUT_SEC1_HYBRID:  # PointFormat::Hybrid (secp256k1.rs:1048-1066, p256.rs:1704-1722)
        return bytes([6 | (y & 1)]) + x.to_bytes(32, "big") + y.to_bytes(32, "big")
    return x.to_bytes(32, "big") + y.to_bytes(32, "big")


def sec1_decompress(c: WeierstrassCurve, data: bytes):
    """33-byte SEC1 compressed -> affine or None (invalid / identity)."""
    if data[0] not in (2, 3):
        return None
    x = int.from_bytes(data[1:33], "big")
    if x >= c.p:
        return None
    rhs = (x * x * x + c.a * x + c.b) % c.p
    y = pow(rhs, (c.p + 1) // 4, c.p)
    if y * y % c.p != rhs:
        return None
    if (y & 1) != (data[0] & 1):
        y = c.p - y
    return (x, y)


# status codes of fec_batch_mul (include/forge_ec_cuda.h)
ST_OK, ST_SCALAR_RANGE, ST_POINT_INVALID, ST_INFINITY = 0, 1, 2, 3


def batch_mul(curve, scalars: bytes, points: bytes | None, n: int, fmt: int = 0, points_fmt: int = 0):
    """Model of fec_batch_mul: returns (out_bytes, status_bytes)."""
    c = CURVES[curve]
    osz = OUT_SIZE[fmt]
    out = bytearray(n * osz)
    st = bytearray(n)
    psz = 64 if points_fmt == 0 else 33
    for i in range(n):
        k = int.from_bytes(scalars[32 * i : 32 * i + 32], "big")
        if k >= c.n:
            st[i] = ST_SCALAR_RANGE
            continue
        if points is None:
            P = (c.gx, c.gy)
        else:
            raw = points[psz * i : psz * (i + 1)]
            # the identity as an input point: the all-zero X || Y record (the engine's own identity encoding, point.rs:41-47)
            # resp. prefix 0x00 (point.rs:82-87); Curve::multiply(identity, k) is the identity (lib.rs:820-832)
            if (points_fmt == 0 and raw == bytes(64)) or (points_fmt != 0 and raw[0] == 0):
                st[i] = ST_INFINITY
                continue
            if points_fmt == 0:
                P = (int.from_bytes(raw[:32], "big"), int.from_bytes(raw[32:], "big"))
                if not w_is_on_curve(c, P):
                    P = "bad"
            else:
                P = sec1_decompress(c, raw)
                if P is None:
                    P = "bad"
            if P == "bad":
                st[i] = ST_POINT_INVALID
                continue
        R = w_mul(c, k, P)
        if R is None:
            st[i] = ST_INFINITY
        out[osz * i : osz * (i + 1)] = sec1_encode(R, fmt)
    return bytes(out), bytes(st)


def point_decode(curve, points: bytes, n: int, fmt: int = OUT_AFFINE_XY_BE, points_fmt: int = 1):
    """Model of fec_batch_point_decode: CompressedPoint::to_affine (forge-ec-encoding/src/point.rs:70-144) / on-curve validation
    (Curve::validate_point, forge-ec-core/src/lib.rs:905-923, cofactor 1) and re-encoding: (out_bytes, status_bytes)."""
    c = CURVES[curve]
    osz, psz = OUT_SIZE[fmt], (64 if points_fmt == 0 else 33)
    out, st = bytearray(n * osz), bytearray(n)
    for i in range(n):
        raw = points[psz * i : psz * (i + 1)]
        if (points_fmt == 0 and raw == bytes(64)) or (points_fmt != 0 and raw[0] == 0):
            st[i] = ST_INFINITY
            continue
        if points_fmt == 0:
            P = (int.from_bytes(raw[:32], "big"), int.from_bytes(raw[32:], "big"))
            if not w_is_on_curve(c, P):
                P = None
        else:
            P = sec1_decompress(c, raw)
        if P is None:
            st[i] = ST_POINT_INVALID
            continue
        out[osz * i : osz * (i + 1)] = sec1_encode(P, fmt)
    return bytes(out), bytes(st)


def w_add_affine(c: WeierstrassCurve, P, Q):
    """affine addition with every special case (None = identity)"""
    if P is None:
        return Q
    if Q is None:
        return P
    if P[0] == Q[0]:
        if (P[1] + Q[1]) % c.p == 0:
            return None
        lam = (3 * P[0] * P[0] + c.a) * pow(2 * P[1], -1, c.p) % c.p
    else:
        lam = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, c.p) % c.p
    x = (lam * lam - P[0] - Q[0]) % c.p
    return x, (lam * (P[0] - x) - P[1]) % c.p


def msm(curve, scalars: bytes, points: bytes, n: int, fmt: int = 0, points_fmt: int = 0, products: bytes | None = None):
    """Model of fec_msm: Curve::multi_scalar_multiply (lib.rs:934-951) adds multiply(points[i], scalars[i]) one by one.
    `products` (n x 64 B affine + status computed elsewhere, e.g. by the C oracle) skips the slow Python multiplications."""
    c = CURVES[curve]
    if products is None:
        xy, st = batch_mul(curve, scalars, points, n, OUT_AFFINE_XY_BE, points_fmt)
    else:
        xy, st = products
    if any(s in (ST_SCALAR_RANGE, ST_POINT_INVALID) for s in st):
        return bytes(OUT_SIZE[fmt]), (ST_SCALAR_RANGE if ST_SCALAR_RANGE in st else ST_POINT_INVALID)
    acc = None
    for i in range(n):
        if st[i] == 0:
            acc = w_add_affine(c, acc, (int.from_bytes(xy[64 * i : 64 * i + 32], "big"), int.from_bytes(xy[64 * i + 32 : 64 * i + 64], "big")))
    return sec1_encode(acc, fmt), (ST_INFINITY if acc is None else 0)


# --------------------------------------------------------------------------------------
# Batched ECDH: KeyExchange::derive_shared_secret + derive_key (secp256k1.rs:1846-1905, p256.rs:2281-2344)
# --------------------------------------------------------------------------------------
KDF_NONE, KDF_HKDF_SHA256, KDF_XOR_INFO = 0, 1, 2


def hkdf_sha256_zero_salt(secret: bytes, info: bytes, out_len: int) -> bytes:
    """Secp256k1::derive_key: RFC 5869 with a salt of 32 zero bytes, restated with hashlib/hmac."""
    import hashlib
    import hmac
    prk = hmac.new(bytes(32), secret, hashlib.sha256).digest()
    okm, t, counter = b"", b"", 1
    while len(okm) < out_len:
        t = hmac.new(prk, t + info + bytes([counter]), hashlib.sha256).digest()
        okm += t
        counter += 1
    return okm[:out_len]


def kdf_xor_info(secret: bytes, info: bytes, out_len: int) -> bytes:
    """P256::derive_key: the reference's placeholder (p256.rs:2314-2344)."""
    okm = bytearray(out_len)
    for i, b in enumerate(secret):
        if i < out_len:
            okm[i] ^= b
    for i, b in enumerate(info):
        if i < out_len:
            okm[i] ^= b
    return bytes(okm)


def batch_ecdh(curve, scalars: bytes, points: bytes, n: int, kdf: int = KDF_NONE, info: bytes = b"", out_len: int = 32,
               points_fmt: int = 0):
    """Model of fec_batch_ecdh: returns (out_bytes, status_bytes); rejected lanes give zero records."""
    xs, st = batch_mul(curve, scalars, points, n, OUT_X_BE, points_fmt)
    if kdf == KDF_NONE:
        return xs, st
    out = bytearray(n * out_len)
    for i in range(n):
        if st[i] == 0:
            f = hkdf_sha256_zero_salt if kdf == KDF_HKDF_SHA256 else kdf_xor_info
            out[out_len * i : out_len * (i + 1)] = f(xs[32 * i : 32 * i + 32], info, out_len)
    return bytes(out), st


# --------------------------------------------------------------------------------------
# X25519 (RFC 7748 section 5) -- what curve25519.rs:1624-1716 is shaped after
# --------------------------------------------------------------------------------------


def x25519(k: bytes, u: bytes) -> bytes:
    kb = bytearray(k)
    kb[0] &= 248
    kb[31] &= 127
    kb[31] |= 64
    kn = int.from_bytes(kb, "little")
    x1 = int.from_bytes(u, "little") & ((1 << 255) - 1)
    x1 %= P25519
    x2, z2, x3, z3, swap = 1, 0, x1, 1, 0
    p = P25519
    for t in range(254, -1, -1):
        kt = (kn >> t) & 1
        swap ^= kt
        if swap:
            x2, x3, z2, z3 = x3, x2, z3, z2
        swap = kt
        A = (x2 + z2) % p
        AA = A * A % p
        B = (x2 - z2) % p
        BB = B * B % p
        E = (AA - BB) % p
        C = (x3 + z3) % p
        D = (x3 - z3) % p
        DA = D * A % p
        CB = C * B % p
        x3 = (DA + CB) ** 2 % p
        z3 = x1 * (DA - CB) ** 2 % p
        x2 = AA * BB % p
        z2 = E * (AA + A24 * E) % p
    if swap:
        x2, x3, z2, z3 = x3, x2, z3, z2
    return (x2 * pow(z2, p - 2, p) % p).to_bytes(32, "little")


X25519_BASE = (9).to_bytes(32, "little")

# --------------------------------------------------------------------------------------
# Ed25519 public key (RFC 8032 section 5.1.5)
# --------------------------------------------------------------------------------------


def _ed_add(P, Q):
    p = P25519
    X1, Y1, Z1, T1 = P
    X2, Y2, Z2, T2 = Q
    A = (Y1 - X1) * (Y2 - X2) % p
    B = (Y1 + X1) * (Y2 + X2) % p
    C = 2 * T1 * T2 * ED_D % p
    D = 2 * Z1 * Z2 % p
    E, F, G, H = B - A, D - C, D + C, B + A
    return (E * F % p, G * H % p, F * G % p, E * H % p)


def ed_scalar_mul_base(s: int):
    Q = (0, 1, 1, 0)
    B = (ED_BX, ED_BY, 1, ED_BX * ED_BY % P25519)
    while s:
        if s & 1:
            Q = _ed_add(Q, B)
        B = _ed_add(B, B)
        s >>= 1
    return Q


def ed_encode(Q) -> bytes:
    X, Y, Z, _ = Q
    zi = pow(Z, -1, P25519)
    x, y = X * zi % P25519, Y * zi % P25519
    return (y | ((x & 1) << 255)).to_bytes(32, "little")


def ed25519_clamp(h32: bytes) -> int:
    a = bytearray(h32)
    a[0] &= 248
    a[31] &= 127
    a[31] |= 64
    return int.from_bytes(a, "little")


def ed25519_pubkey(data: bytes, input_is_seed: bool = True) -> bytes:
    """seed -> SHA-512 -> clamp(first half) -> [s]B -> y || sign(x).  With
    input_is_seed=False ``data`` is a 32-byte little-endian scalar used as is (mod L
    semantics are those of the group: any 256-bit integer is accepted)."""
    if input_is_seed:
        s = ed25519_clamp(hashlib.sha512(data).digest()[:32])
    else:
        s = int.from_bytes(data, "little")
    return ed_encode(ed_scalar_mul_base(s))


# --------------------------------------------------------------------------------------
# generic field model for the field microbench (config 2)
# --------------------------------------------------------------------------------------

PRIMES = {0: SECP256K1.p, 1: P256.p, 2: P25519}
FIELD_OP_MUL, FIELD_OP_SQR, FIELD_OP_MULSQR_CHAIN = 0, 1, 2


# --------------------------------------------------------------------------------------
# Signing-side nonce commitment: the R = k*G / [r]B step of a signature and the scalar that goes with it
# --------------------------------------------------------------------------------------
NONCE_ECDSA, NONCE_BIP340, NONCE_EDDSA = 0, 1, 2


def nonce_commit(curve, mode: int, nonces: bytes, n: int):
    """Model of fec_batch_nonce_commit: (out_R, out_r, status).
    ECDSA (forge-ec-signature/src/ecdsa.rs:108-131): R = k G compressed, r = x(R) as a scalar -- rejected (status 1) where the
    reference's Scalar::from_bytes fails (x >= n), status 3 for k = 0 or r = 0.  BIP-340 (schnorr.rs:372-385): x-only R and
    k' = n - k when y(R) is odd.  EdDSA (RFC 8032 5.1.6; eddsa.rs:120-131 shape): r = 64-byte little-endian value mod L, R = [r]B."""
    if mode == NONCE_EDDSA:
        R, r = bytearray(32 * n), bytearray(32 * n)
        for i in range(n):
            v = int.from_bytes(nonces[64 * i : 64 * i + 64], "little") % ED_L
            r[32 * i : 32 * i + 32] = v.to_bytes(32, "little")
            R[32 * i : 32 * i + 32] = ed_encode(ed_scalar_mul_base(v))
        return bytes(R), bytes(r), bytes(n)
    c = CURVES[curve]
    xy, st = batch_mul(curve, nonces, None, n, OUT_AFFINE_XY_BE)
    st = bytearray(st)
    rsz = 33 if mode == NONCE_ECDSA else 32
    R, r = bytearray(rsz * n), bytearray(32 * n)
    for i in range(n):
        if st[i]:
            continue
        x = int.from_bytes(xy[64 * i : 64 * i + 32], "big")
        y = int.from_bytes(xy[64 * i + 32 : 64 * i + 64], "big")
        k = int.from_bytes(nonces[32 * i : 32 * i + 32], "big")
        if mode == NONCE_ECDSA:
            if x >= c.n:
                st[i] = ST_SCALAR_RANGE
                continue
            if x == 0:
                st[i] = ST_INFINITY
                continue
            R[33 * i : 33 * i + 33] = bytes([2 | (y & 1)]) + x.to_bytes(32, "big")
            r[32 * i : 32 * i + 32] = x.to_bytes(32, "big")
        else:
            R[32 * i : 32 * i + 32] = x.to_bytes(32, "big")
            r[32 * i : 32 * i + 32] = ((c.n - k) if (y & 1) else k).to_bytes(32, "big")
    return bytes(R), bytes(r), bytes(st)


def field_bench_model(prime_id: int, op: int, iters: int, a: int, b: int) -> int:
    """What fec_field_bench computes per lane (little-endian 32-byte in/out):
    op=mul: x=a; repeat iters: x = x*b.   op=sqr: x=a; repeat: x = x^2.
    op=chain: x=a; repeat: x = (x*b)^2."""
    p = PRIMES[prime_id]
    x = a % p
    b %= p
    for _ in range(iters):
        if op == FIELD_OP_MUL:
            x = x * b % p
        elif op == FIELD_OP_SQR:
            x = x * x % p
        else:
            x = x * b % p
            x = x * x % p
    return x


# --------------------------------------------------------------------------------------
# Philox-4x32-10 (seed "forge-ec", stream = config id, counter = lane)
# --------------------------------------------------------------------------------------

PHILOX_SEED = 0x666F7267652D6563
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32(counter_lo: np.ndarray, stream: int, block: int, seed: int = PHILOX_SEED) -> np.ndarray:
    """counter = (lane_lo32, lane_hi32, block, stream); key = seed split in two words.
    Returns uint32 array [len(counter), 4]."""
    lane = np.asarray(counter_lo, dtype=np.uint64)
    c0 = lane & np.uint64(0xFFFFFFFF)
    c1 = lane >> np.uint64(32)
    c2 = np.full_like(c0, block)
    c3 = np.full_like(c0, stream)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & mask
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & mask
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def philox_bytes32(n: int, stream: int, first_lane: int = 0, sub: int = 0) -> np.ndarray:
    """n x 32 uniform bytes; lane i uses blocks 2*sub and 2*sub+1 of its counter."""
    lanes = np.arange(first_lane, first_lane + n, dtype=np.uint64)
    lo = philox4x32(lanes, stream, 2 * sub)
    hi = philox4x32(lanes, stream, 2 * sub + 1)
    words = np.concatenate([lo, hi], axis=1).astype("<u4")
    return words.view(np.uint8).reshape(n, 32)


def random_scalars_be(curve, n: int, stream: int, first_lane: int = 0, edge: bool = True, sub: int = 0) -> np.ndarray:
    """n x 32 big-endian scalars < order: value read big-endian, one conditional subtract of
    n (core lib.rs:326-332 `from_bytes_reduced` < n case); lanes 0..7 of the *global* range
    are the edge scalars {0,1,2,3,n-1,n-2,(n+1)/2,2^255 mod n}."""
    c = CURVES[curve]
    raw = philox_bytes32(n, stream, first_lane, sub)
    nb = np.frombuffer(c.n.to_bytes(32, "big"), dtype=np.uint8)
    # lexicographic compare raw >= n (vectorised)
    diff = raw.astype(np.int16) - nb.astype(np.int16)
    nz = diff != 0
    first = np.where(nz.any(axis=1), nz.argmax(axis=1), 0)
    ge = np.where(nz.any(axis=1), diff[np.arange(n), first] > 0, True)
    out = raw.copy()
    for i in np.nonzero(ge)[0]:
        v = int.from_bytes(raw[i].tobytes(), "big") - c.n
        out[i] = np.frombuffer(v.to_bytes(32, "big"), dtype=np.uint8)
    if edge:
        edges = [0, 1, 2, 3, c.n - 1, c.n - 2, (c.n + 1) // 2, (1 << 255) % c.n]
        for j, v in enumerate(edges):
            g = j - first_lane
            if 0 <= g < n:
                out[g] = np.frombuffer(v.to_bytes(32, "big"), dtype=np.uint8)
    return out


def random_field_le(prime_id: int, n: int, stream: int, sub: int = 0) -> np.ndarray:
    """n x 32 little-endian field elements in [0,p) (top bits masked then one conditional
    subtract -- uniform enough for a benchmark, exact for parity)."""
    p = PRIMES[prime_id]
    raw = philox_bytes32(n, stream, 0, sub).copy()
    if prime_id == 2:
        raw[:, 31] &= 0x7F
    out = raw
    vals = raw.view("<u8").reshape(n, 4)
    # rows that might be >= p: top limb all ones (cheap prefilter), fix exactly
    cand = np.nonzero(vals[:, 3] >= np.uint64((p >> 192) & 0xFFFFFFFFFFFFFFFF))[0]
    for i in cand:
        v = int.from_bytes(raw[i].tobytes(), "little")
        if v >= p:
            out[i] = np.frombuffer((v - p).to_bytes(32, "little"), dtype=np.uint8)
    return out


# ----------------------------------------------------------------------------------------------
# forge-ec's LITERAL P-256 code path as big-integer semantics (forge-ec-curves/src/p256.rs): what oracle/ref/
# oracle_ref_p256.c and FEC_SEM_FORGE_QUIRK replay limb by limb.  Add, Mul and pow are the field up to the corner
# cases of reduce_wide_p256 (its closing +-p loops wrap modulo 2^256, :672-696); Sub (:470-496) adds p through `+=`,
# which reduces it away again, so a < b gives a - b + 2^256; double (:1869-1912) is the a = 0 formula; Add
# (:1941-2006) starts with a projective equality test; multiply (:2120-2156) is MSB-first double-and-add.
# Reproduces the transliteration goldens of SURVEY.md Appendix A.2 (tests/golden/forge_ec_literal.json).
# ----------------------------------------------------------------------------------------------
p256lit_p = 2**256 - 2**224 + 2**192 + 2**96 - 1
p256lit_n = 0xffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551
p256lit_M = 2**256
def p256lit_reduce_wide(T):
    c = [(T >> (32*i)) & 0xFFFFFFFF for i in range(16)]
    a = [0]*8
    for i in range(8): a[i] += c[i]
    a[3]+=2*c[11]; a[4]+=2*c[12]; a[5]+=2*c[13]; a[6]+=2*c[14]; a[7]+=2*c[15]
    a[3]+=2*c[12]; a[4]+=2*c[13]; a[5]+=2*c[14]; a[6]+=2*c[15]
    a[0]+=c[8]; a[1]+=c[9]; a[2]+=c[10]; a[6]+=c[14]; a[7]+=c[15]
    a[0]+=c[9]; a[1]+=c[10]; a[2]+=c[11]; a[3]+=c[13]; a[4]+=c[14]; a[5]+=c[15]; a[6]+=c[13]; a[7]+=c[8]
    a[0]-=c[11]; a[1]-=c[12]; a[2]-=c[13]; a[6]-=c[8]; a[7]-=c[10]
    a[0]-=c[12]; a[1]-=c[13]; a[2]-=c[14]; a[3]-=c[15]; a[6]-=c[9]; a[7]-=c[11]
    a[0]-=c[13]; a[1]-=c[14]; a[2]-=c[15]; a[3]-=c[8]; a[4]-=c[9]; a[5]-=c[10]; a[7]-=c[12]
    a[0]-=c[14]; a[1]-=c[15]; a[3]-=c[9]; a[4]-=c[10]; a[5]-=c[11]; a[7]-=c[13]
    V = sum(a[i] << (32*i) for i in range(8))
    carry, L = V >> 256, V & (p256lit_M-1)       # floor semantics, as the arithmetic shifts give
    R = (L - carry*p256lit_p) % p256lit_M
    while R >= p256lit_p: R -= p256lit_p
    return R
def p256lit_mul(a,b): return p256lit_reduce_wide(a*b)
def p256lit_sqr(a): return p256lit_mul(a,a)
def p256lit_add(a,b): return (a+b) % p256lit_p
def p256lit_sub(a,b):
    if a < b: a = a % p256lit_p
    return (a - b) % p256lit_M
def p256lit_neg(a): return 0 if a == 0 else (p256lit_p - a) % p256lit_M
def p256lit_powl(a, e):
    r, base = 1, a
    for i in range(256):
        if (e >> i) & 1: r = p256lit_mul(r, base)
        base = p256lit_sqr(base)
    return r
def p256lit_inv(a): return p256lit_powl(a, p256lit_p-2)
p256lit_ID = (0,1,0)
def p256lit_dbl(P):
    x,y,z = P
    if z == 0: return p256lit_ID
    xx=p256lit_sqr(x); yy=p256lit_sqr(y); yyyy=p256lit_sqr(yy)
    xy2=p256lit_sqr(p256lit_add(x,yy)); t=p256lit_sub(p256lit_sub(xy2,xx),yyyy); d=p256lit_add(t,t)
    e=p256lit_mul(3,xx); ee=p256lit_sqr(e)
    x3=p256lit_sub(p256lit_sub(ee,d),d)
    y3=p256lit_sub(p256lit_mul(e,p256lit_sub(d,x3)), p256lit_mul(8,yyyy))
    z3=p256lit_add(y,y)
    if z != 1: z3 = p256lit_mul(z3, z)
    return (x3,y3,z3)
def p256lit_pt_eq(P,Q):
    if P[2]==0 and Q[2]==0: return True
    if P[2]==0 or Q[2]==0: return False
    z1z1=p256lit_sqr(P[2]); z2z2=p256lit_sqr(Q[2])
    u1=p256lit_mul(P[0],z2z2); u2=p256lit_mul(Q[0],z1z1)
    s1=p256lit_mul(p256lit_mul(P[1],Q[2]),z2z2); s2=p256lit_mul(p256lit_mul(Q[1],P[2]),z1z1)
    return u1==u2 and s1==s2
def p256lit_padd(P,Q):
    if P[2]==0: return Q
    if Q[2]==0: return P
    if p256lit_pt_eq(P,Q): return p256lit_dbl(P)
    z1z1=p256lit_sqr(P[2]); z2z2=p256lit_sqr(Q[2])
    u1=p256lit_mul(P[0],z2z2); u2=p256lit_mul(Q[0],z1z1)
    s1=p256lit_mul(p256lit_mul(P[1],Q[2]),z2z2); s2=p256lit_mul(p256lit_mul(Q[1],P[2]),z1z1)
    if u1==u2 and s1==p256lit_neg(s2): return p256lit_ID
    h=p256lit_sub(u2,u1); i=p256lit_sqr(p256lit_add(h,h)); j=p256lit_mul(h,i)
    r=p256lit_add(p256lit_sub(s2,s1),p256lit_sub(s2,s1)); v=p256lit_mul(u1,i)
    x3=p256lit_sub(p256lit_sub(p256lit_sub(p256lit_sqr(r),j),v),v)
    y3=p256lit_sub(p256lit_mul(r,p256lit_sub(v,x3)), p256lit_mul(p256lit_add(s1,s1),j))
    z3=p256lit_mul(p256lit_sub(p256lit_sub(p256lit_sqr(p256lit_add(P[2],Q[2])),z1z1),z2z2),h)
    return (x3,y3,z3)
def p256lit_multiply(P,k):
    if P[2]==0 or k==0: return p256lit_ID
    R=p256lit_ID
    for i in range(255,-1,-1):
        R=p256lit_dbl(R)
        if (k>>i)&1: R=p256lit_padd(R,P)
    return R
def p256lit_to_bytes(P):
    if P[2]==0: return bytes(33)
    zi=p256lit_inv(P[2]); zi2=p256lit_sqr(zi); zi3=p256lit_mul(zi2,zi)
    x=p256lit_mul(P[0],zi2); y=p256lit_mul(P[1],zi3)
    return bytes([3 if y&1 else 2])+x.to_bytes(32,'big')
p256lit_G=(0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5,1)
