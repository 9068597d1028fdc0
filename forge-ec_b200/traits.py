"""Mirror of the forge-ec-core trait surface for the batched path (value types are plain bytes).

reference                                             here
---------------------------------------------------   ------------------------------------------
Scalar::from_bytes (secp256k1.rs:2271, p256.rs:1041)   Scalar.from_bytes(b) -> Scalar (32 B BE, < n)
Curve::generator / multiply / to_affine (lib.rs:801-832) Secp256k1.generator(), .multiply(P, k), .to_affine
PointAffine::to_bytes[_with_format] (lib.rs:606-632)   AffinePoint.to_bytes(), .to_bytes_with_format
(new) BatchMul::batch_mul                              Secp256k1.batch_mul(points, scalars)
x25519(scalar,u) (curve25519.rs:1624)                  x25519(scalar, u); batch_x25519(scalars, us)
Ed25519Signature::derive_public_key (eddsa.rs:450)     Ed25519Signature.derive_public_key(seed)
KeyExchange::derive_shared_secret / derive_key          Secp256k1.derive_shared_secret(k, P), .derive_key(secret, info, n),
  (secp256k1.rs:1846-1905, p256.rs:2281-2344)            .batch_ecdh(scalars, points, info, out_len)   (new, batched)

Every call goes to the GPU through the C ABI; single-point calls are batches of one.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import engine as _e

_DEFAULT: Optional[_e.Engine] = None


def default_engine() -> _e.Engine:
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = _e.Engine()
    return _DEFAULT


@dataclass(frozen=True)
class Scalar:
    """32-byte big-endian scalar (the layout of Scalar::to_bytes / from_bytes)."""
    be: bytes

    @staticmethod
    def from_bytes(b: bytes) -> "Scalar":
        if len(b) != 32:
            raise ValueError("InvalidEncoding: scalar must be 32 bytes")
        return Scalar(bytes(b))

    @staticmethod
    def from_int(v: int) -> "Scalar":
        return Scalar(v.to_bytes(32, "big"))

    def to_bytes(self) -> bytes:
        return self.be


@dataclass(frozen=True)
class AffinePoint:
    """Affine point as X || Y big-endian (64 B); ``infinity`` mirrors the reference's flag."""
    xy: bytes
    infinity: bool = False

    def x(self) -> bytes:
        return self.xy[:32]

    def y(self) -> bytes:
        return self.xy[32:]

    def is_identity(self) -> bool:
        return self.infinity

    def to_bytes(self) -> bytes:
        """SEC1 compressed, identity -> 33 zero bytes (secp256k1.rs:875-894, point.rs:39-67)."""
        return self.to_bytes_with_format("compressed")

    def to_bytes_with_format(self, fmt: str) -> bytes:
        if fmt == "compressed":
            return bytes(33) if self.infinity else bytes([2 | (self.xy[63] & 1)]) + self.xy[:32]
        if fmt == "uncompressed":
            return bytes(65) if self.infinity else b"\x04" + self.xy
        if fmt == "hybrid":  # PointFormat::Hybrid (forge-ec-core/src/lib.rs:536-543, secp256k1.rs:1048-1066)
            return bytes(65) if self.infinity else bytes([6 | (self.xy[63] & 1)]) + self.xy
        raise ValueError("UnsupportedOperation: format")


class _WeierstrassCurve:
    CURVE_ID = -1

    @classmethod
    def generator(cls) -> Optional[AffinePoint]:
        return None  # marker: multiply(None, k) is the fixed-base path

    @classmethod
    def batch_mul(cls, points: Optional[Sequence[AffinePoint]], scalars: Sequence[Scalar], engine: Optional[_e.Engine] = None):
        """BatchMul::batch_mul: k_i*G when points is None, else k_i*P_i.  Returns AffinePoints.
        Raises ValueError (Error::InvalidScalar / PointNotOnCurve) if any lane is rejected."""
        eng = engine or default_engine()
        n = len(scalars)
        s = np.frombuffer(b"".join(k.be for k in scalars), dtype=np.uint8)
        p = None if points is None else np.frombuffer(b"".join(q.xy for q in points), dtype=np.uint8)
        out, st = eng.batch_mul(cls.CURVE_ID, s, p, n=n, out_fmt=_e.OUT_AFFINE_XY_BE)
        if np.any(st == _e.LANE_SCALAR_RANGE):
            raise ValueError("InvalidScalar")
        if np.any(st == _e.LANE_POINT_INVALID):
            raise ValueError("PointNotOnCurve")
        return [AffinePoint(out[i].tobytes(), bool(st[i] == _e.LANE_INFINITY)) for i in range(n)]

    @classmethod
    def multiply(cls, point: Optional[AffinePoint], scalar: Scalar, engine: Optional[_e.Engine] = None) -> AffinePoint:
        """Curve::multiply followed by Curve::to_affine (secp256k1.rs:2635-2692 + 1342-1363)."""
        return cls.batch_mul(None if point is None else [point], [scalar], engine)[0]

    @staticmethod
    def to_affine(p: AffinePoint) -> AffinePoint:
        return p  # results already come back normalised (batch inversion on the device)

    @classmethod
    def multi_scalar_multiply(cls, points: Sequence[AffinePoint], scalars: Sequence[Scalar], engine: Optional[_e.Engine] = None) -> AffinePoint:
        """Curve::multi_scalar_multiply (forge-ec-core/src/lib.rs:934-951): sum of scalars[i] * points[i]; the identity for
        mismatched or empty inputs, as the reference."""
        if len(points) != len(scalars) or not points:
            return AffinePoint(bytes(64), True)
        eng = engine or default_engine()
        s = np.frombuffer(b"".join(k.be for k in scalars), dtype=np.uint8)
        p = np.frombuffer(b"".join(q.xy for q in points), dtype=np.uint8)
        rec, st = eng.msm(cls.CURVE_ID, s, p, n=len(scalars), out_fmt=_e.OUT_AFFINE_XY_BE)
        if st == _e.LANE_SCALAR_RANGE:
            raise ValueError("InvalidScalar")
        if st == _e.LANE_POINT_INVALID:
            raise ValueError("PointNotOnCurve")
        return AffinePoint(rec, st == _e.LANE_INFINITY)

    @classmethod
    def batch_decode(cls, records: Sequence[bytes], engine: Optional[_e.Engine] = None):
        """CompressedPoint::to_affine for a batch of 33-byte SEC1 records (forge-ec-encoding/src/point.rs:70-144), validated as
        Curve::validate_point does (forge-ec-core/src/lib.rs:905-923).  Raises ValueError("InvalidEncoding") for a rejected record."""
        eng = engine or default_engine()
        out, st = eng.batch_point_decode(cls.CURVE_ID, np.frombuffer(b"".join(records), dtype=np.uint8), n=len(records))
        if np.any(st == _e.LANE_POINT_INVALID):
            raise ValueError("InvalidEncoding")
        return [AffinePoint(out[i].tobytes(), bool(st[i] == _e.LANE_INFINITY)) for i in range(len(records))]

    # ---- KeyExchange (secp256k1.rs:1843-1906, p256.rs:2278-2345) ------------------------------------------
    KDF = _e.KDF_HKDF_SHA256

    @classmethod
    def batch_ecdh(cls, scalars: Sequence[Scalar], points: Sequence[AffinePoint], info: Optional[bytes] = None, out_len: int = 32,
                   engine: Optional[_e.Engine] = None):
        """derive_shared_secret (info is None) or derive_key(derive_shared_secret(..), info, out_len) for every pair.
        Raises ValueError where the reference returns Err (scalar out of range, invalid public key, identity)."""
        eng = engine or default_engine()
        n = len(scalars)
        s = np.frombuffer(b"".join(k.be for k in scalars), dtype=np.uint8)
        p = np.frombuffer(b"".join(q.xy for q in points), dtype=np.uint8)
        out, st = eng.batch_ecdh(cls.CURVE_ID, s, p, n=n, kdf=_e.KDF_NONE if info is None else cls.KDF, info=info or b"",
                                 out_len=out_len)
        if np.any(st == _e.LANE_SCALAR_RANGE):
            raise ValueError("InvalidScalar")
        if np.any(st == _e.LANE_POINT_INVALID):
            raise ValueError("InvalidPublicKey")
        if np.any(st == _e.LANE_INFINITY):
            raise ValueError("KeyExchangeError")
        return [out[i].tobytes() for i in range(n)]

    @classmethod
    def derive_shared_secret(cls, private_key: Scalar, public_key: AffinePoint, engine: Optional[_e.Engine] = None) -> bytes:
        return cls.batch_ecdh([private_key], [public_key], None, 32, engine)[0]


def batch_nonce_commit(curve: int, mode: int, nonces: Sequence[bytes], engine: Optional[_e.Engine] = None):
    """The fixed-base step of `sign` for a batch of nonces: ECDSA r = x(k G) (ecdsa.rs:108-131), BIP-340
    x-only R with k negated when y(R) is odd (schnorr.rs:372-385), EdDSA R = [r mod L]B from 64-byte hash values
    (eddsa.rs:120-131, RFC 8032 5.1.6).  Returns a list of (R bytes, scalar bytes); raises ValueError where the reference
    returns Err (nonce out of range / zero, x(R) not a scalar)."""
    eng = engine or default_engine()
    R, r, st = eng.batch_nonce_commit(curve, mode, np.frombuffer(b"".join(nonces), dtype=np.uint8), n=len(nonces))
    if np.any(st == _e.LANE_SCALAR_RANGE):
        raise ValueError("InvalidScalar")
    if np.any(st == _e.LANE_INFINITY):
        raise ValueError("InvalidSignature")
    return [(R[i].tobytes(), r[i].tobytes()) for i in range(len(nonces))]


class Secp256k1(_WeierstrassCurve):
    CURVE_ID = _e.CURVE_SECP256K1


class P256(_WeierstrassCurve):
    CURVE_ID = _e.CURVE_P256
    KDF = _e.KDF_XOR_INFO  # the reference's P-256 derive_key is a placeholder (p256.rs:2314-2344)


def batch_x25519(scalars: Sequence[bytes], us: Optional[Sequence[bytes]] = None, engine: Optional[_e.Engine] = None):
    eng = engine or default_engine()
    n = len(scalars)
    s = np.frombuffer(b"".join(scalars), dtype=np.uint8)
    u = None if us is None else np.frombuffer(b"".join(us), dtype=np.uint8)
    out = eng.batch_x25519(s, u, n=n)
    return [out[i].tobytes() for i in range(n)]


def x25519(scalar: bytes, u: bytes, engine: Optional[_e.Engine] = None) -> bytes:
    """x25519(&scalar, &u) (curve25519.rs:1624-1716), RFC 7748 semantics."""
    return batch_x25519([scalar], [u], engine)[0]


class Curve25519:
    @staticmethod
    def generator() -> bytes:
        return (9).to_bytes(32, "little")

    @staticmethod
    def multiply(u: bytes, scalar: bytes, engine: Optional[_e.Engine] = None) -> bytes:
        return x25519(scalar, u, engine)


class Ed25519Signature:
    @staticmethod
    def derive_public_key(seed: bytes, engine: Optional[_e.Engine] = None) -> bytes:
        """RFC 8032 5.1.5 (what eddsa.rs:450-509 specifies)."""
        eng = engine or default_engine()
        return eng.batch_ed25519_pubkey(np.frombuffer(seed, dtype=np.uint8), True, n=1)[0].tobytes()

    @staticmethod
    def batch_derive_public_key(seeds: Sequence[bytes], engine: Optional[_e.Engine] = None):
        eng = engine or default_engine()
        out = eng.batch_ed25519_pubkey(np.frombuffer(b"".join(seeds), dtype=np.uint8), True, n=len(seeds))
        return [out[i].tobytes() for i in range(len(seeds))]
