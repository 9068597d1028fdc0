"""forge-ec_b200 -- host side of the B200 batched scalar-multiplication engine.

The product is ``lib/libforge_ec_cuda.so`` (hand-written sm_100a kernels behind the C ABI of
``include/forge_ec_cuda.h``).  This package is the thin Python binding used by the tests and
``bench.py`` and a mirror of the forge-ec-core trait surface for this path
(``Curve::multiply`` / ``to_affine`` / ``PointAffine::to_bytes`` / ``x25519`` /
``Ed25519Signature::derive_public_key``) with the new ``batch_mul`` / ``batch_x25519`` entry points.

There is no CPU fallback: importing works anywhere (so that symbol checks can run without a
GPU), but creating an :class:`Engine` without a CUDA device raises :class:`FecError`.

The directory name contains a hyphen; import it with
``importlib.import_module("forge-ec_b200")``.
"""
from .engine import (  # noqa: F401
    CURVE_P256,
    CURVE_SECP256K1,
    Engine,
    FecError,
    FIELD_MUL,
    FIELD_MULSQR,
    FIELD_SQR,
    LANE_INFINITY,
    LANE_OK,
    LANE_POINT_INVALID,
    LANE_SCALAR_RANGE,
    NONCE_BIP340,
    NONCE_ECDSA,
    NONCE_EDDSA,
    KDF_HKDF_SHA256,
    KDF_NONE,
    KDF_XOR_INFO,
    OUT_AFFINE_XY_BE,
    OUT_X_BE,
    OUT_SEC1_COMPRESSED,
    OUT_SEC1_HYBRID,
    OUT_SEC1_UNCOMPRESSED,
    OUT_SIZE,
    PRIME_25519,
    PRIME_P256,
    PRIME_SECP256K1,
    PT_AFFINE_XY_BE,
    PT_SEC1_COMPRESSED,
    SEM_FORGE_QUIRK,
    SEM_STD,
    EXPORTED_SYMBOLS,
    library_path,
    load_library,
    shard_range,
)
from .traits import (  # noqa: F401
    AffinePoint,
    Curve25519,
    Ed25519Signature,
    P256,
    Scalar,
    Secp256k1,
    batch_nonce_commit,
    batch_x25519,
    default_engine,
    x25519,
)
