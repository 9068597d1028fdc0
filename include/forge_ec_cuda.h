/* forge_ec_cuda.h -- C ABI of the B200 batched scalar-multiplication engine.
 *
 * This is the drop-in boundary for forge-ec's one data-parallel hot path.  The reference has no
 * FFI of its own (every crate is #![forbid(unsafe_code)]); the seam it offers is the `Curve`
 * trait, one call per scalar.  Each entry point below is the batched form of a reference call
 * chain (file:line relative to the forge-ec repository):
 *
 *   fec_batch_mul            n x { Scalar::from_bytes            secp256k1.rs:2271-2297 / p256.rs:1041-1055
 *                                  Curve::generator | from_affine secp256k1.rs:2608-2625, 1365-1373 / p256.rs:2092-2110
 *                                  Curve::multiply               secp256k1.rs:2635-2692 / p256.rs:2120-2156
 *                                  Curve::to_affine              secp256k1.rs:1342-1363 / p256.rs:1835-1857
 *                                  PointAffine::to_bytes[_with_format] / CompressedPoint::from_affine /
 *                                  UncompressedPoint::from_affine   secp256k1.rs:875-894, 1010-1067;
 *                                                                   forge-ec-encoding/src/point.rs:39-67, 186-211 }
 *                            with points_fmt = SEC1 compressed additionally
 *                                  CompressedPoint::to_affine / Curve::validate_point
 *                                                                   point.rs:70-144; forge-ec-core/src/lib.rs:905-923
 *   fec_batch_x25519         n x   x25519(scalar, u)             curve25519.rs:1624-1716  (RFC 7748)
 *   fec_batch_ed25519_pubkey n x   Ed25519Signature::derive_public_key   forge-ec-signature/src/eddsa.rs:450-509 (RFC 8032 5.1.5)
 *   fec_batch_point_decode   n x   CompressedPoint::to_affine / Curve::validate_point   point.rs:70-144; forge-ec-core/src/lib.rs:905-923
 *   fec_batch_nonce_commit   n x   the R = k*G step of EcdsaSignature / BipSchnorr / EdDSA sign   ecdsa.rs:108-131, schnorr.rs:372-385,
 *                                                                eddsa.rs:120-131
 *   fec_field_bench          n x   FieldElement::mul / square    secp256k1.rs:442-507, 634-713; p256.rs:498-534, 772-776;
 *                                                                curve25519.rs:227-314, 490-494
 *
 * Conventions
 *   - Every buffer is caller-owned, tightly packed, lane-major (lane i at offset i*record_size).
 *   - Host entry points take HOST pointers (pageable or pinned), shard the lanes in contiguous
 *     ranges over the context's devices, copy in, compute, copy out, and return when the result
 *     is in `out`.  `*_dev` entry points take DEVICE pointers valid on the context's device
 *     `dev_index` (16-byte aligned), enqueue on `stream` (a cudaStream_t, NULL = legacy default
 *     stream) and return without synchronising.
 *   - Return value: 0 on success, else a fec_err (whole-call failure).  Per-lane outcomes that
 *     depend on the data go to `status` (may be NULL): the device never branches on them.
 *   - A lane whose status is not FEC_LANE_OK has an all-zero output record, which is also the
 *     reference's encoding of the identity (point.rs:41-47).
 *   - There is no CPU fallback: without a usable CUDA device fec_ctx_create fails.
 *   - Threads: the host entry points of one context are serialised on an internal mutex (they share the context's pipeline
 *     buffers); several contexts may coexist.  The `*_dev` entry points take no lock: they only enqueue work on the caller's
 *     stream and may be called concurrently for different streams (the launch counter is atomic).  fec_ctx_stage_timing /
 *     fec_ctx_stage_ms are single-threaded measurement aids and must not overlap other calls on the same context.
 *   - `*_dev` calls never synchronise, with one exception: fec_msm_dev with n >= 2^16 points waits once for the stream (the
 *     bucket method reads back its largest bucket size to decide on the fallback), so it cannot be captured in a CUDA graph.
 *   - Scratch memory comes from a stream-ordered pool the library creates per device (cudaMallocFromPoolAsync; the
 *     process-wide default pool and its attributes are not touched).  The pool keeps its memory between calls and gives
 *     it back to the driver in fec_ctx_destroy.
 *   - Side channels: control flow never depends on secrets.  By default the fixed-base entry points (k*G, Ed25519 key
 *     generation, X25519 with the base point, nonce commitments) gather table entries by secret digits; the environment
 *     variable FORGE_EC_CT_TABLE=1 (read once per process) selects scanned 4-bit tables in shared memory for all of them
 *     at about a quarter of the throughput.
 */
#ifndef FORGE_EC_CUDA_H
#define FORGE_EC_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The library is linked with hidden default visibility: exactly the functions declared here are exported. */
#if defined(__GNUC__)
#define FEC_API __attribute__((visibility("default")))
#else
#define FEC_API
#endif

typedef struct fec_ctx fec_ctx;

typedef enum { FEC_SECP256K1 = 0, FEC_P256 = 1 } fec_curve;
typedef enum { FEC_PRIME_SECP256K1 = 0, FEC_PRIME_P256 = 1, FEC_PRIME_25519 = 2 } fec_prime;
typedef enum {
  FEC_OUT_SEC1_COMPRESSED = 0,   /* 33 B: 02|03 || X          (PointFormat::Compressed)   */
  FEC_OUT_SEC1_UNCOMPRESSED = 1, /* 65 B: 04 || X || Y        (PointFormat::Uncompressed) */
  FEC_OUT_AFFINE_XY_BE = 2,      /* 64 B: X || Y, big-endian  (FieldElement::to_bytes x2) */
  FEC_OUT_X_BE = 3,              /* 32 B: X, big-endian       (the ECDH shared secret, secp256k1.rs:1885-1905) */
  FEC_OUT_SEC1_HYBRID = 4        /* 65 B: 06|07 || X || Y     (PointFormat::Hybrid, forge-ec-core/src/lib.rs:536-543;
                                    secp256k1.rs:1048-1066, p256.rs:1704-1722) */
} fec_out;
typedef enum {
  FEC_KDF_NONE = 0,        /* the 32-byte shared secret itself (KeyExchange::derive_shared_secret)                       */
  FEC_KDF_HKDF_SHA256 = 1, /* Secp256k1::derive_key: RFC 5869 HKDF-SHA-256, salt = 32 zero bytes (secp256k1.rs:1846-1883) */
  FEC_KDF_XOR_INFO = 2     /* P256::derive_key: the reference's placeholder, secret XOR info (p256.rs:2314-2344)         */
} fec_kdf;
typedef enum { FEC_PT_AFFINE_XY_BE = 0 /* 64 B */, FEC_PT_SEC1_COMPRESSED = 1 /* 33 B */ } fec_points_fmt;
typedef enum {
  FEC_SEM_STD = 0,        /* SEC2 / FIPS 186-4 / RFC 7748 / RFC 8032 results (what the reference specifies) */
  FEC_SEM_FORGE_QUIRK = 1 /* replay of the reference's LITERAL limb code: secp256k1 (secp256k1.rs:442-507, 634-713 and the
                             256-step ladder :2635-2692) and P-256 (p256.rs:416-534, 544-704, the a = 0 doubling :1869-1912,
                             Add with its projective equality test :1941-2067, MSB-first double-and-add :2120-2156), whose
                             outputs are not SEC2 / FIPS 186-4 results.  k*G or affine X||Y points
                             only, anything else: FEC_ERR_UNSUPPORTED.  A parity artefact (BASELINE.json configs[0]), one
                             thread per lane, not a fast path.  Lane status: 1 scalar >= n, 3 identity (all-zero record in
                             both cases), no point validation (the reference does none). */
} fec_sem;
typedef enum { FEC_FIELD_MUL = 0, FEC_FIELD_SQR = 1, FEC_FIELD_MULSQR = 2 } fec_field_op;
typedef enum {
  FEC_NONCE_ECDSA = 0,  /* R = k G, r = x(R) as a scalar            (forge-ec-signature/src/ecdsa.rs:108-131)  */
  FEC_NONCE_BIP340 = 1, /* R = k G, x-only R, k <- n - k if y odd   (schnorr.rs:372-385; secp256k1 only)       */
  FEC_NONCE_EDDSA = 2   /* r = 64-byte hash value mod L, R = [r]B   (eddsa.rs:120-131, 322-330; RFC 8032 5.1.6) */
} fec_nonce_mode;

typedef enum {
  FEC_OK = 0,
  FEC_ERR_INVALID_ARG = 1, /* NULL / bad enum / misaligned device pointer  -> Error::InvalidEncoding */
  FEC_ERR_CUDA = 2,        /* a CUDA runtime call failed                  -> Error::GenericError    */
  FEC_ERR_NO_DEVICE = 3,   /* no usable CUDA device                       -> Error::UnsupportedOperation */
  FEC_ERR_UNSUPPORTED = 4, /* semantics / format not available            -> Error::UnsupportedOperation */
  FEC_ERR_OOM = 5          /* host or device allocation failed            -> Error::GenericError    */
} fec_err;

/* per-lane status bytes */
enum {
  FEC_LANE_OK = 0,
  FEC_LANE_SCALAR_RANGE = 1,  /* scalar >= group order      -> Error::InvalidScalar   */
  FEC_LANE_POINT_INVALID = 2, /* not on curve / bad encoding -> Error::PointNotOnCurve / InvalidEncoding */
  FEC_LANE_INFINITY = 3       /* result is the identity (k = 0)                          */
};

/* device_ids == NULL or n_devices <= 0: use the current CUDA device only. */
FEC_API int fec_ctx_create(const int* device_ids, int n_devices, fec_ctx** out);
FEC_API void fec_ctx_destroy(fec_ctx* ctx);
FEC_API int fec_ctx_device_count(const fec_ctx* ctx);
/* number of kernels this context has launched since creation (for bench.py's gpu_launches) */
FEC_API uint64_t fec_ctx_launch_count(const fec_ctx* ctx);
FEC_API const char* fec_strerror(int err);
/* last CUDA error string seen by this context (thread-unsafe diagnostic) */
FEC_API const char* fec_ctx_last_cuda_error(const fec_ctx* ctx);

/* Sharding rule used by the host entry points (and by bench.py across ranks): shard `index` of `n_shards` owns
 * the contiguous lanes [first, first+count); cuts are multiples of 1024 lanes.  Needs no device. */
FEC_API void fec_shard_range(size_t n, int n_shards, int index, size_t* first, size_t* count);

/* k_i*G when points == NULL, else k_i*P_i.
 * scalars: n x 32 B big-endian (Scalar::from_bytes layout); points: n x 64 B (X||Y big-endian) or
 * n x 33 B SEC1 compressed; out: n x {33,65,64,32,65} B by out_fmt; status: n x u8 or NULL. */
FEC_API int fec_batch_mul(fec_ctx* ctx, fec_curve curve, fec_sem sem, const uint8_t* scalars, const uint8_t* points,
                  fec_points_fmt points_fmt, size_t n, fec_out out_fmt, uint8_t* out, uint8_t* status);
FEC_API int fec_batch_mul_dev(fec_ctx* ctx, int dev_index, fec_curve curve, fec_sem sem, const uint8_t* d_scalars,
                      const uint8_t* d_points, fec_points_fmt points_fmt, size_t n, fec_out out_fmt, uint8_t* d_out,
                      uint8_t* d_status, void* stream);

/* Standalone point decoding and validation: n x { CompressedPoint::to_affine (forge-ec-encoding/src/point.rs:70-144) |
 * PointAffine::new + is_on_curve; Curve::validate_point (forge-ec-core/src/lib.rs:905-923: on these cofactor-1 curves membership
 * of the curve is the whole check) }.  points: n x 33 B SEC1 compressed or n x 64 B X || Y; out: n records of out_fmt (so the same
 * call also converts between the encodings); status: 0 valid, 2 rejected (bad prefix, coordinate >= p, no square root, not on
 * the curve), 3 the identity record (prefix 0x00 / all-zero X || Y); rejected and identity lanes give all-zero records. */
FEC_API int fec_batch_point_decode(fec_ctx* ctx, fec_curve curve, const uint8_t* points, fec_points_fmt points_fmt, size_t n,
                                   fec_out out_fmt, uint8_t* out, uint8_t* status);
FEC_API int fec_batch_point_decode_dev(fec_ctx* ctx, int dev_index, fec_curve curve, const uint8_t* d_points, fec_points_fmt points_fmt,
                                       size_t n, fec_out out_fmt, uint8_t* d_out, uint8_t* d_status, void* stream);

/* Batched ECDH: shared_i = x(k_i * P_i), 32 bytes big-endian, then the reference's key derivation.
 * kdf == FEC_KDF_NONE: out = n x 32 B shared secrets; otherwise out = n x out_len bytes (1 <= out_len <= 64, info_len <= 256;
 * `info` is host memory in both variants).  A lane whose scalar is out of range, whose point is invalid or whose result is
 * the identity gets an all-zero record and its status byte (the reference returns Err for it). */
FEC_API int fec_batch_ecdh(fec_ctx* ctx, fec_curve curve, const uint8_t* scalars, const uint8_t* points, fec_points_fmt points_fmt, size_t n,
                   fec_kdf kdf, const uint8_t* info, size_t info_len, size_t out_len, uint8_t* out, uint8_t* status);
FEC_API int fec_batch_ecdh_dev(fec_ctx* ctx, int dev_index, fec_curve curve, const uint8_t* d_scalars, const uint8_t* d_points,
                       fec_points_fmt points_fmt, size_t n, fec_kdf kdf, const uint8_t* info, size_t info_len, size_t out_len,
                       uint8_t* d_out, uint8_t* d_status, void* stream);

/* Multi-scalar multiplication sum_i k_i * P_i (Curve::multi_scalar_multiply, forge-ec-core/src/lib.rs:934-951):
 * the batched k*P kernels followed by a device-wide reduction.  out: ONE record of out_fmt; status: ONE byte -- 0 ok, 1 / 2 some
 * lane had an out-of-range scalar / an invalid point (out is all-zero), 3 the sum is the identity.  n >= 1.  The host variant
 * sums contiguous lane ranges on every device of the context and adds the per-device sums on the first one. */
FEC_API int fec_msm(fec_ctx* ctx, fec_curve curve, const uint8_t* scalars, const uint8_t* points, fec_points_fmt points_fmt, size_t n,
            fec_out out_fmt, uint8_t* out, uint8_t* status);
FEC_API int fec_msm_dev(fec_ctx* ctx, int dev_index, fec_curve curve, const uint8_t* d_scalars, const uint8_t* d_points,
                fec_points_fmt points_fmt, size_t n, fec_out out_fmt, uint8_t* d_out, uint8_t* d_status, void* stream);

/* Signing-side nonce commitment: the fixed-base multiplication of a signature and the scalar that goes with it.
 *   FEC_NONCE_ECDSA   nonces n x 32 B big-endian k;  out_R n x 33 B SEC1 compressed R = k G;  out_r n x 32 B big-endian r = x(R).
 *                     status: 1 if k >= n or x(R) >= n (where the reference's Scalar::from_bytes fails, ecdsa.rs:116-121), 3 if k = 0
 *                     or r = 0 (ecdsa.rs:125-129); records of such lanes are all zero.
 *   FEC_NONCE_BIP340  (curve must be FEC_SECP256K1) nonces n x 32 B big-endian k;  out_R n x 32 B x(R);  out_r n x 32 B big-endian
 *                     k' = k if y(R) is even, n - k otherwise (the nonce the signer continues with).  status as above (1: k >= n, 3: k = 0).
 *   FEC_NONCE_EDDSA   (curve ignored) nonces n x 64 B: the SHA-512 output r = H(prefix || M) as a little-endian integer;
 *                     out_r n x 32 B little-endian r mod L;  out_R n x 32 B RFC 8032 encoding of [r mod L]B.  status all zero.
 * status may be NULL. */
FEC_API int fec_batch_nonce_commit(fec_ctx* ctx, fec_curve curve, fec_nonce_mode mode, const uint8_t* nonces, size_t n, uint8_t* out_R,
                           uint8_t* out_r, uint8_t* status);
FEC_API int fec_batch_nonce_commit_dev(fec_ctx* ctx, int dev_index, fec_curve curve, fec_nonce_mode mode, const uint8_t* d_nonces, size_t n,
                               uint8_t* d_out_R, uint8_t* d_out_r, uint8_t* d_status, void* stream);

/* RFC 7748 X25519: scalars n x 32 B (clamped inside), u n x 32 B little-endian (NULL = base point 9),
 * out n x 32 B little-endian. */
FEC_API int fec_batch_x25519(fec_ctx* ctx, const uint8_t* scalars, const uint8_t* u, size_t n, uint8_t* out);
FEC_API int fec_batch_x25519_dev(fec_ctx* ctx, int dev_index, const uint8_t* d_scalars, const uint8_t* d_u, size_t n,
                         uint8_t* d_out, void* stream);
/* The same with a semantics selector: FEC_SEM_FORGE_QUIRK replays the reference's LITERAL x25519() (curve25519.rs:1624-1716 on its
 * own FieldElement, :50-494: limb arithmetic without carries, big-endian byte codec, A instead of a24, a hard-coded answer for
 * the scalar 2), whose outputs are not RFC 7748 results.  A parity artefact, one thread per lane. */
FEC_API int fec_batch_x25519_sem(fec_ctx* ctx, fec_sem sem, const uint8_t* scalars, const uint8_t* u, size_t n, uint8_t* out);
FEC_API int fec_batch_x25519_sem_dev(fec_ctx* ctx, int dev_index, fec_sem sem, const uint8_t* d_scalars, const uint8_t* d_u, size_t n,
                             uint8_t* d_out, void* stream);

/* RFC 8032 5.1.5 public key.  input_is_seed != 0: 32-byte seeds (SHA-512 + clamp on device);
 * input_is_seed == 0: 32-byte little-endian scalars used as they are.  out n x 32 B (y || sign x). */
FEC_API int fec_batch_ed25519_pubkey(fec_ctx* ctx, const uint8_t* seeds_or_scalars, int input_is_seed, size_t n,
                             uint8_t* out);
FEC_API int fec_batch_ed25519_pubkey_dev(fec_ctx* ctx, int dev_index, const uint8_t* d_in, int input_is_seed, size_t n,
                                 uint8_t* d_out, void* stream);
/* The same with a semantics selector: FEC_SEM_FORGE_QUIRK (seeds only, else FEC_ERR_UNSUPPORTED) replays the reference's LITERAL
 * Ed25519Signature::derive_public_key (eddsa.rs:450-509 on ed25519.rs: the second half of the hash parsed big-endian, the
 * fabricated base point, d for 2d, a hard-coded key for seeds starting with 0x9d, 32 bytes of a 33-byte SEC1-style record), whose
 * outputs are not RFC 8032 keys.  A parity artefact, one thread per lane. */
FEC_API int fec_batch_ed25519_pubkey_sem(fec_ctx* ctx, fec_sem sem, const uint8_t* seeds_or_scalars, int input_is_seed, size_t n, uint8_t* out);
FEC_API int fec_batch_ed25519_pubkey_sem_dev(fec_ctx* ctx, int dev_index, fec_sem sem, const uint8_t* d_in, int input_is_seed, size_t n,
                                     uint8_t* d_out, void* stream);

/* Field microbenchmark (config 2).  a, b, out: lanes x 32 B little-endian residues in [0,p).
 * Per lane: x = a; repeat iters times { MUL: x = x*b;  SQR: x = x^2;  MULSQR: x = (x*b)^2 }; out = x.
 * *seconds (may be NULL) receives the kernel time measured with CUDA events (host variant only). */
FEC_API int fec_field_bench(fec_ctx* ctx, fec_prime prime, fec_field_op op, size_t lanes, int iters, const uint8_t* a,
                    const uint8_t* b, uint8_t* out, double* seconds);
FEC_API int fec_field_bench_dev(fec_ctx* ctx, int dev_index, fec_prime prime, fec_field_op op, size_t lanes, int iters,
                        const uint8_t* d_a, const uint8_t* d_b, uint8_t* d_out, void* stream);

/* Roofline calibration: a register-resident loop of independent 32x32->64 multiply-adds.
 * variant 0 = IMAD.WIDE.U32 (one instruction per IMUL32), 1 = mad.lo + mad.hi (two IMAD per IMUL32),
 * 2 = carry-chained IMAD.WIDE.X as used by the field code.
 * Returns multiply-add *instructions* per second over the whole device. */
FEC_API int fec_imad_peak(fec_ctx* ctx, int dev_index, int variant, double* instr_per_sec);

/* Measurement aid: while enabled, fec_batch_mul_dev(FEC_SECP256K1, points != NULL) records CUDA events between its three
 * kernels (prepare / ladder / finish).  fec_ctx_stage_ms waits for them, returns the summed milliseconds per stage and
 * the number of timed chunks, and forgets them. */
FEC_API int fec_ctx_stage_timing(fec_ctx* ctx, int enable);
FEC_API int fec_ctx_stage_ms(fec_ctx* ctx, double ms[3], int* chunks);

/* Synthetic benchmark inputs: Philox-4x32-10, key "forge-ec", counter = (lane, block, stream_id);
 * lane i of the call is global lane first_lane + i and receives 32 bytes (blocks 2*sub, 2*sub+1).
 * mode 0: raw bytes; mode 1 / 2: big-endian scalar below the secp256k1 / P-256 group order (what
 * Scalar::random -> from_bytes_reduced yields, secp256k1.rs:2256-2263, forge-ec-core/src/lib.rs:326-332);
 * edge != 0 overwrites global lanes 0..7 with {0,1,2,3,n-1,n-2,(n+1)/2,2^255 mod n}. */
FEC_API int fec_synth_dev(fec_ctx* ctx, int dev_index, int mode, uint32_t stream_id, uint32_t sub, uint64_t first_lane, int edge,
                  size_t n, uint8_t* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FORGE_EC_CUDA_H */
