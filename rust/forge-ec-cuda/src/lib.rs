//! forge-ec-cuda: the `extern "C"` shim and the `BatchMul` extension trait.
//!
//! A new crate because every existing forge-ec crate is `#![forbid(unsafe_code)]`.  It keeps the
//! forge-ec-core trait surface (`Curve`, `Scalar`, `PointAffine`) as the API and adds the batched
//! entry points; byte layouts are the reference's own (`Scalar::to_bytes` big-endian 32 B,
//! `FieldElement::to_bytes` big-endian 32 B, SEC1 compressed 33 B).
//!
//! NOT COMPILED in the build image (no cargo/rustc there); kept in sync with
//! `include/forge_ec_cuda.h` by review.
#![allow(non_camel_case_types)]

use forge_ec_core::{Curve, Error, FieldElement, PointAffine, Result, Scalar};
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct fec_ctx {
    _private: [u8; 0],
}

pub const FEC_SECP256K1: c_int = 0;
pub const FEC_P256: c_int = 1;
pub const FEC_SEM_STD: c_int = 0;
pub const FEC_PT_AFFINE_XY_BE: c_int = 0;
pub const FEC_PT_SEC1_COMPRESSED: c_int = 1;
pub const FEC_OUT_SEC1_COMPRESSED: c_int = 0;
pub const FEC_OUT_SEC1_UNCOMPRESSED: c_int = 1;
pub const FEC_OUT_AFFINE_XY_BE: c_int = 2;
pub const FEC_OUT_X_BE: c_int = 3;
pub const FEC_OUT_SEC1_HYBRID: c_int = 4;
pub const FEC_KDF_NONE: c_int = 0;
pub const FEC_KDF_HKDF_SHA256: c_int = 1;
pub const FEC_KDF_XOR_INFO: c_int = 2;
pub const FEC_NONCE_ECDSA: c_int = 0;
pub const FEC_NONCE_BIP340: c_int = 1;
pub const FEC_NONCE_EDDSA: c_int = 2;
pub const FEC_LANE_OK: u8 = 0;
pub const FEC_LANE_SCALAR_RANGE: u8 = 1;
pub const FEC_LANE_POINT_INVALID: u8 = 2;
pub const FEC_LANE_INFINITY: u8 = 3;

extern "C" {
    pub fn fec_ctx_create(device_ids: *const c_int, n_devices: c_int, out: *mut *mut fec_ctx) -> c_int;
    pub fn fec_ctx_destroy(ctx: *mut fec_ctx);
    pub fn fec_ctx_device_count(ctx: *const fec_ctx) -> c_int;
    pub fn fec_strerror(err: c_int) -> *const c_char;
    pub fn fec_batch_mul(
        ctx: *mut fec_ctx, curve: c_int, sem: c_int, scalars: *const u8, points: *const u8, points_fmt: c_int, n: usize,
        out_fmt: c_int, out: *mut u8, status: *mut u8,
    ) -> c_int;
    pub fn fec_batch_mul_dev(
        ctx: *mut fec_ctx, dev_index: c_int, curve: c_int, sem: c_int, d_scalars: *const u8, d_points: *const u8,
        points_fmt: c_int, n: usize, out_fmt: c_int, d_out: *mut u8, d_status: *mut u8, stream: *mut c_void,
    ) -> c_int;
    pub fn fec_batch_ecdh(
        ctx: *mut fec_ctx, curve: c_int, scalars: *const u8, points: *const u8, points_fmt: c_int, n: usize, kdf: c_int,
        info: *const u8, info_len: usize, out_len: usize, out: *mut u8, status: *mut u8,
    ) -> c_int;
    pub fn fec_msm(
        ctx: *mut fec_ctx, curve: c_int, scalars: *const u8, points: *const u8, points_fmt: c_int, n: usize, out_fmt: c_int,
        out: *mut u8, status: *mut u8,
    ) -> c_int;
    pub fn fec_batch_point_decode(
        ctx: *mut fec_ctx, curve: c_int, points: *const u8, points_fmt: c_int, n: usize, out_fmt: c_int, out: *mut u8, status: *mut u8,
    ) -> c_int;
    pub fn fec_batch_nonce_commit(
        ctx: *mut fec_ctx, curve: c_int, mode: c_int, nonces: *const u8, n: usize, out_r_point: *mut u8, out_r_scalar: *mut u8,
        status: *mut u8,
    ) -> c_int;
    pub fn fec_batch_x25519(ctx: *mut fec_ctx, scalars: *const u8, u: *const u8, n: usize, out: *mut u8) -> c_int;
    /// sem = 1 (FEC_SEM_FORGE_QUIRK): the reference's literal x25519(), a parity artefact
    /// sem = 1 (FEC_SEM_FORGE_QUIRK, seeds only): the reference's literal derive_public_key, a parity artefact
    pub fn fec_batch_ed25519_pubkey_sem(ctx: *mut fec_ctx, sem: c_int, seeds: *const u8, input_is_seed: c_int, n: usize, out: *mut u8) -> c_int;
    pub fn fec_batch_x25519_sem(ctx: *mut fec_ctx, sem: c_int, scalars: *const u8, u: *const u8, n: usize, out: *mut u8) -> c_int;
    pub fn fec_batch_ed25519_pubkey(ctx: *mut fec_ctx, seeds_or_scalars: *const u8, input_is_seed: c_int, n: usize, out: *mut u8) -> c_int;
}

/// Owns one `fec_ctx` (devices, streams, comb tables).  `Send` but not `Sync`-shared without the
/// library's own mutex: calls on one context serialise inside the library.
pub struct Engine {
    ctx: *mut fec_ctx,
}
unsafe impl Send for Engine {}

fn map_err(code: c_int) -> Error {
    match code {
        1 => Error::InvalidEncoding,
        3 | 4 => Error::UnsupportedOperation,
        _ => Error::GenericError,
    }
}

impl Engine {
    /// `devices = &[]` uses the current CUDA device.  Fails (no CPU fallback) when there is none.
    pub fn new(devices: &[i32]) -> Result<Self> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe {
            if devices.is_empty() { fec_ctx_create(std::ptr::null(), 0, &mut ctx) } else { fec_ctx_create(devices.as_ptr(), devices.len() as c_int, &mut ctx) }
        };
        if rc != 0 { Err(map_err(rc)) } else { Ok(Self { ctx }) }
    }
}
impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { fec_ctx_destroy(self.ctx) }
    }
}

/// Curves the engine knows; implemented for the two Weierstrass curves of forge-ec-curves.
pub trait CudaCurve: Curve {
    const CURVE_ID: c_int;
    /// rebuild an affine point from `X || Y` big-endian (64 B) and the identity flag
    fn affine_from_xy(xy: &[u8; 64], infinity: bool) -> Self::PointAffine;
}

/// The new entry point: `k_i * G` (`points = None`) or `k_i * P_i`, one GPU batch.
/// Replaces n sequential `Curve::multiply` + `Curve::to_affine` calls
/// (forge-ec-curves/src/secp256k1.rs:2635-2692 + 1342-1363; p256.rs:2120-2156 + 1835-1857).
pub trait BatchMul: CudaCurve {
    fn batch_mul(engine: &Engine, points: Option<&[Self::PointAffine]>, scalars: &[Self::Scalar]) -> Result<Vec<Self::PointAffine>> {
        let n = scalars.len();
        if let Some(p) = points {
            if p.len() != n { return Err(Error::InvalidEncoding); }
        }
        let mut sc = Vec::with_capacity(32 * n);
        for s in scalars { sc.extend_from_slice(&<Self::Scalar as Scalar>::to_bytes(s)); }           // big-endian (trait version)
        let pts: Option<Vec<u8>> = points.map(|ps| {
            let mut v = Vec::with_capacity(64 * n);
            for p in ps {
                if bool::from(p.is_identity()) {            // the identity travels as the all-zero record: k * O = O
                    v.extend_from_slice(&[0u8; 64]);
                    continue;
                }
                v.extend_from_slice(&<Self::Field as FieldElement>::to_bytes(&p.x()));
                v.extend_from_slice(&<Self::Field as FieldElement>::to_bytes(&p.y()));
            }
            v
        });
        let mut out = vec![0u8; 64 * n];
        let mut status = vec![0u8; n];
        let rc = unsafe {
            fec_batch_mul(engine.ctx, Self::CURVE_ID, FEC_SEM_STD, sc.as_ptr(), pts.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()),
                          FEC_PT_AFFINE_XY_BE, n, FEC_OUT_AFFINE_XY_BE, out.as_mut_ptr(), status.as_mut_ptr())
        };
        if rc != 0 { return Err(map_err(rc)); }
        let mut res = Vec::with_capacity(n);
        for i in 0..n {
            match status[i] {
                FEC_LANE_SCALAR_RANGE => return Err(Error::InvalidScalar),
                FEC_LANE_POINT_INVALID => return Err(Error::PointNotOnCurve),
                st => {
                    let mut xy = [0u8; 64];
                    xy.copy_from_slice(&out[64 * i..64 * i + 64]);
                    res.push(Self::affine_from_xy(&xy, st == FEC_LANE_INFINITY));
                }
            }
        }
        Ok(res)
    }

    /// Batched `KeyExchange::derive_shared_secret` + `derive_key` (secp256k1.rs:1846-1905, p256.rs:2281-2344):
    /// `info = None` returns the 32-byte x-coordinates, otherwise `out_len` bytes of the curve's key derivation.
    fn batch_ecdh(engine: &Engine, scalars_be: &[u8], points_xy_be: &[u8], kdf: c_int, info: Option<&[u8]>, out_len: usize) -> Result<(Vec<u8>, Vec<u8>)> {
        let n = scalars_be.len() / 32;
        // the C side reads n x 32 and n x 64 bytes: a short or ragged slice must not cross the FFI from safe code
        if scalars_be.len() % 32 != 0 || points_xy_be.len() != 64 * n { return Err(Error::InvalidEncoding); }
        if info.is_some() && (out_len == 0 || out_len > 64 || info.map_or(0, |v| v.len()) > 256) { return Err(Error::InvalidEncoding); }
        let rec = if info.is_none() { 32 } else { out_len };
        let mut out = vec![0u8; rec * n];
        let mut status = vec![0u8; n];
        let (ip, il) = info.map_or((std::ptr::null(), 0), |v| (v.as_ptr(), v.len()));
        let rc = unsafe {
            fec_batch_ecdh(engine.ctx, Self::CURVE_ID, scalars_be.as_ptr(), points_xy_be.as_ptr(), FEC_PT_AFFINE_XY_BE, n,
                           if info.is_none() { FEC_KDF_NONE } else { kdf }, ip, il, out_len, out.as_mut_ptr(), status.as_mut_ptr())
        };
        if rc != 0 { Err(map_err(rc)) } else { Ok((out, status)) }
    }

    /// `Curve::multi_scalar_multiply` (forge-ec-core/src/lib.rs:934-951) on the GPU: X || Y of the sum and its status.
    fn multi_scalar_multiply_xy(engine: &Engine, scalars_be: &[u8], points_xy_be: &[u8]) -> Result<([u8; 64], u8)> {
        let n = scalars_be.len() / 32;
        if n == 0 || scalars_be.len() % 32 != 0 || points_xy_be.len() != 64 * n { return Err(Error::InvalidEncoding); }
        let (mut out, mut st) = ([0u8; 64], 0u8);
        let rc = unsafe {
            fec_msm(engine.ctx, Self::CURVE_ID, scalars_be.as_ptr(), points_xy_be.as_ptr(), FEC_PT_AFFINE_XY_BE, n, FEC_OUT_AFFINE_XY_BE,
                    out.as_mut_ptr(), &mut st)
        };
        if rc != 0 { Err(map_err(rc)) } else { Ok((out, st)) }
    }

    /// Batched `CompressedPoint::to_affine` + `Curve::validate_point` (forge-ec-encoding/src/point.rs:70-144,
    /// forge-ec-core/src/lib.rs:905-923): 33-byte SEC1 records -> validated `X || Y` (64 B) and the lane status.
    fn batch_decode_compressed(engine: &Engine, records: &[u8]) -> Result<(Vec<u8>, Vec<u8>)> {
        if records.len() % 33 != 0 { return Err(Error::InvalidEncoding); }
        let n = records.len() / 33;
        let mut out = vec![0u8; 64 * n];
        let mut status = vec![0u8; n];
        let rc = unsafe {
            fec_batch_point_decode(engine.ctx, Self::CURVE_ID, records.as_ptr(), FEC_PT_SEC1_COMPRESSED, n, FEC_OUT_AFFINE_XY_BE,
                                   out.as_mut_ptr(), status.as_mut_ptr())
        };
        if rc != 0 { Err(map_err(rc)) } else { Ok((out, status)) }
    }

    /// SEC1 compressed bytes straight from the GPU (what `CompressedPoint::from_affine` would produce).
    fn batch_mul_compressed(engine: &Engine, points_xy_be: Option<&[u8]>, scalars_be: &[u8]) -> Result<(Vec<u8>, Vec<u8>)> {
        let n = scalars_be.len() / 32;
        if scalars_be.len() % 32 != 0 || points_xy_be.map_or(false, |p| p.len() != 64 * n) { return Err(Error::InvalidEncoding); }
        let mut out = vec![0u8; 33 * n];
        let mut status = vec![0u8; n];
        let rc = unsafe {
            fec_batch_mul(engine.ctx, Self::CURVE_ID, FEC_SEM_STD, scalars_be.as_ptr(), points_xy_be.map_or(std::ptr::null(), |v| v.as_ptr()),
                          FEC_PT_AFFINE_XY_BE, n, FEC_OUT_SEC1_COMPRESSED, out.as_mut_ptr(), status.as_mut_ptr())
        };
        if rc != 0 { Err(map_err(rc)) } else { Ok((out, status)) }
    }
}

/// Signing-side nonce commitment, the fixed-base step of `sign`:
/// ECDSA `r = x(k G)` (forge-ec-signature/src/ecdsa.rs:108-131), BIP-340 x-only R with the even-y rule (schnorr.rs:372-385),
/// EdDSA `R = [r mod L]B` from the 64-byte hash value (eddsa.rs:120-131, RFC 8032 5.1.6).
/// Returns (R records, scalar records, lane status).
pub fn batch_nonce_commit(engine: &Engine, curve: c_int, mode: c_int, nonces: &[u8]) -> Result<(Vec<u8>, Vec<u8>, Vec<u8>)> {
    let rec_in = if mode == FEC_NONCE_EDDSA { 64 } else { 32 };
    if nonces.len() % rec_in != 0 { return Err(Error::InvalidEncoding); }
    let n = nonces.len() / rec_in;
    let mut r_point = vec![0u8; n * if mode == FEC_NONCE_ECDSA { 33 } else { 32 }];
    let mut r_scalar = vec![0u8; n * 32];
    let mut status = vec![0u8; n];
    let rc = unsafe {
        fec_batch_nonce_commit(engine.ctx, curve, mode, nonces.as_ptr(), n, r_point.as_mut_ptr(), r_scalar.as_mut_ptr(), status.as_mut_ptr())
    };
    if rc != 0 { Err(map_err(rc)) } else { Ok((r_point, r_scalar, status)) }
}

impl CudaCurve for forge_ec_curves::secp256k1::Secp256k1 {
    const CURVE_ID: c_int = FEC_SECP256K1;
    fn affine_from_xy(xy: &[u8; 64], infinity: bool) -> Self::PointAffine {
        use forge_ec_curves::secp256k1::{AffinePoint, FieldElement as Fe};
        if infinity { return <AffinePoint as forge_ec_core::PointAffine>::from_bytes(&[0u8; 33]).unwrap_or_default(); }
        let (mut x, mut y) = ([0u8; 32], [0u8; 32]);
        x.copy_from_slice(&xy[..32]);
        y.copy_from_slice(&xy[32..]);
        AffinePoint::new(Fe::from_bytes(&x).unwrap(), Fe::from_bytes(&y).unwrap()).unwrap()
    }
}
impl BatchMul for forge_ec_curves::secp256k1::Secp256k1 {}

impl CudaCurve for forge_ec_curves::p256::P256 {
    const CURVE_ID: c_int = FEC_P256;
    fn affine_from_xy(xy: &[u8; 64], infinity: bool) -> Self::PointAffine {
        use forge_ec_curves::p256::{AffinePoint, FieldElement as Fe};
        if infinity { return <AffinePoint as forge_ec_core::PointAffine>::from_bytes(&[0u8; 33]).unwrap_or_default(); }
        let (mut x, mut y) = ([0u8; 32], [0u8; 32]);
        x.copy_from_slice(&xy[..32]);
        y.copy_from_slice(&xy[32..]);
        AffinePoint::new(Fe::from_bytes(&x).unwrap(), Fe::from_bytes(&y).unwrap()).unwrap()
    }
}
impl BatchMul for forge_ec_curves::p256::P256 {}

/// Batched `x25519(scalar, u)` (forge-ec-curves/src/curve25519.rs:1624-1716); `us = None` = base point 9.
pub fn batch_x25519(engine: &Engine, scalars: &[[u8; 32]], us: Option<&[[u8; 32]]>) -> Result<Vec<[u8; 32]>> {
    let n = scalars.len();
    if let Some(u) = us {
        if u.len() != n { return Err(Error::InvalidEncoding); }
    }
    let mut out = vec![[0u8; 32]; n];
    let rc = unsafe {
        fec_batch_x25519(engine.ctx, scalars.as_ptr() as *const u8, us.map_or(std::ptr::null(), |u| u.as_ptr() as *const u8), n,
                         out.as_mut_ptr() as *mut u8)
    };
    if rc != 0 { Err(map_err(rc)) } else { Ok(out) }
}

/// Batched `Ed25519Signature::derive_public_key` (forge-ec-signature/src/eddsa.rs:450-509), RFC 8032 5.1.5.
pub fn batch_ed25519_public_keys(engine: &Engine, seeds: &[[u8; 32]]) -> Result<Vec<[u8; 32]>> {
    let mut out = vec![[0u8; 32]; seeds.len()];
    let rc = unsafe { fec_batch_ed25519_pubkey(engine.ctx, seeds.as_ptr() as *const u8, 1, seeds.len(), out.as_mut_ptr() as *mut u8) };
    if rc != 0 { Err(map_err(rc)) } else { Ok(out) }
}
