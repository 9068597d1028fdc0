"""ctypes binding of include/forge_ec_cuda.h.  Plain pointers and sizes only; numpy arrays and
torch tensors are accepted as buffers but never appear in the C signatures."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

CURVE_SECP256K1, CURVE_P256 = 0, 1
PRIME_SECP256K1, PRIME_P256, PRIME_25519 = 0, 1, 2
OUT_SEC1_COMPRESSED, OUT_SEC1_UNCOMPRESSED, OUT_AFFINE_XY_BE, OUT_X_BE, OUT_SEC1_HYBRID = 0, 1, 2, 3, 4
KDF_NONE, KDF_HKDF_SHA256, KDF_XOR_INFO = 0, 1, 2
OUT_SIZE = {0: 33, 1: 65, 2: 64, 3: 32, 4: 65}
PT_AFFINE_XY_BE, PT_SEC1_COMPRESSED = 0, 1
PT_SIZE = {0: 64, 1: 33}
SEM_STD, SEM_FORGE_QUIRK = 0, 1
FIELD_MUL, FIELD_SQR, FIELD_MULSQR = 0, 1, 2
LANE_OK, LANE_SCALAR_RANGE, LANE_POINT_INVALID, LANE_INFINITY = 0, 1, 2, 3
NONCE_ECDSA, NONCE_BIP340, NONCE_EDDSA = 0, 1, 2

EXPORTED_SYMBOLS = (
    "fec_ctx_create", "fec_ctx_destroy", "fec_ctx_device_count", "fec_ctx_launch_count", "fec_strerror",
    "fec_ctx_last_cuda_error", "fec_batch_mul", "fec_batch_mul_dev", "fec_batch_x25519", "fec_batch_x25519_dev",
    "fec_batch_ed25519_pubkey", "fec_batch_ed25519_pubkey_dev", "fec_field_bench", "fec_field_bench_dev",
    "fec_imad_peak", "fec_synth_dev", "fec_shard_range", "fec_ctx_stage_timing", "fec_ctx_stage_ms",
    "fec_batch_ecdh", "fec_batch_ecdh_dev", "fec_msm", "fec_msm_dev", "fec_batch_x25519_sem", "fec_batch_x25519_sem_dev",
    "fec_batch_ed25519_pubkey_sem", "fec_batch_ed25519_pubkey_sem_dev", "fec_batch_nonce_commit", "fec_batch_nonce_commit_dev",
    "fec_batch_point_decode", "fec_batch_point_decode_dev",
)


class FecError(RuntimeError):
    """Whole-call failure of the C ABI (maps to forge_ec_core::Error, forge-ec-core/src/lib.rs:70-104)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"forge_ec_cuda error {code}: {msg}")
        self.code = code


def library_path() -> str:
    override = os.environ.get("FORGE_EC_CUDA_LIB")  # A/B experiments with differently tuned builds of the same library
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libforge_ec_cuda.so")


_LIB = None


def load_library() -> ctypes.CDLL:
    """Loads the in-tree shared library; fails loudly if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise FecError(-1, f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C forge-ec_b200/csrc).  There is no CPU fallback.")
    lib = ctypes.CDLL(path)
    vp, sz, ip, u8 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p
    lib.fec_ctx_create.argtypes = [ctypes.POINTER(ctypes.c_int), ip, ctypes.POINTER(vp)]
    lib.fec_ctx_create.restype = ip
    lib.fec_ctx_destroy.argtypes = [vp]
    lib.fec_ctx_destroy.restype = None
    lib.fec_ctx_device_count.argtypes = [vp]
    lib.fec_ctx_device_count.restype = ip
    lib.fec_ctx_launch_count.argtypes = [vp]
    lib.fec_ctx_launch_count.restype = ctypes.c_uint64
    lib.fec_strerror.argtypes = [ip]
    lib.fec_strerror.restype = ctypes.c_char_p
    lib.fec_ctx_last_cuda_error.argtypes = [vp]
    lib.fec_ctx_last_cuda_error.restype = ctypes.c_char_p
    lib.fec_batch_mul.argtypes = [vp, ip, ip, u8, u8, ip, sz, ip, u8, u8]
    lib.fec_batch_mul.restype = ip
    lib.fec_batch_mul_dev.argtypes = [vp, ip, ip, ip, u8, u8, ip, sz, ip, u8, u8, vp]
    lib.fec_batch_mul_dev.restype = ip
    lib.fec_batch_x25519.argtypes = [vp, u8, u8, sz, u8]
    lib.fec_batch_x25519.restype = ip
    lib.fec_batch_x25519_dev.argtypes = [vp, ip, u8, u8, sz, u8, vp]
    lib.fec_batch_x25519_dev.restype = ip
    lib.fec_batch_x25519_sem.argtypes = [vp, ip, u8, u8, sz, u8]
    lib.fec_batch_x25519_sem.restype = ip
    lib.fec_batch_x25519_sem_dev.argtypes = [vp, ip, ip, u8, u8, sz, u8, vp]
    lib.fec_batch_x25519_sem_dev.restype = ip
    lib.fec_batch_ed25519_pubkey_sem.argtypes = [vp, ip, u8, ip, sz, u8]
    lib.fec_batch_ed25519_pubkey_sem.restype = ip
    lib.fec_batch_ed25519_pubkey_sem_dev.argtypes = [vp, ip, ip, u8, ip, sz, u8, vp]
    lib.fec_batch_ed25519_pubkey_sem_dev.restype = ip
    lib.fec_batch_ed25519_pubkey.argtypes = [vp, u8, ip, sz, u8]
    lib.fec_batch_ed25519_pubkey.restype = ip
    lib.fec_batch_ed25519_pubkey_dev.argtypes = [vp, ip, u8, ip, sz, u8, vp]
    lib.fec_batch_ed25519_pubkey_dev.restype = ip
    lib.fec_field_bench.argtypes = [vp, ip, ip, sz, ip, u8, u8, u8, ctypes.POINTER(ctypes.c_double)]
    lib.fec_field_bench.restype = ip
    lib.fec_field_bench_dev.argtypes = [vp, ip, ip, ip, sz, ip, u8, u8, u8, vp]
    lib.fec_field_bench_dev.restype = ip
    lib.fec_imad_peak.argtypes = [vp, ip, ip, ctypes.POINTER(ctypes.c_double)]
    lib.fec_imad_peak.restype = ip
    lib.fec_synth_dev.argtypes = [vp, ip, ip, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, ip, sz, u8, vp]
    lib.fec_synth_dev.restype = ip
    lib.fec_shard_range.argtypes = [sz, ip, ip, ctypes.POINTER(sz), ctypes.POINTER(sz)]
    lib.fec_shard_range.restype = None
    lib.fec_batch_ecdh.argtypes = [vp, ip, u8, u8, ip, sz, ip, ctypes.c_char_p, sz, sz, u8, u8]
    lib.fec_batch_ecdh.restype = ip
    lib.fec_batch_ecdh_dev.argtypes = [vp, ip, ip, u8, u8, ip, sz, ip, ctypes.c_char_p, sz, sz, u8, u8, vp]
    lib.fec_batch_ecdh_dev.restype = ip
    lib.fec_msm.argtypes = [vp, ip, u8, u8, ip, sz, ip, u8, u8]
    lib.fec_msm.restype = ip
    lib.fec_msm_dev.argtypes = [vp, ip, ip, u8, u8, ip, sz, ip, u8, u8, vp]
    lib.fec_msm_dev.restype = ip
    lib.fec_batch_point_decode.argtypes = [vp, ip, u8, ip, sz, ip, u8, u8]
    lib.fec_batch_point_decode.restype = ip
    lib.fec_batch_point_decode_dev.argtypes = [vp, ip, ip, u8, ip, sz, ip, u8, u8, vp]
    lib.fec_batch_point_decode_dev.restype = ip
    lib.fec_batch_nonce_commit.argtypes = [vp, ip, ip, u8, sz, u8, u8, u8]
    lib.fec_batch_nonce_commit.restype = ip
    lib.fec_batch_nonce_commit_dev.argtypes = [vp, ip, ip, ip, u8, sz, u8, u8, u8, vp]
    lib.fec_batch_nonce_commit_dev.restype = ip
    lib.fec_ctx_stage_timing.argtypes = [vp, ip]
    lib.fec_ctx_stage_timing.restype = ip
    lib.fec_ctx_stage_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]
    lib.fec_ctx_stage_ms.restype = ip
    _LIB = lib
    return lib


def shard_range(n: int, n_shards: int, index: int):
    """(first, count) of the contiguous lane range that shard `index` of `n_shards` owns (no device needed)."""
    lib = load_library()
    first, count = ctypes.c_size_t(0), ctypes.c_size_t(0)
    lib.fec_shard_range(n, n_shards, index, ctypes.byref(first), ctypes.byref(count))
    return first.value, count.value


def _host_in(buf, n: int, rec: int, name: str) -> np.ndarray:
    """bytes / bytearray / numpy uint8 -> C-contiguous uint8 array of n*rec bytes (no copy when possible)."""
    if isinstance(buf, np.ndarray):
        a = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    else:
        a = np.frombuffer(buf, dtype=np.uint8)
    if a.size != n * rec:
        raise ValueError(f"{name}: expected {n} x {rec} = {n * rec} bytes, got {a.size}")
    return a


def _ptr(a) -> Optional[int]:
    """address of a numpy array, a torch tensor, or a raw integer pointer"""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


class Engine:
    """Owns one fec_ctx.  ``devices=None`` uses the current CUDA device."""

    def __init__(self, devices: Optional[list] = None):
        self._lib = load_library()
        self._ctx = ctypes.c_void_p()
        if devices:
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = self._lib.fec_ctx_create(arr, len(devices), ctypes.byref(self._ctx))
        else:
            rc = self._lib.fec_ctx_create(None, 0, ctypes.byref(self._ctx))
        if rc != 0:
            self._ctx = ctypes.c_void_p()
            raise FecError(rc, self._lib.fec_strerror(rc).decode())

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.fec_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            extra = self._lib.fec_ctx_last_cuda_error(self._ctx).decode()
            raise FecError(rc, self._lib.fec_strerror(rc).decode() + (f" [{extra}]" if extra else ""))

    @property
    def device_count(self) -> int:
        return self._lib.fec_ctx_device_count(self._ctx)

    @property
    def launch_count(self) -> int:
        return int(self._lib.fec_ctx_launch_count(self._ctx))

    def stage_timing(self, enable: bool):
        """record CUDA events between the three secp256k1 k*P kernels of the following batch_mul_dev calls"""
        self._check(self._lib.fec_ctx_stage_timing(self._ctx, int(bool(enable))))

    def stage_ms(self):
        """(ms_prepare, ms_ladder, ms_finish, chunks) summed over the calls made since stage_timing(True)"""
        ms = (ctypes.c_double * 3)()
        chunks = ctypes.c_int(0)
        self._check(self._lib.fec_ctx_stage_ms(self._ctx, ms, ctypes.byref(chunks)))
        return ms[0], ms[1], ms[2], chunks.value

    # -- host-buffer entry points -----------------------------------------------------
    def batch_mul(self, curve: int, scalars, points=None, *, n: Optional[int] = None, out_fmt: int = OUT_SEC1_COMPRESSED,
                  points_fmt: int = PT_AFFINE_XY_BE, out: Optional[np.ndarray] = None, want_status: bool = True,
                  sem: int = SEM_STD):
        """k_i*G (points is None) or k_i*P_i.  Returns (out[n, rec] uint8, status[n] uint8 or None)."""
        if n is None:
            n = (scalars.nbytes if isinstance(scalars, np.ndarray) else len(scalars)) // 32
        s = _host_in(scalars, n, 32, "scalars")
        p = None if points is None else _host_in(points, n, PT_SIZE[points_fmt], "points")
        rec = OUT_SIZE[out_fmt]
        if out is None:
            out = np.empty((n, rec), dtype=np.uint8)
        st = np.empty(n, dtype=np.uint8) if want_status else None
        self._check(self._lib.fec_batch_mul(self._ctx, curve, sem, _ptr(s), _ptr(p), points_fmt, n, out_fmt, _ptr(out), _ptr(st)))
        return out, st

    def batch_mul_ptr(self, curve: int, scalars_ptr: int, points_ptr: Optional[int], n: int, out_ptr: int,
                      status_ptr: Optional[int] = None, *, out_fmt: int = OUT_SEC1_COMPRESSED,
                      points_fmt: int = PT_AFFINE_XY_BE):
        """Same call on raw HOST addresses (e.g. pinned torch tensors): nothing is allocated or copied here."""
        self._check(self._lib.fec_batch_mul(self._ctx, curve, SEM_STD, scalars_ptr, points_ptr, points_fmt, n, out_fmt, out_ptr,
                                            status_ptr))

    def batch_ecdh(self, curve: int, scalars, points, *, n: Optional[int] = None, kdf: int = KDF_NONE, info: bytes = b"",
                   out_len: int = 32, points_fmt: int = PT_AFFINE_XY_BE):
        """x(k_i*P_i) followed by the reference's key derivation.  Returns (out[n, 32 | out_len] uint8, status[n])."""
        if n is None:
            n = (scalars.nbytes if isinstance(scalars, np.ndarray) else len(scalars)) // 32
        s = _host_in(scalars, n, 32, "scalars")
        p = _host_in(points, n, PT_SIZE[points_fmt], "points")
        rec = 32 if kdf == KDF_NONE else out_len
        out = np.empty((n, rec), dtype=np.uint8)
        st = np.empty(n, dtype=np.uint8)
        self._check(self._lib.fec_batch_ecdh(self._ctx, curve, _ptr(s), _ptr(p), points_fmt, n, kdf, bytes(info), len(info), out_len,
                                             _ptr(out), _ptr(st)))
        return out, st

    def batch_ecdh_dev(self, curve: int, d_scalars, d_points, n: int, d_out, d_status=None, *, kdf: int = KDF_NONE,
                       info: bytes = b"", out_len: int = 32, points_fmt: int = 0, stream: int = 0, dev_index: int = 0):
        self._check(self._lib.fec_batch_ecdh_dev(self._ctx, dev_index, curve, _ptr(d_scalars), _ptr(d_points), points_fmt, n, kdf,
                                                 bytes(info), len(info), out_len, _ptr(d_out), _ptr(d_status), stream))

    def msm(self, curve: int, scalars, points, *, n: Optional[int] = None, out_fmt: int = OUT_SEC1_COMPRESSED,
            points_fmt: int = PT_AFFINE_XY_BE):
        """sum_i k_i*P_i.  Returns (record bytes, status int): status 0 ok, 1/2 a lane was rejected, 3 identity."""
        if n is None:
            n = (scalars.nbytes if isinstance(scalars, np.ndarray) else len(scalars)) // 32
        s = _host_in(scalars, n, 32, "scalars")
        p = _host_in(points, n, PT_SIZE[points_fmt], "points")
        out = np.empty(OUT_SIZE[out_fmt], dtype=np.uint8)
        st = np.zeros(1, dtype=np.uint8)
        self._check(self._lib.fec_msm(self._ctx, curve, _ptr(s), _ptr(p), points_fmt, n, out_fmt, _ptr(out), _ptr(st)))
        return out.tobytes(), int(st[0])

    def msm_dev(self, curve: int, d_scalars, d_points, n: int, d_out, d_status=None, *, out_fmt: int = 0, points_fmt: int = 0,
                stream: int = 0, dev_index: int = 0):
        self._check(self._lib.fec_msm_dev(self._ctx, dev_index, curve, _ptr(d_scalars), _ptr(d_points), points_fmt, n, out_fmt,
                                          _ptr(d_out), _ptr(d_status), stream))

    def batch_point_decode(self, curve: int, points, *, n: Optional[int] = None, points_fmt: int = PT_SEC1_COMPRESSED,
                           out_fmt: int = OUT_AFFINE_XY_BE):
        """SEC1 decoding / curve-membership validation of n points, re-encoded in out_fmt.  Returns (out[n, rec], status[n])."""
        if n is None:
            n = (points.nbytes if isinstance(points, np.ndarray) else len(points)) // PT_SIZE[points_fmt]
        p = _host_in(points, n, PT_SIZE[points_fmt], "points")
        out = np.empty((n, OUT_SIZE[out_fmt]), dtype=np.uint8)
        st = np.empty(n, dtype=np.uint8)
        self._check(self._lib.fec_batch_point_decode(self._ctx, curve, _ptr(p), points_fmt, n, out_fmt, _ptr(out), _ptr(st)))
        return out, st

    def batch_point_decode_dev(self, curve: int, d_points, n: int, d_out, d_status=None, *, points_fmt: int = PT_SEC1_COMPRESSED,
                               out_fmt: int = OUT_AFFINE_XY_BE, stream: int = 0, dev_index: int = 0):
        self._check(self._lib.fec_batch_point_decode_dev(self._ctx, dev_index, curve, _ptr(d_points), points_fmt, n, out_fmt, _ptr(d_out),
                                                         _ptr(d_status), stream))

    def batch_nonce_commit(self, curve: int, mode: int, nonces, *, n: Optional[int] = None):
        """Signing-side nonce commitment (ECDSA r = x(k G), BIP-340 even-y rule, EdDSA [r mod L]B).
        Returns (out_R[n, 33 | 32], out_r[n, 32], status[n])."""
        rec_in = 64 if mode == NONCE_EDDSA else 32
        if n is None:
            n = (nonces.nbytes if isinstance(nonces, np.ndarray) else len(nonces)) // rec_in
        a = _host_in(nonces, n, rec_in, "nonces")
        R = np.empty((n, 33 if mode == NONCE_ECDSA else 32), dtype=np.uint8)
        r = np.empty((n, 32), dtype=np.uint8)
        st = np.empty(n, dtype=np.uint8)
        self._check(self._lib.fec_batch_nonce_commit(self._ctx, curve, mode, _ptr(a), n, _ptr(R), _ptr(r), _ptr(st)))
        return R, r, st

    def batch_nonce_commit_dev(self, curve: int, mode: int, d_nonces, n: int, d_out_R, d_out_r, d_status=None, *, stream: int = 0,
                               dev_index: int = 0):
        self._check(self._lib.fec_batch_nonce_commit_dev(self._ctx, dev_index, curve, mode, _ptr(d_nonces), n, _ptr(d_out_R),
                                                         _ptr(d_out_r), _ptr(d_status), stream))

    def batch_x25519(self, scalars, us=None, *, n: Optional[int] = None, sem: int = SEM_STD) -> np.ndarray:
        if n is None:
            n = (scalars.nbytes if isinstance(scalars, np.ndarray) else len(scalars)) // 32
        s = _host_in(scalars, n, 32, "scalars")
        u = None if us is None else _host_in(us, n, 32, "u")
        out = np.empty((n, 32), dtype=np.uint8)
        if sem == SEM_STD:
            self._check(self._lib.fec_batch_x25519(self._ctx, _ptr(s), _ptr(u), n, _ptr(out)))
        else:
            self._check(self._lib.fec_batch_x25519_sem(self._ctx, sem, _ptr(s), _ptr(u), n, _ptr(out)))
        return out

    def batch_ed25519_pubkey(self, data, input_is_seed: bool = True, *, n: Optional[int] = None, sem: int = SEM_STD) -> np.ndarray:
        if n is None:
            n = (data.nbytes if isinstance(data, np.ndarray) else len(data)) // 32
        d = _host_in(data, n, 32, "seeds")
        out = np.empty((n, 32), dtype=np.uint8)
        if sem == SEM_STD:
            self._check(self._lib.fec_batch_ed25519_pubkey(self._ctx, _ptr(d), int(bool(input_is_seed)), n, _ptr(out)))
        else:
            self._check(self._lib.fec_batch_ed25519_pubkey_sem(self._ctx, sem, _ptr(d), int(bool(input_is_seed)), n, _ptr(out)))
        return out

    def field_bench(self, prime: int, op: int, a, b, iters: int, *, lanes: Optional[int] = None):
        """Returns (out[lanes,32] little-endian canonical residues, kernel seconds)."""
        if lanes is None:
            lanes = (a.nbytes if isinstance(a, np.ndarray) else len(a)) // 32
        aa = _host_in(a, lanes, 32, "a")
        bb = _host_in(b, lanes, 32, "b")
        out = np.empty((lanes, 32), dtype=np.uint8)
        sec = ctypes.c_double(0.0)
        self._check(self._lib.fec_field_bench(self._ctx, prime, op, lanes, iters, _ptr(aa), _ptr(bb), _ptr(out), ctypes.byref(sec)))
        return out, sec.value

    def imad_peak(self, variant: int, dev_index: int = 0) -> float:
        """multiply-add instructions per second of the calibration kernel (see forge_ec_cuda.h)."""
        r = ctypes.c_double(0.0)
        self._check(self._lib.fec_imad_peak(self._ctx, dev_index, variant, ctypes.byref(r)))
        return r.value

    # -- device-pointer entry points (torch tensors or raw addresses; asynchronous on `stream`) ----
    def batch_mul_dev(self, curve: int, d_scalars, d_points, n: int, d_out, d_status=None, *, out_fmt: int = 0,
                      points_fmt: int = 0, stream: int = 0, dev_index: int = 0, sem: int = SEM_STD):
        self._check(self._lib.fec_batch_mul_dev(self._ctx, dev_index, curve, sem, _ptr(d_scalars), _ptr(d_points), points_fmt,
                                                n, out_fmt, _ptr(d_out), _ptr(d_status), stream))

    def batch_x25519_dev(self, d_scalars, d_us, n: int, d_out, *, stream: int = 0, dev_index: int = 0, sem: int = SEM_STD):
        self._check(self._lib.fec_batch_x25519_sem_dev(self._ctx, dev_index, sem, _ptr(d_scalars), _ptr(d_us), n, _ptr(d_out), stream))

    def batch_ed25519_pubkey_dev(self, d_in, input_is_seed: bool, n: int, d_out, *, stream: int = 0, dev_index: int = 0, sem: int = SEM_STD):
        self._check(self._lib.fec_batch_ed25519_pubkey_sem_dev(self._ctx, dev_index, sem, _ptr(d_in), int(bool(input_is_seed)), n,
                                                               _ptr(d_out), stream))

    def field_bench_dev(self, prime: int, op: int, lanes: int, iters: int, d_a, d_b, d_out, *, stream: int = 0, dev_index: int = 0):
        self._check(self._lib.fec_field_bench_dev(self._ctx, dev_index, prime, op, lanes, iters, _ptr(d_a), _ptr(d_b), _ptr(d_out),
                                                  stream))

    def synth_dev(self, mode: int, stream_id: int, sub: int, first_lane: int, n: int, d_out, *, edge: bool = False,
                  stream: int = 0, dev_index: int = 0):
        """Philox inputs on the device: mode 0 raw bytes, 1/2 scalars below the secp256k1/P-256 order."""
        self._check(self._lib.fec_synth_dev(self._ctx, dev_index, mode, stream_id, sub, first_lane, int(edge), n, _ptr(d_out), stream))
