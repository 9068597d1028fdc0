import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no built library (build products are not tracked): build it once with nvcc, exactly as
    __graft_entry__.build() does.  Building is not a fallback -- without nvcc the tests that need the library fail loudly."""
    import shutil
    import subprocess

    lib = os.path.join(ROOT, "forge-ec_b200", "lib", "libforge_ec_cuda.so")
    if not os.path.exists(lib) and shutil.which("nvcc") and shutil.which("make"):
        subprocess.run(["make", "-j", str(min(8, os.cpu_count() or 1)), "-C", os.path.join(ROOT, "forge-ec_b200", "csrc")], check=False,
                       stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def fec():
    """the host package (hyphenated directory name -> importlib)"""
    return importlib.import_module("forge-ec_b200")


@pytest.fixture(scope="session")
def pm():
    """Python big-integer model of the path (oracle/pymodel.py) -- the checker, never the product"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pymodel

    return pymodel


@pytest.fixture(scope="session")
def engine(fec):
    """one context on cuda:0; a missing library or device is an error, not a skip"""
    eng = fec.Engine()
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def ostd():
    """oracle/std/oracle_std.c (threaded C restatement of the specified behaviour) -- the checker, never the product"""
    import ctypes
    import subprocess

    import numpy as np

    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    ncpu = os.cpu_count() or 4
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_std.so"))

    class O:
        @staticmethod
        def mul(curve, sc, pts=None, pfmt=0, ofmt=0, threads=ncpu):
            s = np.frombuffer(bytes(sc), dtype=np.uint8)
            n = s.size // 32
            p = None if pts is None else np.frombuffer(bytes(pts), dtype=np.uint8)
            out = np.empty(n * {0: 33, 1: 65, 2: 64, 3: 32, 4: 65}[ofmt], dtype=np.uint8)
            st = np.empty(n, dtype=np.uint8)
            lib.ostd_batch_mul(curve, vp(s.ctypes.data), vp(p.ctypes.data) if p is not None else None, pfmt, sz(n), ofmt,
                               vp(out.ctypes.data), vp(st.ctypes.data), threads)
            return out.tobytes(), st.tobytes()

        @staticmethod
        def msm(curve, sc, pts, pfmt=0, ofmt=0, threads=ncpu):
            """Curve::multi_scalar_multiply: (one record, status)"""
            s = np.frombuffer(bytes(sc), dtype=np.uint8)
            n = s.size // 32
            p = np.frombuffer(bytes(pts), dtype=np.uint8)
            out = np.empty({0: 33, 1: 65, 2: 64, 3: 32, 4: 65}[ofmt], dtype=np.uint8)
            st = np.empty(1, dtype=np.uint8)
            lib.ostd_msm(curve, vp(s.ctypes.data), vp(p.ctypes.data), pfmt, sz(n), ofmt, vp(out.ctypes.data), vp(st.ctypes.data), threads)
            return out.tobytes(), int(st[0])

        @staticmethod
        def msm_products(curve, xy, st, ofmt=0):
            """the same from products that `mul(..., ofmt=2)` has already computed"""
            a = np.frombuffer(bytes(xy), dtype=np.uint8)
            b = np.frombuffer(bytes(st), dtype=np.uint8)
            out = np.empty({0: 33, 1: 65, 2: 64, 3: 32, 4: 65}[ofmt], dtype=np.uint8)
            s1 = np.empty(1, dtype=np.uint8)
            lib.ostd_msm_products(curve, vp(a.ctypes.data), vp(b.ctypes.data), sz(b.size), ofmt, vp(out.ctypes.data), vp(s1.ctypes.data))
            return out.tobytes(), int(s1[0])

        @staticmethod
        def point_decode(curve, pts, pfmt=1, ofmt=2, threads=ncpu):
            p = np.frombuffer(bytes(pts), dtype=np.uint8)
            n = p.size // (33 if pfmt == 1 else 64)
            out = np.empty(n * {0: 33, 1: 65, 2: 64, 3: 32, 4: 65}[ofmt], dtype=np.uint8)
            st = np.empty(n, dtype=np.uint8)
            lib.ostd_batch_point_decode(curve, vp(p.ctypes.data), pfmt, sz(n), ofmt, vp(out.ctypes.data), vp(st.ctypes.data), threads)
            return out.tobytes(), st.tobytes()

        @staticmethod
        def nonce_commit(curve, mode, nonces, threads=ncpu):
            """(out_R, out_r, status) of the signing-side nonce commitment"""
            a = np.frombuffer(bytes(nonces), dtype=np.uint8)
            n = a.size // (64 if mode == 2 else 32)
            R = np.empty(n * (33 if mode == 0 else 32), dtype=np.uint8)
            r = np.empty(n * 32, dtype=np.uint8)
            st = np.empty(n, dtype=np.uint8)
            lib.ostd_batch_nonce_commit(curve, mode, vp(a.ctypes.data), sz(n), vp(R.ctypes.data), vp(r.ctypes.data), vp(st.ctypes.data), threads)
            return R.tobytes(), r.tobytes(), st.tobytes()

        @staticmethod
        def x25519(ks, us=None, threads=ncpu):
            k = np.frombuffer(bytes(ks), dtype=np.uint8)
            n = k.size // 32
            u = None if us is None else np.frombuffer(bytes(us), dtype=np.uint8)
            out = np.empty(n * 32, dtype=np.uint8)
            lib.ostd_batch_x25519(vp(k.ctypes.data), vp(u.ctypes.data) if u is not None else None, sz(n), vp(out.ctypes.data), threads)
            return out.tobytes()

        @staticmethod
        def ed(data, is_seed=True, threads=ncpu):
            d = np.frombuffer(bytes(data), dtype=np.uint8)
            n = d.size // 32
            out = np.empty(n * 32, dtype=np.uint8)
            lib.ostd_batch_ed25519_pubkey(vp(d.ctypes.data), int(is_seed), sz(n), vp(out.ctypes.data), threads)
            return out.tobytes()

        @staticmethod
        def field(prime, op, iters, a, b, threads=2):
            aa = np.frombuffer(bytes(a), dtype=np.uint8)
            bb = np.frombuffer(bytes(b), dtype=np.uint8)
            n = aa.size // 32
            out = np.empty(n * 32, dtype=np.uint8)
            lib.ostd_field_bench(prime, op, sz(n), iters, vp(aa.ctypes.data), vp(bb.ctypes.data), vp(out.ctypes.data), threads)
            return out.tobytes()

        @staticmethod
        def sha512(m):
            d = ctypes.create_string_buffer(64)
            lib.ostd_sha512(m, sz(len(m)), d)
            return d.raw

    return O
