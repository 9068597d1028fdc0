"""CPU tests of the host side: the C-ABI library loads and exports what include/forge_ec_cuda.h declares, fails
loudly without a device, the device headers' limb arithmetic (compiled for the host by tests/hostsim) is
bit-exact against the oracle, and the multi-rank sharding logic works under gloo with world_size 2."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(fec):
    hdr = open(os.path.join(ROOT, "include", "forge_ec_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)  # declarations only
    declared = sorted(set(re.findall(r"\b(fec_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 17
    lib = fec.load_library()
    for sym in declared:
        assert hasattr(lib, sym), f"libforge_ec_cuda.so does not export {sym}"
    assert sorted(fec.EXPORTED_SYMBOLS) == declared
    # ... and nothing else: the kernels' launchers (namespace fec) stay internal (hidden default visibility)
    nm = subprocess.run(["nm", "-D", "--defined-only", fec.library_path()], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in nm.splitlines() if len(l.split()) == 3 and l.split()[1] in "TW")
    assert [s for s in exported if s.startswith("fec_")] == declared
    assert not [s for s in exported if "3fec" in s], "internal launchers are exported"


def test_strerror_and_argument_checks_need_no_device(fec):
    lib = fec.load_library()
    assert lib.fec_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.fec_strerror(3)
    assert lib.fec_ctx_create(None, 0, None) == 1             # FEC_ERR_INVALID_ARG
    assert lib.fec_batch_mul(None, 0, 0, None, None, 0, 0, 0, None, None) == 1
    assert lib.fec_ctx_device_count(None) == 0 and lib.fec_ctx_launch_count(None) == 0


def test_no_cpu_fallback(fec):
    """without a CUDA device the engine must refuse to exist (with one, it must come up)"""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        fec.Engine().close()
    else:
        with pytest.raises(fec.FecError) as e:
            fec.Engine()
        assert e.value.code == 3


def test_shard_ranges(fec):
    for n in (0, 1, 1023, 1024, 1025, 10_000, 1 << 24, (1 << 24) + 77):
        for g in (1, 2, 3, 4, 8):
            pos = 0
            for i in range(g):
                first, count = fec.shard_range(n, g, i)
                assert first == pos and (first % 1024 == 0 or first == n)
                pos += count
            assert pos == n
    assert fec.shard_range(1 << 24, 8, 3) == (3 << 21, 1 << 21)


def test_trait_mirror_encoders(fec):
    xy = bytes(range(1, 65))
    p = fec.AffinePoint(xy)
    assert p.to_bytes() == bytes([2 | (xy[63] & 1)]) + xy[:32] and p.to_bytes_with_format("uncompressed") == b"\x04" + xy
    inf = fec.AffinePoint(bytes(64), True)
    assert inf.to_bytes() == bytes(33) and inf.to_bytes_with_format("uncompressed") == bytes(65) and inf.is_identity()
    with pytest.raises(ValueError):
        fec.Scalar.from_bytes(b"\x00" * 31)
    assert fec.Scalar.from_int(7).to_bytes() == (7).to_bytes(32, "big")


@pytest.fixture(scope="session")
def hostsim():
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    out_dir = os.path.join(ROOT, "tests", "hostsim", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostsim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-x", "c++", "-shared", "-fPIC", "-o", so, src], check=True)
    return ctypes.CDLL(so)


def test_device_field_code_on_host_matches_model(hostsim, pm):
    import random
    rnd = random.Random(7)
    for prime in (0, 1, 2):
        p = pm.PRIMES[prime]
        vals = [0, 1, 2, p - 1, p - 2, (p + 1) // 2, (1 << 255) % p, (1 << 224) % p, (2**256 - 1) % p] + [rnd.randrange(p) for _ in range(60)]
        for _ in range(600):
            a, b = rnd.choice(vals), rnd.choice(vals)
            exp = [a * b % p, a * a % p, (a + b) % p, (a - b) % p, (-a) % p]
            for op in range(5):
                o = ctypes.create_string_buffer(32)
                hostsim.hs_field_op(prime, op, a.to_bytes(32, "little"), b.to_bytes(32, "little"), o)
                assert int.from_bytes(o.raw, "little") == exp[op], (prime, op, hex(a), hex(b))
        for a in vals[1:20]:
            o = ctypes.create_string_buffer(32)
            hostsim.hs_field_op(prime, 5, a.to_bytes(32, "little"), bytes(32), o)
            assert int.from_bytes(o.raw, "little") == pow(a, -1, p)
        # single-pass small multiples and the zero test, also on non-canonical (weak) representatives
        weak = vals + ([p, p + 1, 2**256 - 1, 2**256 - 2**32 - 978, 2**256 - 2**33] if prime != 1 else []) + ([2 * p] if prime == 2 else [])
        for a in weak:
            for op, m in ((7, 2), (8, 3), (9, 4), (10, 8)):
                o = ctypes.create_string_buffer(32)
                hostsim.hs_field_op(prime, op, a.to_bytes(32, "little"), bytes(32), o)
                assert int.from_bytes(o.raw, "little") == a * m % p, (prime, op, hex(a))
            o = ctypes.create_string_buffer(32)
            hostsim.hs_field_op(prime, 11, a.to_bytes(32, "little"), bytes(32), o)
            assert int.from_bytes(o.raw, "little") == (1 if a % p == 0 else 0), (prime, hex(a))


def test_fused_linear_chains_of_the_secp256k1_field(hostsim, pm):
    """2 (t - a - b), t - 2a, t - a - 2b, t - 8a, +-t - a (and 16 t) with ONE fold of the wrapped value (fields.cuh: fold_signed) against the integers,
    on weak representatives (anything below 2^256) and on the operands that wrap the most often"""
    import random
    rnd = random.Random(23)
    p = pm.PRIMES[0]
    edge = [0, 1, 2, 976, 977, 978, 2**32 + 976, 2**32 + 977, 2**32 + 978, 2**33, p - 1, p, p + 1, 2**256 - 1, 2**256 - 2, 2**256 - 978,
            2**256 - 2**32 - 978, 2**256 - 2**33, 2**255, 2**255 - 1, 2**253, 3 * 2**253, 7 * 2**253, (1 << 224) - 1]
    vals = edge + [rnd.randrange(2**256) for _ in range(40)]
    le = lambda v: v.to_bytes(32, "little")
    exp = [lambda t, a, b: 2 * (t - a - b), lambda t, a, b: t - 2 * a, lambda t, a, b: t - a - 2 * b, lambda t, a, b: t - 8 * a,
           lambda t, a, b: t - a, lambda t, a, b: -t - a, lambda t, a, b: 16 * t,
           lambda t, a, b: (t - 2 * (2 * (t - a - b))) + (2 * (t - a - b) - a)]
    o = ctypes.create_string_buffer(32)
    for _ in range(6000):
        t, a, b = rnd.choice(vals), rnd.choice(vals), rnd.choice(vals)
        for op in range(8):
            hostsim.hs_k1_fused(op, le(t), le(a), le(b), o)
            assert int.from_bytes(o.raw, "little") == exp[op](t, a, b) % p, (op, hex(t), hex(a), hex(b))
    for t in edge:
        for a in edge:
            for b in edge[:12]:
                for op in range(8):
                    hostsim.hs_k1_fused(op, le(t), le(a), le(b), o)
                    assert int.from_bytes(o.raw, "little") == exp[op](t, a, b) % p, (op, hex(t), hex(a), hex(b))


def test_device_lane_code_on_host_matches_model(hostsim, pm):
    """recoding, table building, exceptional additions, encoders, ladder, SHA-512, Ed25519 comb"""
    import random
    rnd = random.Random(11)
    be = lambda v: v.to_bytes(32, "big")
    for curve in (0, 1):
        c = pm.CURVES[curve]
        ks = [0, 1, 2, 3, c.n - 1, c.n - 2, (c.n + 1) // 2, c.n, 2**256 - 1, (30 << 252) - c.n, 15 << 252] + [rnd.randrange(c.n) for _ in range(6)]
        for k in ks:
            o = ctypes.create_string_buffer(33)
            for use_vm in (1, 0):  # accumulator-machine kernel and register-resident kernel
                hostsim.hs_set_var_impl(use_vm)
                st = hostsim.hs_wmul(curve, be(k), None, 0, 1, 0, o)
                assert (o.raw, bytes([st])) == pm.batch_mul(curve, be(k), None, 1, 0), ("fixed", curve, hex(k), use_vm)
            P = pm.w_mul(c, rnd.randrange(1, c.n), (c.gx, c.gy))
            for pt, pfmt in ((be(P[0]) + be(P[1]), 0), (pm.sec1_encode(P, 0), 1), (be(c.gx) + be(c.gy), 0)):
                for use_vm in (1, 0):  # three-stage accumulator-machine path and single-kernel path
                    hostsim.hs_set_var_impl(use_vm)
                    st = hostsim.hs_wmul(curve, be(k), pt, pfmt, 0, 0, o)
                    assert (o.raw, bytes([st])) == pm.batch_mul(curve, be(k), pt, 1, 0, pfmt), ("var", curve, hex(k), pfmt, use_vm)
        o = ctypes.create_string_buffer(33)
        assert hostsim.hs_wmul(curve, be(0), be(2) + be(2), 0, 0, 0, o) == 2 and o.raw == bytes(33)
    for i in range(8):
        k = bytes(rnd.randrange(256) for _ in range(32))
        u = bytes(rnd.randrange(256) for _ in range(32))
        o = ctypes.create_string_buffer(32)
        hostsim.hs_x25519(k, u, o)
        assert o.raw == pm.x25519(k, u)
        hostsim.hs_x25519_vm(k, u, o)
        assert o.raw == pm.x25519(k, u)
    # edge inputs of the ladder (both formulations of the conditional swap): u = 0, 1, p - 1, p, p + 1, 2^255 - 1, 2^256 - 1
    # (non-canonical and bit 255 set: masked / reduced as RFC 7748 says), a point of order 8, scalars of all zeros / all ones
    p25519 = 2**255 - 19
    us = [0, 1, p25519 - 1, p25519, p25519 + 1, 2**255 - 1, 2**256 - 1, 9,
          0x00b8495f16056286fdb1329ceb8d09da6ac49ff1fae35616aeb8413b7c7aebe0]
    for uv in us:
        for k in (bytes(32), b"\xff" * 32, bytes(range(32)), bytes([rnd.randrange(256) for _ in range(32)])):
            u = uv.to_bytes(32, "little")
            o = ctypes.create_string_buffer(32)
            hostsim.hs_x25519(k, u, o)
            assert o.raw == pm.x25519(k, u), (hex(uv), k.hex())
            hostsim.hs_x25519_vm(k, u, o)
            assert o.raw == pm.x25519(k, u), (hex(uv), k.hex())
        hostsim.hs_ed25519_pubkey(k, 1, o)
        assert o.raw == pm.ed25519_pubkey(k)
        hostsim.hs_ed25519_pubkey(k, 0, o)
        assert o.raw == pm.ed25519_pubkey(k, False)
        hostsim.hs_ed25519_pubkey_ct(k, 1, o)    # the 4-bit scanned comb of FORGE_EC_CT_TABLE=1
        assert o.raw == pm.ed25519_pubkey(k)
        hostsim.hs_ed25519_pubkey_ct(k, 0, o)
        assert o.raw == pm.ed25519_pubkey(k, False)


def test_dual_operation_programs_of_the_accumulator_machine(pm):
    """FEC_VM_DUAL (vm.cuh): the secp256k1 doubling / mixed addition written with two independent field operations per
    routine call (A/B form) must give the same k*P as the model, edge scalars included."""
    import random
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    out_dir = os.path.join(ROOT, "tests", "hostsim", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostsim_dual.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-DFEC_VM_DUAL", "-x", "c++", "-shared", "-fPIC", "-o", so, src], check=True)
    hs = ctypes.CDLL(so)
    rnd = random.Random(23)
    be = lambda v: v.to_bytes(32, "big")
    hs.hs_set_var_impl(1)
    o = ctypes.create_string_buffer(33)
    for curve in (0, 1):   # P-256: the interleaved two-multiplication routine FeP256::mul2_inl as well
        c = pm.CURVES[curve]
        ks = [0, 1, 2, 3, c.n - 1, c.n - 2, (c.n + 1) // 2, (30 << 252) - c.n, 15 << 252] + [rnd.randrange(c.n) for _ in range(12)]
        for k in ks:
            P = pm.w_mul(c, rnd.randrange(1, c.n), (c.gx, c.gy))
            for pt in (be(P[0]) + be(P[1]), be(c.gx) + be(c.gy)):
                st = hs.hs_wmul(curve, be(k), pt, 0, 0, 0, o)
                assert (o.raw, bytes([st])) == pm.batch_mul(curve, be(k), pt, 1, 0, 0), (curve, hex(k))
                st = hs.hs_wmul(curve, be(k), None, 0, 1, 0, o)      # fixed base (P-256: the accumulator-machine kernel)
                assert (o.raw, bytes([st])) == pm.batch_mul(curve, be(k), None, 1, 0), (curve, hex(k))


def test_zero_test_filter_of_the_secp256k1_ladder_rare_path(pm):
    """The ladder's mixed addition replaces its three exact zero tests by a two-compare filter and computes the exact flag word
    only in a warp the filter flags (vm.cuh: vm_jac_madd_1<FILTER>, vm_madd_flags).  A false alarm has probability 2^-31 per
    addition, so this build makes the filter fire on EVERY addition: results must not change, edge scalars included."""
    import random
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    out_dir = os.path.join(ROOT, "tests", "hostsim", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostsim_filter.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-DFEC_VM_FILTER_ALWAYS", "-x", "c++", "-shared", "-fPIC", "-o", so, src], check=True)
    hs = ctypes.CDLL(so)
    rnd = random.Random(29)
    be = lambda v: v.to_bytes(32, "big")
    hs.hs_set_var_impl(1)
    o = ctypes.create_string_buffer(33)
    for curve in (0, 1):  # the P-256 ladder and the accumulator-machine fixed-base loop use the same filter
        c = pm.CURVES[curve]
        ks = [0, 1, 2, 3, c.n - 1, c.n - 2, (c.n + 1) // 2, (30 << 252) - c.n, 15 << 252] + [rnd.randrange(c.n) for _ in range(12)]
        for k in ks:
            P = pm.w_mul(c, rnd.randrange(1, c.n), (c.gx, c.gy))
            for pt in (be(P[0]) + be(P[1]), be(c.gx) + be(c.gy)):
                st = hs.hs_wmul(curve, be(k), pt, 0, 0, 0, o)
                assert (o.raw, bytes([st])) == pm.batch_mul(curve, be(k), pt, 1, 0, 0), (curve, hex(k))
            st = hs.hs_wmul(curve, be(k), None, 0, 1, 0, o)
            assert (o.raw, bytes([st])) == pm.batch_mul(curve, be(k), None, 1, 0), (curve, hex(k))


def test_p256_weak_form_on_raw_limbs(hostsim, pm):
    """FeP256 (r2) keeps any representative below 2^256: every operation on raw Montgomery-domain limbs, with operands at the edges
    -- 0, p, p + 1, 2^256 - 1, the pairs that trigger the second correction of add / sub (both operands above p; b - a > p) --
    against the integers: result below 2^256 and congruent; mul / sqr results times 2^256 congruent to the product; canon and to_plain
    below p; is_zero true exactly for 0 and p."""
    import random
    rnd = random.Random(37)
    p = pm.PRIMES[1]
    R = 2**256
    Rinv = pow(R, -1, p)
    le = lambda v: v.to_bytes(32, "little")
    o = ctypes.create_string_buffer(32)

    def run(op, a, b=0):
        hostsim.hs_p256_raw(op, le(a), le(b), o)
        return int.from_bytes(o.raw, "little")

    edge = [0, 1, 2, p - 2, p - 1, p, p + 1, p + 2, R - 1, R - 2, R - p, R - p - 1, R - p + 1, 2**255, 2**224, 2**224 - 1, (p + 1) // 2, p // 2]
    vals = edge + [rnd.randrange(R) for _ in range(50)] + [rnd.randrange(p, R) for _ in range(10)]
    for a in vals:
        for b in vals:
            assert run(0, a, b) % p == a * b * Rinv % p, ("mul", hex(a), hex(b))
            assert run(2, a, b) % p == (a + b) % p, ("add", hex(a), hex(b))
            assert run(3, a, b) % p == (a - b) % p, ("sub", hex(a), hex(b))
        assert run(1, a) % p == a * a * Rinv % p, ("sqr", hex(a))
        assert run(4, a) % p == (-a) % p, ("neg", hex(a))
        assert run(5, a) % p == a * pow(2, -1, p) % p, ("half", hex(a))
        assert run(6, a) == a % p, ("canon", hex(a))
        assert run(7, a) == (1 if a % p == 0 else 0), ("is_zero", hex(a))
        assert run(8, a) == a * Rinv % p, ("to_plain", hex(a))
        assert run(9, a) % p == 3 * a % p and run(10, a) % p == 8 * a % p, ("small", hex(a))


def test_x25519_ladder_field_forms(pm):
    """k_x25519.cu builds Fe25519 with FEC_25519_EVOD_REDUCE (products reduced from their unmerged accumulators) and
    FEC_25519_FOLD255 (results of mul / sqr folded at 2^255, below 2^255 + 2^11; the ladder's additions then need no fold and its
    subtractions one): the same device code on the host -- raw result ranges, add_r1 / sub_r1 on operands at the edges of that range
    (including the pair that makes the addition leave 2^256: the branch no ladder ever takes), and the ladder on the RFC 7748 edge inputs."""
    import random
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    out_dir = os.path.join(ROOT, "tests", "hostsim", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostsim_x255.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-DFEC_25519_EVOD_REDUCE", "-DFEC_25519_FOLD255", "-x", "c++", "-shared", "-fPIC", "-o", so, src], check=True)
    hs = ctypes.CDLL(so)
    rnd = random.Random(31)
    p = 2**255 - 19
    le = lambda v: v.to_bytes(32, "little")
    o = ctypes.create_string_buffer(32)
    top = 2**255 + 2**11 - 1
    edge = [0, 1, 2, 18, 19, 20, p - 1, p, p + 1, 2**255 - 1, 2**255, 2**255 + 1, top, top - 1, 2**254, 2**255 - 2**12, 2**255 - 2**12 + 1]
    weak = [2**256 - 1, 2**256 - 38, 2**256 - 39, 2 * p, 2 * p + 1] + [rnd.randrange(2**256) for _ in range(40)]
    for a in edge + weak:                              # mul / sqr: any weak operand in, range and value out
        for b in edge + weak[:12]:
            hs.hs_f25519_r1(2, le(a), le(b), o)
            r = int.from_bytes(o.raw, "little")
            assert r <= top and r % p == (a * b) % p, (hex(a), hex(b), hex(r))
        hs.hs_f25519_r1(3, le(a), le(a), o)
        r = int.from_bytes(o.raw, "little")
        assert r <= top and r % p == (a * a) % p, (hex(a), hex(r))
    r1 = edge + [rnd.randrange(top + 1) for _ in range(60)]
    for a in r1:                                       # add_r1 / sub_r1 on operands of that range
        for b in r1:
            hs.hs_f25519_r1(0, le(a), le(b), o)
            assert int.from_bytes(o.raw, "little") == (a + b) % p, ("add", hex(a), hex(b))
            hs.hs_f25519_r1(1, le(a), le(b), o)
            assert int.from_bytes(o.raw, "little") == (a - b) % p, ("sub", hex(a), hex(b))
    p25519 = p
    us = [0, 1, p25519 - 1, p25519, p25519 + 1, 2**255 - 1, 2**256 - 1, 9,
          0x00b8495f16056286fdb1329ceb8d09da6ac49ff1fae35616aeb8413b7c7aebe0]
    for uv in us + [rnd.randrange(2**256) for _ in range(6)]:
        for k in (bytes(32), b"\xff" * 32, bytes(range(32)), bytes([rnd.randrange(256) for _ in range(32)])):
            hs.hs_x25519(k, le(uv), o)
            assert o.raw == pm.x25519(k, le(uv)), (hex(uv), k.hex())
    # the Ed25519 comb (k_ed25519.cu builds with FEC_25519_FOLD255 too: ed_madd_signed uses add_r1 / sub_r1), both table forms
    for k in [bytes(32), b"\xff" * 32, bytes(range(32))] + [bytes(rnd.randrange(256) for _ in range(32)) for _ in range(8)]:
        for seed in (1, 0):
            hs.hs_ed25519_pubkey(k, seed, o)
            assert o.raw == pm.ed25519_pubkey(k, bool(seed)), (k.hex(), seed)
            hs.hs_ed25519_pubkey_ct(k, seed, o)
            assert o.raw == pm.ed25519_pubkey(k, bool(seed)), (k.hex(), seed)


def test_karatsuba_variant_of_the_multiplier(pm):
    """u256.cuh keeps a one-level Karatsuba 8x8 multiplier behind FEC_KARATSUBA (measured slower on B200: it trades 16
    IMAD.WIDE for ~70 additions); it must stay correct, including equal and extreme halves."""
    import random
    so = os.path.join(ROOT, "tests", "hostsim", "_build", "libhostsim_karatsuba.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-DFEC_KARATSUBA", "-x", "c++", "-shared", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")], check=True)
    hs = ctypes.CDLL(so)
    rnd = random.Random(9)
    for prime in (0, 1, 2):
        p = pm.PRIMES[prime]
        vals = [0, 1, p - 1, (1 << 128) - 1, 1 << 128, ((1 << 128) - 1) << 128, ((1 << 128) | 1) * 12345 % p, (2**256 - 1) % p]
        vals += [rnd.randrange(p) for _ in range(40)]
        for _ in range(3000):
            a, b = rnd.choice(vals) % p, rnd.choice(vals) % p
            o = ctypes.create_string_buffer(32)
            hs.hs_field_op(prime, 0, a.to_bytes(32, "little"), b.to_bytes(32, "little"), o)
            assert int.from_bytes(o.raw, "little") == a * b % p, (prime, hex(a), hex(b))


def test_device_kdf_code_on_host_matches_model(hostsim, pm):
    """SHA-256 / HMAC / HKDF of csrc/kdf.cuh (compiled for the host) against hashlib/hmac, incl. RFC 5869 test case 3-like
    shapes: empty info, info that crosses block boundaries, every output length up to 64."""
    import random
    rnd = random.Random(3)
    for info_len in (0, 1, 22, 23, 31, 32, 54, 55, 56, 64, 100, 119, 120, 256):
        secret = bytes(rnd.randrange(256) for _ in range(32))
        info = bytes(rnd.randrange(256) for _ in range(info_len))
        for out_len in (1, 16, 31, 32, 33, 48, 64):
            o = ctypes.create_string_buffer(out_len)
            hostsim.hs_kdf(1, secret, info, info_len, out_len, o)
            assert o.raw == pm.hkdf_sha256_zero_salt(secret, info, out_len), (info_len, out_len)
            hostsim.hs_kdf(2, secret, info, info_len, out_len, o)
            assert o.raw == pm.kdf_xor_info(secret, info, out_len)


def test_quirk_device_code_on_host_matches_literal_oracle(hostsim, pm):
    """FEC_SEM_FORGE_QUIRK (csrc/quirk_k1.cuh, compiled for the host) against oracle/ref, the literal restatement
    of forge-ec's own secp256k1 code, and against the transliteration goldens of SURVEY.md Appendix A.2."""
    import json
    import random
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    oref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_ref.so"))
    lit = json.load(open(os.path.join(ROOT, "tests", "golden", "forge_ec_literal.json")))
    o = ctypes.create_string_buffer(33)
    for ks, exp in lit["secp256k1_kG_compressed"].items():
        assert hostsim.hs_quirk_k1(int(ks).to_bytes(32, "big"), None, 0, o) == 0 and o.raw.hex() == exp
    rnd = random.Random(5)
    c = pm.CURVES[0]
    # the reference's N constant has its two top limbs swapped (secp256k1.rs:27-28): that is the bound its parser applies
    N_REF = (0xFFFFFFFFFFFFFFFE << 192) | (0xFFFFFFFFFFFFFFFF << 128) | (c.n & ((1 << 128) - 1))
    ks = [0, 1, 2, c.n - 1, c.n, N_REF - 1, N_REF, 2**256 - 1] + [rnd.randrange(N_REF) for _ in range(6)]
    for k in ks:
        kb = k.to_bytes(32, "big")
        P = pm.w_mul(c, rnd.randrange(1, c.n), (c.gx, c.gy))
        for xy in (None, P[0].to_bytes(32, "big") + P[1].to_bytes(32, "big")):
            exp = ctypes.create_string_buffer(33)
            oref.oref_secp256k1_batch_mul(kb, xy, ctypes.c_size_t(1), exp, 1)
            st = hostsim.hs_quirk_k1(kb, xy, 0, o)
            assert o.raw == exp.raw, (hex(k), xy is None)
            assert st == (1 if k >= N_REF else 3 if k == 0 else 0)
    # the other two encodings carry the same x (and the y whose parity the compressed prefix reports)
    kb = (0xDEADBEEF).to_bytes(32, "big")
    o65, o64 = ctypes.create_string_buffer(65), ctypes.create_string_buffer(64)
    hostsim.hs_quirk_k1(kb, None, 0, o); hostsim.hs_quirk_k1(kb, None, 1, o65); hostsim.hs_quirk_k1(kb, None, 2, o64)
    assert o65.raw[0] == 4 and o65.raw[1:33] == o.raw[1:] == o64.raw[:32] and o65.raw[33:] == o64.raw[32:]
    assert (o64.raw[63] & 1) == o.raw[0] - 2


def test_quirk_p256_device_code_on_host_matches_literal_oracle(hostsim, pm):
    """FEC_SEM_FORGE_QUIRK for P-256 (csrc/quirk_p256.cuh, compiled for the host) against oracle/ref/oracle_ref_p256.c and
    the transliteration goldens of SURVEY.md Appendix A.2; arbitrary (also off-curve) input points, as the reference's
    from_affine validates nothing."""
    import json
    import random
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    oref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_ref.so"))
    lit = json.load(open(os.path.join(ROOT, "tests", "golden", "forge_ec_literal.json")))
    o = ctypes.create_string_buffer(33)
    for ks, exp in lit["p256_kG_compressed"].items():
        assert hostsim.hs_quirk_p256(int(ks).to_bytes(32, "big"), None, 0, o) == 0 and o.raw.hex() == exp
    rnd = random.Random(6)
    c = pm.CURVES[1]
    ks = [0, 1, 2, c.n - 1, c.n, 2**256 - 1] + [rnd.randrange(c.n) for _ in range(6)]
    for k in ks:
        kb = k.to_bytes(32, "big")
        P = pm.w_mul(c, rnd.randrange(1, c.n), (c.gx, c.gy))
        junk = rnd.randrange(c.p).to_bytes(32, "big") + rnd.randrange(c.p).to_bytes(32, "big")
        for xy in (None, P[0].to_bytes(32, "big") + P[1].to_bytes(32, "big"), junk):
            exp = ctypes.create_string_buffer(33)
            oref.oref_p256_batch_mul(kb, xy, ctypes.c_size_t(1), exp, 1)
            st = hostsim.hs_quirk_p256(kb, xy, 0, o)
            assert o.raw == exp.raw, (hex(k), xy is None)
            assert st == (1 if k >= c.n else 3 if o.raw == bytes(33) else 0)
    kb = (0xDEADBEEF).to_bytes(32, "big")
    o65, o64 = ctypes.create_string_buffer(65), ctypes.create_string_buffer(64)
    hostsim.hs_quirk_p256(kb, None, 0, o); hostsim.hs_quirk_p256(kb, None, 1, o65); hostsim.hs_quirk_p256(kb, None, 2, o64)
    assert o65.raw[0] == 4 and o65.raw[1:33] == o.raw[1:] == o64.raw[:32] and o65.raw[33:] == o64.raw[32:]
    assert (o64.raw[63] & 1) == o.raw[0] - 2


def test_quirk_x25519_device_code_on_host_matches_literal_oracle(hostsim):
    """FEC_SEM_FORGE_QUIRK for X25519 (csrc/quirk_x25519.cuh, compiled for the host) against oracle/ref/oracle_ref_x25519.c
    and the transliteration goldens of SURVEY.md Appendix A.2."""
    import json
    import random
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    oref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_ref.so"))
    lit = json.load(open(os.path.join(ROOT, "tests", "golden", "forge_ec_literal.json")))
    o, exp = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
    for ks, want in lit["x25519_base_point"].items():
        hostsim.hs_quirk_x25519(bytes.fromhex(ks), None, o)
        assert o.raw.hex() == want
    rnd = random.Random(8)
    special_u = [bytes(32), (1).to_bytes(32, "little"), (9).to_bytes(32, "little"), b"\xff" * 32, (2**255 - 19).to_bytes(32, "big")]
    for i in range(40):
        k = bytes(rnd.randrange(256) for _ in range(32))
        u = special_u[i] if i < len(special_u) else bytes(rnd.randrange(256) for _ in range(32))
        oref.oref_x25519_batch(k, u, ctypes.c_size_t(1), exp, 1)
        hostsim.hs_quirk_x25519(k, u, o)
        assert o.raw == exp.raw, (k.hex(), u.hex())


def test_quirk_ed25519_device_code_on_host_matches_literal_oracle(hostsim):
    """FEC_SEM_FORGE_QUIRK for Ed25519 key generation (csrc/quirk_ed25519.cuh, compiled for the host) against
    oracle/ref/oracle_ref_ed25519.c and the transliteration goldens of SURVEY.md Appendix A.2."""
    import json
    import random
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    oref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_ref.so"))
    lit = json.load(open(os.path.join(ROOT, "tests", "golden", "forge_ec_literal.json")))
    o, exp = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
    for seed, want in lit["ed25519_derive_public_key"].items():
        hostsim.hs_quirk_ed25519(bytes.fromhex(seed), o)
        assert o.raw.hex() == want
    rnd = random.Random(10)
    for i in range(12):
        seed = bytes(rnd.randrange(256) for _ in range(32)) if i else bytes(32)
        oref.oref_ed25519_batch_pubkey(seed, ctypes.c_size_t(1), exp, 1)
        hostsim.hs_quirk_ed25519(seed, o)
        assert o.raw == exp.raw, seed.hex()


def _gloo_worker(rank, world, port, n, tmp):
    import importlib
    import sys

    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fec = importlib.import_module("forge-ec_b200")
    import pymodel as pm

    first, count = fec.shard_range(n, world, rank)
    # every rank generates ONLY its own lanes (the Philox counter is the global lane index) ...
    k = pm.random_scalars_be(0, count, stream=3, first_lane=first)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_std.so"))
    out = np.empty(count * 33, dtype=np.uint8)
    lib.ostd_batch_mul(0, ctypes.c_void_p(k.ctypes.data), None, 0, ctypes.c_size_t(count), 0, ctypes.c_void_p(out.ctypes.data), None, 1)
    # ... and the results are gathered on rank 0 (the path's only collective step, and an optional one)
    pad = torch.zeros(((n + world - 1) // world + 1024) * 33, dtype=torch.uint8)
    pad[: out.size] = torch.from_numpy(out)
    bufs = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the max-over-ranks timing reduction of bench.py
    if rank == 0:
        parts = [bufs[r][: fec.shard_range(n, world, r)[1] * 33] for r in range(world)]
        np.save(tmp, torch.cat(parts).numpy())
        assert t.item() == float(world)
    dist.destroy_process_group()


def test_two_rank_sharding_gloo(fec, pm, tmp_path):
    import torch.multiprocessing as mp

    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    n, world = 3000, 2
    tmp = str(tmp_path / "gathered.npy")
    mp.spawn(_gloo_worker, args=(world, 29591, n, tmp), nprocs=world, join=True)
    got = np.load(tmp)
    k = pm.random_scalars_be(0, n, stream=3)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_std.so"))
    exp = np.empty(n * 33, dtype=np.uint8)
    lib.ostd_batch_mul(0, ctypes.c_void_p(k.ctypes.data), None, 0, ctypes.c_size_t(n), 0, ctypes.c_void_p(exp.ctypes.data), None, 4)
    assert np.array_equal(got, exp)


def _build_cpp_traits_test():
    """compiles tests/cpp/test_traits.cpp (the C++ host side over the C ABI, include/forge_ec.hpp) against the built library"""
    out_dir = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "test_traits")
    lib_dir = os.path.join(ROOT, "forge-ec_b200", "lib")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "test_traits.cpp"), "-L", lib_dir, "-lforge_ec_cuda", "-Wl,-rpath," + lib_dir],
                   check=True)
    return exe


def test_cpp_host_side_compiles_and_fails_loudly_without_a_device():
    """The C++ mirror of the trait surface builds against include/*.h and links every symbol it uses; without a GPU its
    Engine constructor must throw UnsupportedOperation (exit code 77 of the test program), never compute on the CPU."""
    exe = _build_cpp_traits_test()
    r = subprocess.run([exe], capture_output=True, text=True)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 77 and "no usable" in r.stdout, (r.returncode, r.stdout, r.stderr)
