"""The HOST side of the scalar-field and hash-to-curve C ABI (capi_fr.cu, capi_h2c.cu: argument checks, staging, offset
rebasing, table caching, the in-place NTT path) compiled with g++ against a mock CUDA runtime (tests/emul/mock) and
driven through the SAME Engine methods and the SAME test bodies as the -m gpu tests of those rows
(tests/test_gpu_zz_fr.py, tests/test_gpu_zz_h2c.py).  Kernels run as (block, thread) loops over the device source.
What stays GPU-only: real streams/asynchrony, PTX->SASS, launch limits."""
import ctypes as C

import numpy as np
import pytest

import tests.test_gpu_zz_fr as FR
import tests.test_gpu_zz_h2c as H2C
from tests.emul import build as emul_build


@pytest.fixture(scope="module")
def eng():
    from bls12_381_b200 import _lib
    from bls12_381_b200.engine import Engine
    lib = C.CDLL(emul_build.build_cabi())
    for name, args in _lib.SIGNATURES.items():
        if hasattr(lib, name):
            f = getattr(lib, name)
            f.argtypes = args
            f.restype = _lib._RESTYPE.get(name, C.c_int)

    class MockEngine(Engine):
        def __init__(self):
            self.lib = lib
            h = C.c_void_p()
            assert lib.b200_ctx_create(0, C.byref(h)) == 0
            self.h = h

        def close(self):
            if self.h:
                lib.b200_ctx_destroy(self.h)
                self.h = None

    e = MockEngine()
    yield e
    e.close()


def _body(fn):
    return getattr(fn, "__wrapped__", fn)


@pytest.mark.parametrize("op", ["mul", "add", "sub", "square", "neg", "double", "invert"])
def test_fr_ops(eng, orc, op):
    FR.test_fr_ops(eng, orc, op)


def test_fr_kats(eng):
    FR.test_fr_kats_on_gpu(eng)


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 7, 9, 12])
def test_ntt(eng, orc, log_n):
    FR.test_ntt_parity(eng, orc, log_n)


def test_ntt_cache_and_in_place(eng, orc):
    FR.test_ntt_table_cache_switches_sizes(eng, orc)
    rng = np.random.default_rng(11)
    a = FR.rand_fr(rng, 1 << 8)
    buf = a.copy()
    # in == out through the device-pointer entry point ("device" memory is host memory under the mock)
    assert eng.lib.b200_fr_ntt_dev(eng.h, buf.ctypes.data, 8, 0, 1, buf.ctypes.data) == 0
    assert np.array_equal(buf, orc.fr_ntt(a, coset=True))
    out = np.empty_like(a)
    assert eng.lib.b200_fr_op_dev(eng.h, 0, a.ctypes.data, buf.ctypes.data, a.shape[0], out.ctypes.data) == 0
    assert np.array_equal(out, orc.fr_op("mul", a, buf))


def test_fr_argument_errors(eng):
    FR.test_fr_argument_errors(eng)
    a = np.zeros((4, 4), np.uint64)
    assert eng.lib.b200_fr_op(eng.h, 0, a.ctypes.data, None, 4, a.ctypes.data) == -1     # binary op without b
    assert eng.lib.b200_fr_op(eng.h, 0, None, None, 0, None) == 0                          # n = 0
    assert eng.lib.b200_fr_ntt(None, a.ctypes.data, 2, 0, 0, a.ctypes.data) == -1          # no ctx


def test_h2c_expand(eng):
    H2C.test_expand_message_vectors_and_hashlib(eng)


@pytest.mark.parametrize("k,fn,encode", [
    (1, "hash_to_curve_g1.rs::hash_to_curve_works_for_draft16_testvectors_g1_sha256_ro", False),
    (2, "hash_to_curve_g2.rs::encode_to_curve_works_for_draft16_testvectors_g2_sha256_nu", True)])
def test_h2c_vectors(eng, orc, k, fn, encode):
    from tests.test_oracle_h2c import VEC
    v = VEC[fn]
    G = orc.G1 if k == 1 else orc.G2
    msgs = [bytes.fromhex(c["msg"]) for c in v["cases"]]
    pr = eng.hash_to_curve(k, msgs, bytes.fromhex(v["dst"]), encode=encode)
    xy, inf = G.to_affine(pr)
    for i, c in enumerate(v["cases"]):
        assert G.to_uncompressed(xy[i], inf[i]).tobytes().hex() == c["expected"]


def test_h2c_batches_stages_and_offsets(eng, orc):
    H2C.test_batches_against_oracle(eng, orc, 1, 120)
    H2C.test_batches_against_oracle(eng, orc, 2, 12)
    # offsets that do not start at 0 (a window into a larger buffer), and a descending pair -> EINVAL
    msgs = [b"alpha", b"", b"gamma-gamma"]
    cat = np.frombuffer(b"JUNK" + b"".join(msgs) + b"TAIL", np.uint8).copy()
    off = np.array([4, 9, 9, 20], np.uint64)
    dst = np.frombuffer(b"tag", np.uint8).copy()
    out = np.empty((3, 18), np.uint64)
    assert eng.lib.b200_g1_hash_to_curve(eng.h, cat.ctypes.data, off.ctypes.data, 3, dst.ctypes.data, 3, 0, out.ctypes.data) == 0
    assert np.array_equal(out, orc.hash_to_curve(1, msgs, b"tag"))
    bad = np.array([4, 3, 9, 20], np.uint64)
    assert eng.lib.b200_g1_hash_to_curve(eng.h, cat.ctypes.data, bad.ctypes.data, 3, dst.ctypes.data, 3, 0, out.ctypes.data) == -1
    assert eng.lib.b200_h2c_stage(eng.h, 3, 0, out.ctypes.data, 1, out.ctypes.data) == -1
    rng = np.random.default_rng(12)
    from tests import util
    u = util.rand_fp(rng, 5)
    assert np.array_equal(eng.h2c_stage(1, "map_to_curve", u), orc.h2c_stage("g1_map_to_curve", u))
    u2 = util.rand_fp(rng, 2, 2)
    assert np.array_equal(eng.h2c_stage(2, "sswu", u2), orc.h2c_stage("g2_sswu", u2))


def test_hash_to_scalar(eng, orc):
    H2C.test_hash_to_scalar(eng, orc)


def test_gt_mul(eng, orc):
    from tests.test_gt_mul import _inputs
    rng, _, g, s = _inputs(orc, 4, 12400)
    assert np.array_equal(eng.gt_mul_batch(g, s), orc.gt_mul(g, s, threads=4))
    assert eng.gt_mul_batch(g[:0], s[:0]).shape == (0, 72)
    with pytest.raises(ValueError):
        eng.gt_mul_batch(g, s[:2])


def test_serialization_units(eng, orc):
    """capi_serial.cu (validated on hardware in round 1) through the mock runtime: the reference's golden .dat files
    byte for byte, deserialization round trip and rejects, subgroup checks — a CPU regression net for later edits"""
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "dat_vectors.npz"))
    for k, G in ((1, orc.G1), (2, orc.G2)):
        m = 40
        gen = G.generator()
        pts = [G.identity(1)]
        for _ in range(m - 1):
            pts.append(G.add(pts[-1], gen))
        xy, inf = G.batch_normalize(np.concatenate(pts))
        for compressed, name in ((True, "g%d_compressed" % k), (False, "g%d_uncompressed" % k)):
            w = (48 if compressed else 96) * k
            want = gold[name][:m * w].reshape(m, w)
            got = eng.serialize(k, xy, inf, compressed=compressed)
            assert np.array_equal(got, want)                                   # src/tests/mod.rs:3-76
            dxy, dinf, st = eng.deserialize(k, want, compressed=compressed)
            assert (st == 3).all() and np.array_equal(dinf, inf) and np.array_equal(dxy[inf == 0], xy[inf == 0])
        assert (eng.check(k, xy, inf) == 3).all()
        bad = gold["g%d_compressed" % k][:48 * k].copy().reshape(1, 48 * k)
        bad[0, 0] &= 0x7f                                                      # compression flag cleared
        assert eng.deserialize(k, bad, compressed=True)[2][0] & 1 == 0
