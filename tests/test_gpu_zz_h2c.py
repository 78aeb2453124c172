"""GPU parity tests (-m gpu) for batched hash to curve through the C ABI: the RFC 9380 vectors the
reference's integration tests hold, and the oracle on seeded inputs — limb-exact on the projective coordinates.

STATUS: hardware-validated (round 2, first GPU call); the device source is also covered on the CPU harness
(tests/test_device_h2c_cpu.py)."""
import numpy as np
import pytest

from tests import util
from tests.test_oracle_h2c import VEC, xmd_py, _L

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.fixture(scope="module")
def eng():
    import bls12_381_b200
    e = bls12_381_b200.Engine()
    yield e
    e.close()


def test_expand_message_vectors_and_hashlib(eng):
    for fn in ["expand_msg_xmd_works_for_draft16_testvectors_sha256", "expand_msg_xmd_works_for_draft16_testvectors_sha256_long_dst"]:
        v = VEC["expand_msg.rs::" + fn]
        dst = bytes.fromhex(v["dst"])
        for n in (0x20, 0x80):
            cases = [c for c in v["cases"] if c["len_in_bytes"] == n]
            got = eng.expand_message_xmd([bytes.fromhex(c["msg"]) for c in cases], dst, n)
            for g, c in zip(got, cases):
                assert g.tobytes().hex() == c["uniform_bytes"]
    rng = np.random.default_rng(9500)
    msgs = [rng.bytes(int(l)) for l in list(range(0, 130)) + [500, 1000, 4096]]
    for dst, n in ((b"", 1), (b"x" * 255, 33), (b"y" * 256, 64), (b"tag", 256), (b"tag", 8160)):
        got = eng.expand_message_xmd(msgs, dst, n)
        for g, m in zip(got, msgs):
            assert g.tobytes() == xmd_py(m, dst, n)
    rc = eng.lib.b200_expand_message_xmd_sha256(eng.h, None, None, 0, None, 0, 8161, None)
    assert rc == -1                                   # ell > 255: the reference panics


@pytest.mark.parametrize("k,fn,encode", [
    (1, "hash_to_curve_g1.rs::encode_to_curve_works_for_draft16_testvectors_g1_sha256_nu", True),
    (1, "hash_to_curve_g1.rs::hash_to_curve_works_for_draft16_testvectors_g1_sha256_ro", False),
    (2, "hash_to_curve_g2.rs::encode_to_curve_works_for_draft16_testvectors_g2_sha256_nu", True),
    (2, "hash_to_curve_g2.rs::hash_to_curve_works_for_draft16_testvectors_g2_sha256_ro", False)])
def test_rfc_vectors(eng, orc, k, fn, encode):
    v = VEC[fn]
    G = orc.G1 if k == 1 else orc.G2
    msgs = [bytes.fromhex(c["msg"]) for c in v["cases"]]
    dst = bytes.fromhex(v["dst"])
    pr = eng.hash_to_curve(k, msgs, dst, encode=encode)
    assert np.array_equal(pr, orc.hash_to_curve(k, msgs, dst, encode=encode, threads=4))
    # canonical bytes through the library's own batch_normalize + serialize (the step after hashing)
    xy, inf = eng.batch_normalize(k, pr)
    ser = eng.serialize(k, xy, inf, compressed=False)
    for i, c in enumerate(v["cases"]):
        assert ser[i].tobytes().hex() == c["expected"]
    assert (eng.check(k, xy, inf) == 3).all()


def test_stages(eng, orc):
    rng = np.random.default_rng(9600)
    k = "hash_to_curve/map_g1.rs::test_simple_swu_expected"
    u1 = np.concatenate([util.rand_fp(rng, 300), np.zeros((1, 6), np.uint64), _L(k, 3)[None], _L(k, 4)[None], _L(k, 5)[None]])
    s = eng.h2c_stage(1, "sswu", u1)
    assert np.array_equal(s, orc.h2c_stage("g1_sswu", u1))
    assert np.array_equal(s[-1], np.concatenate([_L(k, 6), _L(k, 7), _L(k, 8)]))
    iso = eng.h2c_stage(1, "iso_map", s)
    assert np.array_equal(iso, orc.h2c_stage("g1_iso_map", s))
    assert np.array_equal(eng.h2c_stage(1, "map_to_curve", u1), iso)
    assert np.array_equal(eng.h2c_stage(1, "clear_cofactor", iso), orc.h2c_stage("g1_clear_cofactor", iso))
    u2 = np.concatenate([util.rand_fp(rng, 100, 2), np.zeros((1, 12), np.uint64)])
    s2 = eng.h2c_stage(2, "sswu", u2)
    assert np.array_equal(s2, orc.h2c_stage("g2_sswu", u2))
    iso2 = eng.h2c_stage(2, "iso_map", s2)
    assert np.array_equal(iso2, orc.h2c_stage("g2_iso_map", s2))
    assert np.array_equal(eng.h2c_stage(2, "map_to_curve", u2), iso2)
    assert np.array_equal(eng.h2c_stage(2, "clear_cofactor", iso2), orc.h2c_stage("g2_clear_cofactor", iso2))


@pytest.mark.parametrize("k,n", [(1, 3000), (2, 600)])
def test_batches_against_oracle(eng, orc, k, n):
    rng = np.random.default_rng(9700 + k)
    msgs = [rng.bytes(int(l)) for l in rng.integers(0, 200, n - 3)] + [b"", b"", b"a" * 5000]
    dst = b"BLS_SIG_BLS12381G%d_XMD:SHA-256_SSWU_RO_NUL_" % k
    for encode in (False, True):
        got = eng.hash_to_curve(k, msgs, dst, encode=encode)
        assert np.array_equal(got, orc.hash_to_curve(k, msgs, dst, encode=encode, threads=8))
    long_dst = b"Q" * 400
    assert np.array_equal(eng.hash_to_curve(k, msgs[:5], long_dst), orc.hash_to_curve(k, msgs[:5], long_dst))
    assert eng.hash_to_curve(k, [], dst).shape == (0, 18 * k)


def test_signature_shaped_flow(eng, orc):
    """hash 64 messages to G2, multiply each by a secret-free test scalar on the GPU, and check e(g1, [s]H(m)) ==
    e([s]g1, H(m)) with the pairing path — the BLS-signature shape the row exists for"""
    rng = np.random.default_rng(9800)
    msgs = [rng.bytes(32) for _ in range(64)]
    h = eng.hash_to_curve(2, msgs, b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_NUL_")
    s = util.rand_scalars(rng, 1)
    sig = eng.mul_batch(2, h, np.repeat(s, 64, 0))
    hxy, hinf = eng.batch_normalize(2, h)
    sxy, sinf = eng.batch_normalize(2, sig)
    g1 = orc.G1.generator()
    pk = orc.G1.mul(g1, s)
    gxy, ginf = orc.G1.to_affine(g1)
    pxy, pinf = orc.G1.to_affine(pk)
    lhs = eng.pairing_batch(np.repeat(gxy, 64, 0), np.repeat(ginf, 64), sxy, sinf)
    rhs = eng.pairing_batch(np.repeat(pxy, 64, 0), np.repeat(pinf, 64), hxy, hinf)
    assert np.array_equal(lhs, rhs)
    assert not np.array_equal(lhs[0], lhs[1])


def test_hash_to_scalar(eng, orc):
    rng = np.random.default_rng(9900)
    okm = np.frombuffer(rng.bytes(48 * 3000), np.uint8).reshape(3000, 48).copy()
    okm[0] = 0xff
    okm[1] = 0
    assert np.array_equal(eng.fr_from_okm(okm), orc.fr_from_okm(okm))
    for c in VEC["map_scalar.rs::test_hash_to_scalar"]:          # src/hash_to_curve/map_scalar.rs:25-45
        got = eng.fr_to_bytes(eng.fr_from_okm(np.frombuffer(bytes.fromhex(c["okm"]), np.uint8)))
        assert int.from_bytes(got[0].tobytes(), "little") == int(c["expected"], 16)
    msgs = [rng.bytes(int(l)) for l in rng.integers(0, 150, 500)]
    dst = b"QUUX-V01-CS02-with-BLS12381SCALAR_XMD:SHA-256_"
    for count in (1, 2, 5):
        assert np.array_equal(eng.fr_hash_to_field(msgs, dst, count), orc.fr_hash_to_field(msgs, dst, count))
    assert eng.lib.b200_fr_hash_to_field(eng.h, None, None, 0, None, 0, 171, None) == -1
