"""The oracle's hash-to-curve path (oracle/h2c_oracle.hpp) pinned by the vectors the reference's own tests hold
(RFC 9380 / draft-16: tests/expand_msg.rs, tests/hash_to_curve_g1.rs, tests/hash_to_curve_g2.rs, extracted into
tests/golden/h2c_vectors.json; src/hash_to_curve/map_g1.rs::test_simple_swu_expected in kat.json), by Python's hashlib,
and by reference-independent properties of the generated constants (the isogeny lands on E)."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests import pyref, util

GOLD = os.path.join(os.path.dirname(__file__), "golden")
VEC = json.load(open(os.path.join(GOLD, "h2c_vectors.json")))
KAT = json.load(open(os.path.join(GOLD, "kat.json")))
P = pyref.P


def xmd_py(msg, dst, n):
    """RFC 9380 section 5.3.1 / 5.3.3 on hashlib — independent of the oracle"""
    if len(dst) > 255:
        dst = hashlib.sha256(b"H2C-OVERSIZE-DST-" + dst).digest()
    dp = dst + bytes([len(dst)])
    b0 = hashlib.sha256(bytes(64) + msg + n.to_bytes(2, "big") + b"\0" + dp).digest()
    b = [hashlib.sha256(b0 + b"\1" + dp).digest()]
    for i in range(2, (n + 31) // 32 + 1):
        b.append(hashlib.sha256(bytes(x ^ y for x, y in zip(b0, b[-1])) + bytes([i]) + dp).digest())
    return b"".join(b)[:n]


def test_sha256_and_xmd_against_hashlib(orc):
    rng = np.random.default_rng(8100)
    for n in [0, 1, 3, 55, 56, 57, 63, 64, 65, 119, 120, 127, 128, 1000, 4097]:
        m = rng.bytes(n)
        assert orc.sha256(m) == hashlib.sha256(m).digest()
    for msg_len, dst_len, n in [(0, 0, 1), (5, 1, 32), (100, 255, 33), (7, 256, 64), (64, 700, 128), (300, 16, 255),
                                (1, 43, 256), (0, 43, 8160)]:
        msg, dst = rng.bytes(msg_len), rng.bytes(dst_len)
        assert orc.expand_message_xmd(msg, dst, n).tobytes() == xmd_py(msg, dst, n)
    with pytest.raises(ValueError):
        orc.expand_message_xmd(b"", b"dst", 8161)          # ell > 255: the reference panics (expand_msg.rs:243)


@pytest.mark.parametrize("fn", ["expand_msg_xmd_works_for_draft16_testvectors_sha256",
                                "expand_msg_xmd_works_for_draft16_testvectors_sha256_long_dst"])
def test_expand_message_rfc_vectors(orc, fn):     # tests/expand_msg.rs:50-358
    v = VEC["expand_msg.rs::" + fn]
    dst = bytes.fromhex(v["dst"])
    for c in v["cases"]:
        got = orc.expand_message_xmd(bytes.fromhex(c["msg"]), dst, c["len_in_bytes"]).tobytes()
        assert got.hex() == c["uniform_bytes"]


@pytest.mark.parametrize("k,fn,encode", [
    (1, "hash_to_curve_g1.rs::encode_to_curve_works_for_draft16_testvectors_g1_sha256_nu", True),
    (1, "hash_to_curve_g1.rs::hash_to_curve_works_for_draft16_testvectors_g1_sha256_ro", False),
    (2, "hash_to_curve_g2.rs::encode_to_curve_works_for_draft16_testvectors_g2_sha256_nu", True),
    (2, "hash_to_curve_g2.rs::hash_to_curve_works_for_draft16_testvectors_g2_sha256_ro", False)])
def test_hash_to_curve_rfc_vectors(orc, k, fn, encode):   # tests/hash_to_curve_g1.rs:30-167, _g2.rs:30-187
    v = VEC[fn]
    G = orc.G1 if k == 1 else orc.G2
    msgs = [bytes.fromhex(c["msg"]) for c in v["cases"]]
    pr = orc.hash_to_curve(k, msgs, bytes.fromhex(v["dst"]), encode=encode, threads=4)
    xy, inf = G.to_affine(pr)
    for i, c in enumerate(v["cases"]):
        assert G.to_uncompressed(xy[i], inf[i]).tobytes().hex() == c["expected"]
    assert (G.checks(xy, inf) == 3).all()          # on the curve and in the prime-order subgroup


def _L(key, i):
    return np.array([int(x, 16) for x in KAT[key][i]], dtype=np.uint64)


def test_simple_swu_kats(orc):       # src/hash_to_curve/map_g1.rs:654-759
    k = "hash_to_curve/map_g1.rs::test_simple_swu_expected"
    want = np.concatenate([_L(k, 0), _L(k, 1), _L(k, 2)])
    zero = np.zeros(6, np.uint64)
    assert np.array_equal(orc.h2c_stage("g1_sswu", zero)[0], want)
    assert np.array_equal(orc.h2c_stage("g1_sswu", _L(k, 3))[0], want)                 # sqrt(-1/XI), positive
    neg = orc.h2c_stage("g1_sswu", _L(k, 4))[0]                                         # ... negative: y flips
    assert np.array_equal(neg[:6], want[:6]) and np.array_equal(neg[12:], want[12:])
    assert np.array_equal(neg[6:12], orc.tower(1, "neg", want[6:12])[0])
    assert np.array_equal(orc.h2c_stage("g1_sswu", _L(k, 5))[0], np.concatenate([_L(k, 6), _L(k, 7), _L(k, 8)]))


def test_sgn0(orc):                  # src/hash_to_curve/map_g1.rs:789-806, map_g2.rs:533-600
    half = (P - 1) // 2
    m = pyref.to_mont
    vals = [0, 1, P - 1, half, half + 1]
    assert list(orc.sgn0(1, np.stack([m(v) for v in vals]))) == [0, 1, 0, 1, 0]
    f2 = lambda a, b: np.concatenate([m(a), m(b)])
    cases = [(0, 0, 0), (1, 0, 1), (half, 0, 1), (half, 1, 1), (0, half, 1), (1, half, 1), (half + 1, 0, 0),
             (half + 1, 1, 0), (0, half + 1, 0), (1, half + 1, 1)]
    got = orc.sgn0(2, np.stack([f2(a, b) for a, b, _ in cases]))
    assert list(got) == [c for _, _, c in cases]


def test_from_okm_against_python(orc):
    rng = np.random.default_rng(8200)
    okm = np.frombuffer(rng.bytes(64 * 20), np.uint8).reshape(20, 64).copy()
    okm[0] = 0xff
    okm[1] = 0
    got = orc.fp_from_okm(okm)
    for i in range(20):
        assert pyref.from_mont(got[i]) == int.from_bytes(okm[i].tobytes(), "big") % P


def test_constants_without_the_reference(orc):
    """The generated isogeny / curve constants checked by what they must DO, not by where they came from: SSWU lands on
    E': y^2 = x^3 + A'x + B', the isogeny sends E' to E (y^2 = x^3 + 4, resp. 4(1+u)), cofactor clearing lands in the
    r-torsion.  Any wrong coefficient breaks these with overwhelming probability.  (map_g1.rs::test_osswu_semirandom)"""
    c = json.load(open(os.path.join(os.path.dirname(GOLD), "..", "tools", "h2c_constants.json")))
    rng = np.random.default_rng(8300)
    # --- G1, in Python integers
    A, B = int(c["g1"]["SSWU_ELLP_A"][0], 16), int(c["g1"]["SSWU_ELLP_B"][0], 16)
    u = util.rand_fp(rng, 16)
    pts = orc.h2c_stage("g1_sswu", u)
    for row in pts:
        x, y, z = (pyref.from_mont(row[6 * i:6 * i + 6]) for i in range(3))
        assert (y * y * z - (x ** 3 + A * x * z * z + B * z ** 3)) % P == 0
    iso = orc.h2c_stage("g1_iso_map", pts)
    xy, inf = orc.G1.to_affine(iso)
    assert (orc.G1.checks(xy, inf) & 1).all() and not inf.any()
    cl = orc.h2c_stage("g1_clear_cofactor", iso)
    assert (orc.G1.checks(*orc.G1.to_affine(cl)) == 3).all()
    assert np.array_equal(orc.h2c_stage("g1_map_to_curve", u), iso)
    # --- G2: on E' via the oracle's Fp2 arithmetic, on E via is_on_curve
    u2 = util.rand_fp(rng, 8, 2)
    pts2 = orc.h2c_stage("g2_sswu", u2)
    m = pyref.to_mont
    A2 = np.concatenate([m(int(v, 16)) for v in c["g2"]["SSWU_ELLP_A"]])
    B2 = np.concatenate([m(int(v, 16)) for v in c["g2"]["SSWU_ELLP_B"]])
    T = lambda op, a, b=None: orc.tower(2, op, a, b)
    for row in pts2:
        x, y, z = row[:12], row[12:24], row[24:]
        zz = T("square", z)
        lhs = T("mul", T("square", y), z)
        rhs = T("add", T("add", T("mul", T("square", x), x), T("mul", T("mul", A2, x), zz)), T("mul", T("mul", B2, zz), z))
        assert np.array_equal(lhs, rhs)
    iso2 = orc.h2c_stage("g2_iso_map", pts2)
    xy2, inf2 = orc.G2.to_affine(iso2)
    assert (orc.G2.checks(xy2, inf2) & 1).all() and not inf2.any()
    cl2 = orc.h2c_stage("g2_clear_cofactor", iso2)
    assert (orc.G2.checks(*orc.G2.to_affine(cl2)) == 3).all()
    # clear_cofactor == multiplication by the effective cofactor h_eff of RFC 9380 section 8.8.2 (G2) / 1 - z (G1)
    z = -0xd201000000010000
    s = np.frombuffer(((1 - z) % pyref.Q).to_bytes(32, "little"), np.uint8)
    # (1 - z) reduced mod q is NOT the same scalar on the cofactor part, so compare through the subgroup part only:
    # for a point already in G1, clear_cofactor(P) = [1 - z]P
    g = orc.G1.mul(orc.G1.generator(), util.rand_scalars(rng, 1))
    assert np.array_equal(orc.G1.to_affine(orc.h2c_stage("g1_clear_cofactor", g))[0], orc.G1.to_affine(orc.G1.mul(g, s))[0])


def test_batch_layout_and_edge_messages(orc):
    dst = b"QUUX-V01-CS02-with-BLS12381G1_XMD:SHA-256_SSWU_RO_"
    msgs = [b"", b"a", b"", b"x" * 1000, b"abc"]
    both = orc.hash_to_curve(1, msgs, dst, threads=3)
    for i, m in enumerate(msgs):
        assert np.array_equal(both[i], orc.hash_to_curve(1, [m], dst)[0])
    assert np.array_equal(both[0], both[2]) and not np.array_equal(both[0], both[1])
    long_dst = b"D" * 300
    a = orc.hash_to_curve(2, [b"msg"], long_dst)
    b = orc.hash_to_curve(2, [b"msg"], hashlib.sha256(b"H2C-OVERSIZE-DST-" + long_dst).digest())
    assert np.array_equal(a, b)


def test_hash_to_scalar(orc):          # src/hash_to_curve/map_scalar.rs:25-45
    Q = pyref.Q
    for c in VEC["map_scalar.rs::test_hash_to_scalar"]:
        got = orc.fr_from_okm(np.frombuffer(bytes.fromhex(c["okm"]), np.uint8))
        assert int.from_bytes(orc.scalar_to_bytes(got)[0].tobytes(), "little") == int(c["expected"], 16)
    assert not orc.fr_from_okm(np.zeros(48, np.uint8)).any()            # the all-zero case of the same test
    rng = np.random.default_rng(8400)
    okm = np.frombuffer(rng.bytes(48 * 50), np.uint8).reshape(50, 48).copy()
    okm[0] = 0xff
    got = orc.scalar_to_bytes(orc.fr_from_okm(okm))
    for i in range(50):
        assert int.from_bytes(got[i].tobytes(), "little") == int.from_bytes(okm[i].tobytes(), "big") % Q
    # hash_to_field: count scalars per message from one expand_message_xmd call (src/hash_to_curve/mod.rs:41-66)
    msgs, dst = [b"", b"abc", b"x" * 100], b"QUUX-V01-CS02-with-BLS12381SCALAR_XMD:SHA-256_"
    for count in (1, 2, 3):
        out = orc.scalar_to_bytes(orc.fr_hash_to_field(msgs, dst, count))
        for i, m in enumerate(msgs):
            okm_i = xmd_py(m, dst, 48 * count)
            for c in range(count):
                want = int.from_bytes(okm_i[48 * c:48 * c + 48], "big") % Q
                assert int.from_bytes(out[i * count + c].tobytes(), "little") == want
