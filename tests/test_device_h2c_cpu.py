"""The device-side hash-to-curve source (bls12_381_b200/csrc/h2c.cuh: SHA-256 / expand_message_xmd, hash_to_field, SSWU,
isogenies, cofactor clearing and the two-kernel launch plan) run on the HOST through tests/emul/ against the RFC 9380
vectors of the reference's tests and against the oracle — limb-exact on the projective coordinates.
The real-GPU counterpart is tests/test_gpu_zz_h2c.py."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from tests import util
from tests.emul import build as emul_build
from tests.test_oracle_h2c import VEC, KAT, xmd_py, _L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EmulH2C:
    def __init__(self):
        self.lib = C.CDLL(emul_build.build("default"))

    def expand(self, orc, msgs, dst, n):
        cat, off = orc.pack_messages(msgs)
        d = np.frombuffer(bytes(dst) or b"\0", np.uint8).copy()
        out = np.empty((len(msgs), n), np.uint8)
        rc = self.lib.emul_h2c_expand(_p(cat), _p(off), C.c_size_t(len(msgs)), _p(d), C.c_size_t(len(dst)), n, _p(out))
        assert rc == 0
        return out

    def hash(self, orc, k, msgs, dst, encode=False, threads=4):
        cat, off = orc.pack_messages(msgs)
        d = np.frombuffer(bytes(dst) or b"\0", np.uint8).copy()
        out = np.empty((len(msgs), 18 * k), np.uint64)
        rc = self.lib.emul_h2c_hash(k, _p(cat), _p(off), C.c_size_t(len(msgs)), _p(d), C.c_size_t(len(dst)), int(encode),
                                    _p(out), threads)
        assert rc == 2
        return out

    def stage(self, k, kind, a):
        a = np.ascontiguousarray(a, np.uint64)
        n = a.shape[0]
        out = np.empty((n, 18 * k), np.uint64)
        assert self.lib.emul_h2c_stage(k, kind, _p(a), C.c_size_t(n), _p(out)) == 0
        return out


@pytest.fixture(scope="module")
def em():
    return EmulH2C()


def test_expand_message(em, orc):
    for fn in ["expand_msg_xmd_works_for_draft16_testvectors_sha256", "expand_msg_xmd_works_for_draft16_testvectors_sha256_long_dst"]:
        v = VEC["expand_msg.rs::" + fn]
        dst = bytes.fromhex(v["dst"])
        for n in (0x20, 0x80):
            cases = [c for c in v["cases"] if c["len_in_bytes"] == n]
            got = em.expand(orc, [bytes.fromhex(c["msg"]) for c in cases], dst, n)
            for g, c in zip(got, cases):
                assert g.tobytes().hex() == c["uniform_bytes"]
    rng = np.random.default_rng(9100)
    msgs = [rng.bytes(int(l)) for l in [0, 1, 54, 55, 56, 63, 64, 65, 119, 120, 121, 500]]
    for dst, n in ((b"", 1), (b"x" * 255, 33), (b"y" * 256, 64), (b"tag", 256), (b"tag", 300)):
        got = em.expand(orc, msgs, dst, n)
        for g, m in zip(got, msgs):
            assert g.tobytes() == xmd_py(m, dst, n)


@pytest.mark.parametrize("k,fn,encode", [
    (1, "hash_to_curve_g1.rs::encode_to_curve_works_for_draft16_testvectors_g1_sha256_nu", True),
    (1, "hash_to_curve_g1.rs::hash_to_curve_works_for_draft16_testvectors_g1_sha256_ro", False),
    (2, "hash_to_curve_g2.rs::encode_to_curve_works_for_draft16_testvectors_g2_sha256_nu", True),
    (2, "hash_to_curve_g2.rs::hash_to_curve_works_for_draft16_testvectors_g2_sha256_ro", False)])
def test_rfc_vectors_through_the_launch_plan(em, orc, k, fn, encode):
    v = VEC[fn]
    G = orc.G1 if k == 1 else orc.G2
    msgs = [bytes.fromhex(c["msg"]) for c in v["cases"]]
    dst = bytes.fromhex(v["dst"])
    pr = em.hash(orc, k, msgs, dst, encode=encode)
    assert np.array_equal(pr, orc.hash_to_curve(k, msgs, dst, encode=encode, threads=4))     # limb-exact (x, y, z)
    xy, inf = G.to_affine(pr)
    for i, c in enumerate(v["cases"]):
        assert G.to_uncompressed(xy[i], inf[i]).tobytes().hex() == c["expected"]


def test_stages_against_oracle(em, orc):
    rng = np.random.default_rng(9200)
    k = "hash_to_curve/map_g1.rs::test_simple_swu_expected"
    u1 = np.concatenate([util.rand_fp(rng, 12), np.zeros((1, 6), np.uint64), _L(k, 3)[None], _L(k, 4)[None], _L(k, 5)[None]])
    s = em.stage(1, 0, u1)
    assert np.array_equal(s, orc.h2c_stage("g1_sswu", u1))
    assert np.array_equal(s[-1], np.concatenate([_L(k, 6), _L(k, 7), _L(k, 8)]))             # the reference's KAT
    iso = em.stage(1, 1, s)
    assert np.array_equal(iso, orc.h2c_stage("g1_iso_map", s))
    assert np.array_equal(em.stage(1, 2, u1), iso)
    assert np.array_equal(em.stage(1, 3, iso), orc.h2c_stage("g1_clear_cofactor", iso))
    u2 = np.concatenate([util.rand_fp(rng, 6, 2), np.zeros((1, 12), np.uint64)])
    s2 = em.stage(2, 0, u2)
    assert np.array_equal(s2, orc.h2c_stage("g2_sswu", u2))
    iso2 = em.stage(2, 1, s2)
    assert np.array_equal(iso2, orc.h2c_stage("g2_iso_map", s2))
    assert np.array_equal(em.stage(2, 2, u2), iso2)
    assert np.array_equal(em.stage(2, 3, iso2), orc.h2c_stage("g2_clear_cofactor", iso2))
    # cofactor clearing of the identity and of a subgroup point
    ident = orc.G2.identity(1)
    assert np.array_equal(em.stage(2, 3, ident), orc.h2c_stage("g2_clear_cofactor", ident))


def test_batches_with_ragged_messages(em, orc):
    rng = np.random.default_rng(9300)
    msgs = [rng.bytes(int(l)) for l in rng.integers(0, 200, 40)] + [b"", b"", b"a" * 1000]
    for k, dst in ((1, b"BLS_SIG_BLS12381G1_XMD:SHA-256_SSWU_RO_NUL_"), (2, b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_NUL_")):
        sub = msgs if k == 1 else msgs[:8] + msgs[-3:]
        for encode in (False, True):
            assert np.array_equal(em.hash(orc, k, sub, dst, encode=encode, threads=8),
                                  orc.hash_to_curve(k, sub, dst, encode=encode, threads=8))
    long_dst = b"Q" * 400
    assert np.array_equal(em.hash(orc, 1, msgs[:3], long_dst), orc.hash_to_curve(1, msgs[:3], long_dst))
