"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs — bit-exact (integer work).  Also the reference's own KATs straight on the GPU."""
import json
import os

import numpy as np
import pytest

from tests import pyref, util

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def eng():
    import bls12_381_b200
    e = bls12_381_b200.Engine()
    yield e
    e.close()


def eq(a, b):
    return np.array_equal(np.asarray(a, np.uint64).reshape(-1), np.asarray(b, np.uint64).reshape(-1))


# ----------------------------------------------------------------------------- field tower
@pytest.mark.parametrize("op", ["mul", "add", "sub", "square", "neg", "invert"])
def test_fp_ops(eng, orc, op):
    rng = np.random.default_rng(100)
    e = util.edge_fp()
    a = np.concatenate([util.rand_fp(rng, 2048), np.repeat(e, len(e), 0)])
    b = np.concatenate([util.rand_fp(rng, 2048), np.tile(e, (len(e), 1))])
    if op == "invert":
        a, b = a[:300], None
    elif op in ("square", "neg"):
        b = None
    got = eng.tower(1, op, a, b)
    assert eq(got, orc.tower(1, op, a, b))


def test_fp_kat_on_gpu(eng):
    kat = json.load(open(os.path.join(GOLD, "kat.json")))
    L = lambda k, i: np.array([int(x, 16) for x in kat[k][i]], dtype=np.uint64)
    k = "fp.rs::test_multiplication"
    assert eq(eng.tower(1, "mul", L(k, 0), L(k, 1)), L(k, 2))          # src/fp.rs:721-749
    k = "fp.rs::test_squaring"
    assert eq(eng.tower(1, "square", L(k, 0)), L(k, 1))                # src/fp.rs:699-719
    k = "fp.rs::test_inversion"
    assert eq(eng.tower(1, "invert", L(k, 0)), L(k, 1))                # src/fp.rs:919-940
    k = "fp2.rs::test_multiplication"
    c = lambda i: np.concatenate([L(k, i), L(k, i + 1)])
    assert eq(eng.tower(2, "mul", c(0), c(2)), c(4))                   # src/fp2.rs:464-522


@pytest.mark.parametrize("level,ops,n", [
    (2, ["mul", "add", "sub", "square", "neg", "invert", "frobenius", "conjugate", "mul_by_nonresidue"], 600),
    (6, ["mul", "add", "sub", "square", "neg", "invert", "frobenius", "mul_by_nonresidue"], 200),
    (12, ["mul", "square", "invert", "frobenius", "conjugate", "cyclotomic_square"], 100),
])
def test_tower_ops(eng, orc, level, ops, n):
    rng = np.random.default_rng(200 + level)
    a, b = util.rand_fp(rng, n, level), util.rand_fp(rng, n, level)
    a[0] = 0
    a[1, :] = 0
    a[1, :6] = orc.R_LIMBS
    for op in ops:
        bb = b if op in ("mul", "add", "sub") else None
        aa = a[2:] if op == "invert" else a
        assert eq(eng.tower(level, op, aa, None if bb is None else bb[:aa.shape[0]]),
                  orc.tower(level, op, aa, None if bb is None else bb[:aa.shape[0]])), (level, op)


# ----------------------------------------------------------------------------- group ops, limb-exact
@pytest.mark.parametrize("k", [1, 2])
def test_group_ops(eng, orc, k):
    rng = np.random.default_rng(300 + k)
    G = orc.G1 if k == 1 else orc.G2
    n = 96
    pr, xy, inf = util.rand_points(orc, k, rng, n)
    p = util.randomize_z(orc, k, rng, pr)
    q = np.roll(util.randomize_z(orc, k, rng, pr), 7, 0)
    # edge cases: identity operands, P + P, P + (-P)
    ident = G.identity()
    p[0], q[1] = ident, ident
    p[2], q[2] = ident, ident
    q[3] = p[3]
    neg = p[4].copy()
    w = 6 * k
    neg[w:2 * w] = orc.tower(k, "neg", p[4, w:2 * w])
    q[4] = neg
    assert eq(eng.double_batch(k, p), G.double(p))
    assert eq(eng.add_batch(k, p, q), G.add(p, q))
    qxy, qinf = xy.copy(), inf.copy()
    qinf[5] = 1
    qxy[6] = G.to_affine(p[6:7])[0]                       # P + P through the mixed formula
    assert eq(eng.add_mixed_batch(k, p, qxy, qinf), G.add_mixed(p, qxy, qinf))
    # batch_normalize, identities included (src/g1.rs:1690-1727)
    gx, gi = eng.batch_normalize(k, p)
    ox, oi = G.batch_normalize(p)
    assert eq(gx, ox) and eq(gi, oi)
    gx, gi = eng.batch_normalize(k, p[:1])
    assert gi[0] == 1
    gx, gi = eng.batch_normalize(k, np.concatenate([p] * 30))
    assert eq(gx, np.concatenate([ox] * 30)) and eq(gi, np.concatenate([oi] * 30))


@pytest.mark.parametrize("k,n", [(1, 1024), (2, 192), (1, 6500)])
def test_mul_batch_config1(eng, orc, k, n):
    """BASELINE config 1: 1024 x G1Projective * Scalar, raw (x,y,z) limb-exact vs multiply() (src/g1.rs:754)."""
    rng = np.random.default_rng(400 + k)
    G = orc.G1 if k == 1 else orc.G2
    pr, _, _ = util.rand_points(orc, k, rng, n)
    p = util.randomize_z(orc, k, rng, pr)
    s = util.rand_scalars(rng, n)
    s[0] = 0
    s[1] = util.scalar_bytes(1)
    s[2] = util.scalar_bytes(pyref.Q - 1)
    p[3] = G.identity()
    got = eng.mul_batch(k, p, s)
    exp = G.mul(p, s, threads=8)
    assert eq(got, exp)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 2])
@pytest.mark.parametrize("groups", [-1, 1, 2, 3, 5, 6])
def test_mul_batch_items_per_warp(eng, orc, k, groups):
    """every shape of the scalar-multiplication batch (tuning key mul_groups: thread per item, 1..5 items per warp of the
    six-lane group kernel, one warp per item) gives the reference's raw (x, y, z) limbs — ragged last warp, zero / one / q-1
    scalars, an identity point, items of one warp with different scalar bits"""
    rng = np.random.default_rng(470 + 10 * k + groups)
    G = orc.G1 if k == 1 else orc.G2
    n = 43 if k == 1 else 17
    pr, _, _ = util.rand_points(orc, k, rng, n)
    p = util.randomize_z(orc, k, rng, pr)
    s = util.rand_scalars(rng, n)
    s[0] = 0
    s[1] = util.scalar_bytes(1)
    s[2] = util.scalar_bytes(pyref.Q - 1)
    p[3] = G.identity()
    s[5] = 0
    eng.set_tuning("mul_groups", groups)
    try:
        got = eng.mul_batch(k, p, s)
    finally:
        eng.set_tuning("mul_groups", 0)
    assert eq(got, G.mul(p, s, threads=8))


# ----------------------------------------------------------------------------- MSM
def _msm_case(eng, orc, k, xy, inf, s, cs=(0,)):
    G = orc.G1 if k == 1 else orc.G2
    n = xy.shape[0]
    if n <= 600:
        exp = G.msm_naive(xy, inf, s, threads=8)
    else:
        exp = G.msm_pippenger(xy, inf, s, c=10, threads=8)
    ea = G.to_affine(exp)
    for c in cs:
        eng.set_msm_window(c)
        got = eng.msm(k, xy, inf, s)
        ga = G.to_affine(got)
        assert eq(ga[0], ea[0]) and ga[1][0] == ea[1][0], (k, n, c)
    eng.set_msm_window(0)


@pytest.mark.parametrize("k", [1, 2])
def test_msm_small_and_edges(eng, orc, k):
    rng = np.random.default_rng(500 + k)
    G = orc.G1 if k == 1 else orc.G2
    _, xy, inf = util.rand_points(orc, k, rng, 300)
    s = util.rand_scalars(rng, 300)
    # n = 0 -> identity
    got = eng.msm(k, xy[:0], None, s[:0])
    assert G.to_affine(got)[1][0] == 1
    for n in (1, 2, 3, 37, 300):
        _msm_case(eng, orc, k, xy[:n], inf[:n], s[:n], cs=(0, 4, 7, 13) if n <= 37 else (0, 9))
    # edge cases: zero scalars, q-1, identity points, duplicates, P and -P, all-same scalar
    n = 64
    xy2, inf2, s2 = xy[:n].copy(), inf[:n].copy(), s[:n].copy()
    s2[0] = 0
    s2[1] = util.scalar_bytes(pyref.Q - 1)
    s2[2] = util.scalar_bytes(1)
    inf2[3] = 1
    xy2[5] = xy2[4]
    s2[5] = s2[4]                                   # duplicate point, same scalar -> same bucket (P + P)
    xy2[7] = xy2[6]
    w = 6 * k
    xy2[7, w:] = orc.tower(k, "neg", xy2[6, w:])
    s2[7] = s2[6]                                   # P and -P with the same scalar (P + (-P) in a bucket)
    s2[8:24] = s2[8]                                # all-same scalar
    _msm_case(eng, orc, k, xy2, inf2, s2, cs=(0, 5, 8, 16))
    # everything cancels -> identity
    xy3 = np.concatenate([xy[:8], xy[:8]])
    xy3[8:, w:] = orc.tower(k, "neg", xy[:8, w:])
    s3 = np.concatenate([s[:8], s[:8]])
    got = eng.msm(k, xy3, None, s3)
    assert G.to_affine(got)[1][0] == 1


@pytest.mark.parametrize("k,n", [(1, 5000), (2, 2000)])
def test_msm_medium(eng, orc, k, n):
    rng = np.random.default_rng(600 + k)
    _, xy, inf = util.rand_points(orc, k, rng, n)
    s = util.rand_scalars(rng, n)
    _msm_case(eng, orc, k, xy, inf, s, cs=(0, 11))


# ----------------------------------------------------------------------------- pairings
@pytest.mark.parametrize("variant", [7, 4])      # 7 = six lanes per pairing (default), 4 = one thread per pairing
def test_pairing_parity(eng, orc, variant):
    eng.set_tuning("pairing_variant", variant)
    try:
        _pairing_parity(eng, orc, 700 + variant, 43 if variant == 7 else 40)
    finally:
        eng.set_tuning("pairing_variant", 0)


def _pairing_parity(eng, orc, seed, n):
    rng = np.random.default_rng(seed)
    _, pxy, pinf = util.rand_points(orc, 1, rng, n)
    _, qxy, qinf = util.rand_points(orc, 2, rng, n)
    pinf[1] = 1
    qinf[2] = 1
    pinf[3], qinf[3] = 1, 1
    ml = eng.miller_loop_batch(pxy, pinf, qxy, qinf)
    oml = orc.miller_loop(pxy, pinf, qxy, qinf, threads=8)
    assert eq(ml, oml)                                           # MillerLoopResult limb-exact
    fe = eng.final_exponentiation_batch(ml)
    ofe = orc.final_exponentiation(oml, threads=8)
    assert eq(fe, ofe)
    assert eq(eng.pairing_batch(pxy, pinf, qxy, qinf), ofe)      # == pairing() (src/pairings.rs:607)
    assert eq(orc.pairing(pxy, pinf, qxy, qinf, threads=8), ofe)
    # product mode: multi_miller_loop over prepared terms, identities skipped (src/pairings.rs:554-603)
    mm = eng.multi_miller_loop(pxy[:9], pinf[:9], qxy[:9], qinf[:9])
    omm = orc.multi_miller_loop(pxy[:9], pinf[:9], qxy[:9], qinf[:9])
    assert eq(mm, omm)
    assert eq(eng.final_exponentiation_batch(mm), orc.final_exponentiation(omm))
    one = np.zeros(72, np.uint64)
    one[:6] = orc.R_LIMBS
    assert eq(eng.multi_miller_loop(pxy[:0], None, qxy[:0], None), one)    # MillerLoopResult::default()


def test_msm_rejects_non_canonical_scalars(eng, orc):
    """Scalar::to_bytes() is canonical (< q): raw strings >= q are refused (B200_EINVAL), q - 1 is fine"""
    from bls12_381_b200 import B200Error
    rng = np.random.default_rng(7200)
    n = 300
    q = pyref.Q
    for k in (1, 2):
        _, xy, inf = util.rand_points(orc, k, rng, n)
        s = util.rand_scalars(rng, n)
        for bad in (q, q + 7, (1 << 256) - 1):
            s2 = s.copy()
            s2[123] = np.frombuffer(bad.to_bytes(32, "little"), np.uint8)
            with pytest.raises(B200Error):
                eng.msm(k, xy, inf, s2)
        s2 = s.copy()
        s2[123] = np.frombuffer((q - 1).to_bytes(32, "little"), np.uint8)
        G = orc.G1 if k == 1 else orc.G2
        assert eq(G.to_affine(eng.msm(k, xy, inf, s2))[0], G.to_affine(G.msm_pippenger(xy, inf, s2, c=8, threads=8))[0])


def test_pairing_products_shared_squaring(eng, orc):
    """multi_miller_loop in its reference shape — one squaring of the accumulator per bit for ALL terms (src/pairings.rs:
    554-603) — for n in {0, 1, 2, 9, 1000} with identity terms, and batches of 3- / 4-term products (Groth16 shape)"""
    rng = np.random.default_rng(7100)
    n = 1000
    _, pxy, pinf = util.rand_points(orc, 1, rng, n)
    _, qxy, qinf = util.rand_points(orc, 2, rng, n)
    for i in (1, 17, 500):
        pinf[i] = 1
    for i in (2, 17, 999):
        qinf[i] = 1
    one = np.zeros(72, np.uint64)
    one[:6] = orc.R_LIMBS
    for m in (0, 1, 2, 9, 1000):
        got = eng.multi_miller_loop(pxy[:m], pinf[:m], qxy[:m], qinf[:m]).reshape(-1)
        want = orc.multi_miller_loop(pxy[:m], pinf[:m], qxy[:m], qinf[:m]).reshape(-1) if m else one
        assert eq(got, want), m
    for terms in (3, 4):
        npr = 60
        a = tuple(x[:npr * terms] for x in (pxy, pinf, qxy, qinf))
        want = np.concatenate([orc.multi_miller_loop(*(x[i * terms:(i + 1) * terms] for x in a)).reshape(1, 72)
                               for i in range(npr)])
        assert eq(eng.pairing_product_batch(*a, terms, final_exp=False), want)
        assert eq(eng.pairing_product_batch(*a, terms, final_exp=True), orc.final_exponentiation(want, threads=8))


def test_gt_generator_kat_on_gpu(eng, orc):
    kat = json.load(open(os.path.join(GOLD, "kat.json")))
    exp = np.concatenate([np.array([int(x, 16) for x in g], dtype=np.uint64) for g in kat["pairings.rs::generator"]])
    gxy, ginf = orc.G1.to_affine(orc.G1.generator())
    hxy, hinf = orc.G2.to_affine(orc.G2.generator())
    assert eq(eng.pairing_batch(gxy, None, hxy, None), exp)      # src/pairings.rs:827-832


def test_imad_peak_runs(eng):
    v, ms = eng.imad_peak(500)
    print("IMAD.WIDE peak: %.3e /s (%.3f ms)" % (v, ms))
    assert v > 1e12


# ----------------------------------------------------------------------------- (de)serialization
@pytest.mark.parametrize("k", [1, 2])
def test_serialization_golden_files(eng, orc, k):
    """[i]G, i < 1000, serialized ON THE GPU must reproduce the reference's .dat golden files byte for byte
    (src/tests/mod.rs:3-76), and the goldens must deserialize on the GPU to the same points."""
    dat = np.load(os.path.join(GOLD, "dat_vectors.npz"))
    unc = dat["g%d_uncompressed" % k].reshape(1000, 96 * k)
    cmp_ = dat["g%d_compressed" % k].reshape(1000, 48 * k)
    G = orc.G1 if k == 1 else orc.G2
    # [i]G on the GPU: prefix sums by repeated mixed addition would be serial; use i as the scalar instead
    s = np.zeros((1000, 32), np.uint8)
    s[:, 0] = np.arange(1000) & 0xff
    s[:, 1] = np.arange(1000) >> 8
    xy, inf = eng.batch_normalize(k, eng.mul_batch(k, np.repeat(G.generator(), 1000, 0), s))
    assert inf[0] == 1 and not inf[1:].any()
    assert np.array_equal(eng.serialize(k, xy, inf, compressed=False), unc)
    assert np.array_equal(eng.serialize(k, xy, inf, compressed=True), cmp_)
    for compressed, data in ((False, unc), (True, cmp_)):
        dxy, dinf, st = eng.deserialize(k, data, compressed=compressed)
        assert (st == 3).all() and eq(dinf, inf) and eq(dxy, xy)


@pytest.mark.parametrize("k", [1, 2])
def test_deserialization_rejects(eng, orc, k):
    G = orc.G1 if k == 1 else orc.G2
    rng = np.random.default_rng(900 + k)
    _, xy, inf = util.rand_points(orc, k, rng, 8)
    unc = eng.serialize(k, xy, inf, compressed=False)
    cmp_ = eng.serialize(k, xy, inf, compressed=True)
    for i in range(8):
        assert np.array_equal(unc[i], G.to_uncompressed(xy[i], inf[i])) and np.array_equal(cmp_[i], G.to_compressed(xy[i], inf[i]))
    bad_u, bad_c = unc.copy(), cmp_.copy()
    bad_u[0, 0] |= 0x80                          # compression flag on an uncompressed encoding
    bad_u[1, 0] |= 0x20                          # sort flag on an uncompressed encoding
    bad_u[2, :48 * k] = 0xff                     # non-canonical x (>= p)
    bad_u[2, 0] = 0x1f
    bad_u[3, -1] ^= 1                            # y off the curve: Some, but not on the curve
    bad_u[4, 0] |= 0x40                          # infinity flag with non-zero coordinates
    bad_c[0, 0] &= 0x7f                          # compression flag missing
    bad_c[1, :] = 0
    bad_c[1, 0] = 0xc0                           # the identity, valid
    bad_c[2, :] = 0
    bad_c[2, 0] = 0xe0                           # identity with the sort flag: invalid
    _, _, st_u = eng.deserialize(k, bad_u, compressed=False)
    dxy, dinf, st_c = eng.deserialize(k, bad_c, compressed=True)
    for i in range(8):
        ok, p, pinf = G.from_uncompressed(bad_u[i])          # oracle: unchecked + on-curve in one flag
        assert bool(st_u[i] == 3) == ok, ("uncompressed", i, st_u[i])
        ok, p, pinf = G.from_compressed(bad_c[i])
        assert bool(st_c[i] & 1) == ok, ("compressed", i, st_c[i])
        if ok:
            assert dinf[i] == pinf and (pinf or eq(dxy[i], p))
    assert st_u[3] == 1 and st_u[0] == 0 and st_u[1] == 0 and st_u[2] == 0 and st_u[4] == 0
    assert st_c[0] == 0 and st_c[1] == 3 and dinf[1] == 1 and st_c[2] == 0


@pytest.mark.parametrize("mode", [2, 3, 4])
def test_g2_msm_bucket_kernel_variants(eng, orc, mode):
    """the three G2 bucket kernels (accumulator in registers / in shared memory, 3 or 2 blocks per SM) agree with
    the oracle, including the same-x exceptional cases"""
    rng = np.random.default_rng(1000)
    k = 2
    _, xy, inf = util.rand_points(orc, k, rng, 200)
    s = util.rand_scalars(rng, 200)
    xy[5] = xy[4]
    s[5] = s[4]                                    # P + P inside a bucket
    xy[7] = xy[6]
    xy[7, 12:] = orc.tower(2, "neg", xy[6, 12:])
    s[7] = s[6]                                    # P + (-P) inside a bucket
    s[8:40] = s[8]                                 # one crowded bucket per window
    inf[9] = 1
    eng.set_tuning("g2_acc_blocks", mode)
    try:
        _msm_case(eng, orc, k, xy, inf, s, cs=(0, 6, 12))
    finally:
        eng.set_tuning("g2_acc_blocks", 4)


def test_glv_decompose(eng):
    """k = k1 + k2*lambda (mod q) with |k1|, |k2| < 2^127 for random and extreme scalars (csrc/glv.cuh)"""
    z = 0xd201000000010000
    lam = z * z - 1
    assert lam * lam + lam + 1 == pyref.Q
    rng = np.random.default_rng(77)
    vals = [0, 1, 2, pyref.Q - 1, pyref.Q - 2, lam, lam + 1, lam - 1, pyref.Q // 2, pyref.Q // 2 + 1, (pyref.Q + 1) // 2 - 1,
            lam * lam, lam * (lam // 2), (lam // 2) * (lam + 1), 1 << 254, pyref.Q - lam, lam // 2, lam // 2 + 1,
            (lam // 2) * lam + lam // 2, (lam // 2 + 1) * lam - 1, (1 << 128) - 1, 1 << 128, 1 << 127]
    vals += [int.from_bytes(rng.bytes(40), "little") % pyref.Q for _ in range(4000)]
    s = np.stack([util.scalar_bytes(v) for v in vals])
    for v, (k1, k2) in zip(vals, eng.glv_decompose(s)):
        assert (k1 + k2 * lam - v) % pyref.Q == 0, hex(v)
        assert abs(k1) < (1 << 127) and abs(k2) < (1 << 127), hex(v)


@pytest.mark.parametrize("glv", [0, 1])
def test_g1_msm_with_and_without_glv(eng, orc, glv):
    rng = np.random.default_rng(1200)
    _, xy, inf = util.rand_points(orc, 1, rng, 700)
    s = util.rand_scalars(rng, 700)
    s[0] = 0
    s[1] = util.scalar_bytes(pyref.Q - 1)
    s[2] = util.scalar_bytes(0xac45a4010001a40200000000ffffffff)          # lambda itself
    s[3] = util.scalar_bytes(0xac45a4010001a40200000000ffffffff + 1)
    xy[5], s[5] = xy[4], s[4]
    inf[6] = 1
    eng.set_tuning("g1_glv", glv)
    try:
        _msm_case(eng, orc, 1, xy, inf, s, cs=(0, 5, 8, 13, 16))
        _msm_case(eng, orc, 1, xy[:3], inf[:3], s[:3], cs=(0, 4))
    finally:
        eng.set_tuning("g1_glv", 0)


@pytest.mark.parametrize("k", [1, 2])
def test_subgroup_checks(eng, orc, k):
    """is_on_curve / is_torsion_free on the device vs the oracle, including the reference's KAT of a curve point
    outside the q-order subgroup (src/g1.rs:1598-1623, src/g2.rs:1862-1907)"""
    kat = json.load(open(os.path.join(GOLD, "kat.json")))
    key = "g%d.rs::test_is_torsion_free" % k
    bad = np.concatenate([np.array([int(x, 16) for x in g], dtype=np.uint64) for g in kat[key][:2 * k]])
    G = orc.G1 if k == 1 else orc.G2
    rng = np.random.default_rng(1300 + k)
    _, xy, inf = util.rand_points(orc, k, rng, 20)
    xy = np.concatenate([xy, bad[None, :]])
    inf = np.concatenate([inf, [0]]).astype(np.uint8)
    inf[3] = 1
    xy[5, -1] ^= 1                                   # off the curve
    got = eng.check(k, xy, inf)
    exp = G.checks(xy, inf)
    assert np.array_equal(got, exp)
    assert got[-1] == 1 and got[0] == 3 and got[3] == 3 and (got[5] & 1) == 0


def test_g2_prepared(eng, orc):
    """G2Prepared coefficients limb-exact vs the oracle (src/pairings.rs:504-546, 68 triples :539), and
    multi_miller_loop over prepared terms, identities skipped (:554-603)"""
    rng = np.random.default_rng(1400)
    n = 7
    _, pxy, pinf = util.rand_points(orc, 1, rng, n)
    _, qxy, qinf = util.rand_points(orc, 2, rng, n)
    qinf[2] = 1
    pinf[4] = 1
    co = eng.g2_prepare(qxy, qinf)
    for i in range(n):
        assert eq(co[i], orc.g2_prepare(qxy[i], qinf[i])), i
    mm = eng.multi_miller_loop_prepared(pxy, pinf, co, qinf)
    assert eq(mm, orc.multi_miller_loop(pxy, pinf, qxy, qinf))
    assert eq(mm, eng.multi_miller_loop(pxy, pinf, qxy, qinf))          # == the unprepared product mode
    import torch
    dev = torch.device("cuda", eng.device)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64) if a.dtype == np.uint64 else a).to(dev)
    dco = torch.empty((n, 68 * 36), dtype=torch.int64, device=dev)
    eng.g2_prepare_dev(t(qxy), t(qinf), n, dco)
    assert eq(dco.cpu().numpy().view(np.uint64), co)
    ml = torch.empty((n, 72), dtype=torch.int64, device=dev)
    eng.miller_loop_prepared_batch_dev(t(pxy), t(pinf), dco, t(qinf), n, ml)
    assert eq(ml.cpu().numpy().view(np.uint64), orc.miller_loop(pxy, pinf, qxy, qinf, threads=4))


@pytest.mark.gpu
@pytest.mark.parametrize("prepare_max", [0, 1000000])
def test_g2_prepared_both_kernels(eng, orc, prepare_max):
    """the two kernels behind b200_g2_prepare — one thread per Q (pairing_v4.cu, large batches) and six lanes per Q
    (k_coop_g2_prepare, small batches; tuning key coop_prepare_max picks) — give the same 68 coefficient triples as the oracle,
    identity included (prepared as the generator, src/pairings.rs:528-544); ragged batch (not a multiple of 5 Qs per warp)"""
    rng = np.random.default_rng(1410 + (prepare_max > 0))
    n = 23
    _, qxy, qinf = util.rand_points(orc, 2, rng, n)
    qinf[5] = 1
    eng.set_tuning("coop_prepare_max", prepare_max)
    try:
        co = eng.g2_prepare(qxy, qinf)
    finally:
        eng.set_tuning("coop_prepare_max", 5000)
    for i in range(n):
        assert eq(co[i], orc.g2_prepare(qxy[i], qinf[i])), i


@pytest.mark.parametrize("k,n", [(1, 40000), (2, 12000)])
def test_msm_giant_buckets(eng, orc, k, n):
    """skewed scalars: all-equal scalars put every point of a window into ONE bucket (>= 1023 points -> split across
    blocks by k_msm_giant_parts); half-equal/half-random mixes giant and ordinary buckets"""
    import time
    rng = np.random.default_rng(1500 + k)
    base = util.rand_points(orc, k, rng, 64)
    reps = (n + 63) // 64
    xy = np.tile(base[1], (reps, 1))[:n].copy()
    inf = np.tile(base[2], reps)[:n].copy()
    same = util.rand_scalars(rng, 1)
    s = np.repeat(same, n, 0)
    G = orc.G1 if k == 1 else orc.G2
    for variant in range(2):
        if variant == 1:
            s[n // 2:] = rng.integers(0, 256, (n - n // 2, 32), dtype=np.uint8)
            s[n // 2:, 31] &= 0x3f
        t0 = time.perf_counter()
        got = G.to_affine(eng.msm(k, xy, inf, s))
        dt = time.perf_counter() - t0
        exp = G.to_affine(G.msm_pippenger(xy, inf, s, c=12, threads=16))
        assert eq(got[0], exp[0]) and got[1][0] == exp[1][0], (k, variant)
        assert dt < 2.0, "giant buckets must not serialise on one thread (took %.2f s)" % dt


def test_fp_invert_fast(eng, orc):
    """binary-GCD inverse (csrc/fp_inv.cuh) == Fermat inverse == oracle, incl. 0, 1, p-1, R"""
    rng = np.random.default_rng(1600)
    a = np.concatenate([util.rand_fp(rng, 4000), util.edge_fp()])
    got = eng.tower(1, "invert_fast", a)
    assert eq(got, orc.tower(1, "invert", a))
    assert eq(got, eng.tower(1, "invert", a))


@pytest.mark.parametrize("k", [1, 2])
@pytest.mark.parametrize("levels", [1, 2, 3])
def test_msm_batched_affine_levels(eng, orc, k, levels):
    """batched-affine tree levels in front of the bucket kernel (csrc/msm_affine.cuh): same group element, including
    tangent pairs (P + P), cancelling pairs (P + (-P)), identity operands inside the tree, odd bucket populations,
    empty buckets and giant buckets"""
    rng = np.random.default_rng(1700 + k)
    n = 900 if k == 1 else 500
    _, xy, inf = util.rand_points(orc, k, rng, n)
    s = util.rand_scalars(rng, n)
    w = 6 * k
    xy[5], s[5] = xy[4], s[4]                      # P + P at level 1 (adjacent only by chance -> several copies)
    xy[6], s[6] = xy[4], s[4]
    xy[7], s[7] = xy[4], s[4]
    xy[9] = xy[8]
    xy[9, w:] = orc.tower(k, "neg", xy[8, w:])
    s[9] = s[8]                                    # P + (-P)
    xy[11] = xy[10]
    xy[11, w:] = orc.tower(k, "neg", xy[10, w:])
    s[11] = s[10]
    xy[12], s[12] = xy[10], s[10]                  # (P + (-P)) + P : identity operand at level 2
    s[20:60] = s[20]                               # a crowded bucket
    s[0] = 0
    inf[30] = 1
    eng.set_tuning("msm_affine_levels", levels)
    try:
        _msm_case(eng, orc, k, xy, inf, s, cs=(4, 6, 9))
        _msm_case(eng, orc, k, xy[:3], inf[:3], s[:3], cs=(4,))
        if levels == 3 and k == 1:                 # giant buckets next to the affine path
            big = np.tile(xy[:64], (40, 1))
            sb = np.repeat(s[100:101], 2560, 0)
            sb[1280:] = rng.integers(0, 256, (1280, 32), dtype=np.uint8)
            sb[1280:, 31] &= 0x3f
            _msm_case(eng, orc, k, big, None, sb, cs=(4, 8))
    finally:
        eng.set_tuning("msm_affine_levels", 0)


@pytest.mark.parametrize("k", [1, 2])
def test_msm_through_byte_encodings(eng, orc, k):
    """the call sequence of the Rust wrapper (bindings/rust/bls12381-b200/src/lib.rs), which marshals ONLY through the
    reference's public byte encodings: to_uncompressed bytes -> deserialize (GPU) -> MSM -> batch_normalize ->
    serialize (GPU) == to_uncompressed(sum_i p_i * s_i) computed by the oracle"""
    G = orc.G1 if k == 1 else orc.G2
    rng = np.random.default_rng(1800 + k)
    n = 150
    _, xy, inf = util.rand_points(orc, k, rng, n)
    inf[3] = 1
    s = util.rand_scalars(rng, n)
    enc = np.stack([G.to_uncompressed(xy[i], inf[i]) for i in range(n)])          # what G*Affine::to_uncompressed() gives
    dxy, dinf, st = eng.deserialize(k, enc, compressed=False)
    assert (st == 3).all()
    res = eng.msm(k, dxy, dinf, s)
    axy, ainf = eng.batch_normalize(k, res)
    got = eng.serialize(k, axy, ainf, compressed=False)[0]
    exp_aff = G.to_affine(G.msm_naive(xy, inf, s, threads=8))
    assert np.array_equal(got, G.to_uncompressed(exp_aff[0][0], exp_aff[1][0]))
    # and the batched validation of untrusted compressed encodings (from_compressed semantics: decode + subgroup)
    cmp_ = np.stack([G.to_compressed(xy[i], inf[i]) for i in range(8)])
    dxy, dinf, st = eng.deserialize(k, cmp_, compressed=True)
    assert (st == 3).all() and (eng.check(k, dxy, dinf) == 3).all()


def test_two_contexts_concurrently(orc):
    """distinct b200_ctx are independent (Send + Sync on the Rust side): two engines driven from two host threads at
    the same time give the same, correct results"""
    import threading
    import bls12_381_b200
    rng = np.random.default_rng(1900)
    _, xy, inf = util.rand_points(orc, 1, rng, 3000)
    s = util.rand_scalars(rng, 3000)
    exp = orc.G1.to_affine(orc.G1.msm_pippenger(xy, inf, s, c=10, threads=8))
    results, errors = {}, []

    def work(tag):
        try:
            e = bls12_381_b200.Engine()
            for it in range(6):
                got = orc.G1.to_affine(e.msm(1, xy, inf, s))
                assert eq(got[0], exp[0]) and got[1][0] == exp[1][0], (tag, it)
            results[tag] = True
            e.close()
        except Exception as ex:                      # noqa: BLE001
            errors.append((tag, repr(ex)))

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors and len(results) == 2, errors
