"""Pins the CPU oracle (oracle/bls12_381_oracle.hpp) against the reference's OWN known-answer tests and
golden vector files, extracted by tests/golden/make_golden.py, and against an
independent Python big-integer implementation (tests/pyref.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from tests import pyref

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KAT = json.load(open(os.path.join(GOLD, "kat.json")))
DAT = np.load(os.path.join(GOLD, "dat_vectors.npz"))


def L(key, i):
    """i-th limb group of a KAT as a uint64 row."""
    return np.array([int(x, 16) for x in KAT[key][i]], dtype=np.uint64)


def cat(key, idx):
    return np.concatenate([L(key, i) for i in idx])


def eq(a, b):
    return np.array_equal(np.asarray(a, np.uint64).reshape(-1), np.asarray(b, np.uint64).reshape(-1))


# ----------------------------------------------------------------------------- Fp   (src/fp.rs:699-979)
def test_fp_kats(orc):
    k = "fp.rs::test_squaring"            # a, expected a^2
    assert eq(orc.tower(1, "square", L(k, 0)), L(k, 1))
    k = "fp.rs::test_multiplication"      # a, b, c = a*b
    assert eq(orc.tower(1, "mul", L(k, 0), L(k, 1)), L(k, 2))
    k = "fp.rs::test_addition"
    assert eq(orc.tower(1, "add", L(k, 0), L(k, 1)), L(k, 2))
    k = "fp.rs::test_subtraction"
    assert eq(orc.tower(1, "sub", L(k, 0), L(k, 1)), L(k, 2))
    k = "fp.rs::test_negation"
    assert eq(orc.tower(1, "neg", L(k, 0)), L(k, 1))
    k = "fp.rs::test_inversion"
    assert eq(orc.tower(1, "invert", L(k, 0)), L(k, 1))
    assert eq(orc.tower(1, "invert", np.zeros(6, np.uint64)), np.zeros(6, np.uint64))
    k = "fp.rs::test_sqrt"                # a = 4 ; -sqrt(a) == 2
    ok, s = orc.fp_sqrt(L(k, 0))
    assert ok and eq(orc.tower(1, "neg", s), L(k, 1))
    k = "fp.rs::test_lexicographic_largest"   # zero, one: false ; then three elements: false,true,true
    assert not orc.fp_lex_largest(np.zeros(6, np.uint64))
    assert not orc.fp_lex_largest(orc.R_LIMBS)
    assert not orc.fp_lex_largest(L(k, 0)) and orc.fp_lex_largest(L(k, 1)) and orc.fp_lex_largest(L(k, 2))


def test_fp_against_bigints(orc):
    rng = np.random.default_rng(1)
    vals = [int.from_bytes(rng.bytes(48), "little") % pyref.P for _ in range(64)] + [0, 1, pyref.P - 1]
    a = np.stack([pyref.to_mont(v) for v in vals])
    b = np.stack([pyref.to_mont(v) for v in reversed(vals)])
    m = orc.tower(1, "mul", a, b)
    s = orc.tower(1, "sub", a, b)
    for i, (x, y) in enumerate(zip(vals, reversed(vals))):
        assert pyref.from_mont(m[i]) == x * y % pyref.P
        assert pyref.from_mont(s[i]) == (x - y) % pyref.P
        assert pyref.limbs_to_int(m[i]) < pyref.P       # canonical
    # byte encoding: big-endian canonical (src/fp.rs:851-890)
    one = orc.fp_to_bytes(orc.R_LIMBS)
    assert one[-1] == 1 and not one[:-1].any()
    ok, _ = orc.fp_from_bytes(np.frombuffer(pyref.P.to_bytes(48, "big"), np.uint8))
    assert not ok
    ok, v = orc.fp_from_bytes(np.frombuffer((pyref.P - 1).to_bytes(48, "big"), np.uint8))
    assert ok and pyref.from_mont(v[0]) == pyref.P - 1


# ----------------------------------------------------------------------------- Fp2  (src/fp2.rs:422-887)
def test_fp2_kats(orc):
    k = "fp2.rs::test_squaring"
    assert eq(orc.tower(2, "square", cat(k, [0, 1])), cat(k, [2, 3]))
    k = "fp2.rs::test_multiplication"
    assert eq(orc.tower(2, "mul", cat(k, [0, 1]), cat(k, [2, 3])), cat(k, [4, 5]))
    k = "fp2.rs::test_addition"
    assert eq(orc.tower(2, "add", cat(k, [0, 1]), cat(k, [2, 3])), cat(k, [4, 5]))
    k = "fp2.rs::test_subtraction"
    assert eq(orc.tower(2, "sub", cat(k, [0, 1]), cat(k, [2, 3])), cat(k, [4, 5]))
    k = "fp2.rs::test_negation"
    assert eq(orc.tower(2, "neg", cat(k, [0, 1])), cat(k, [2, 3]))
    k = "fp2.rs::test_inversion"
    assert eq(orc.tower(2, "invert", cat(k, [0, 1])), cat(k, [2, 3]))
    k = "fp2.rs::test_sqrt"   # groups: a.c0,a.c1 | b.c0 (c1=0) | c.c0 (c1=0) | nonsquare c0,c1  (src/fp2.rs:686-766)
    z = np.zeros(6, np.uint64)
    for x in (cat(k, [0, 1]), np.concatenate([L(k, 2), z]), np.concatenate([L(k, 3), z])):
        ok, s = orc.fp2_sqrt(x)
        assert ok and eq(orc.tower(2, "square", s), x)
    ok, _ = orc.fp2_sqrt(cat(k, [4, 5]))
    assert not ok


# ----------------------------------------------------------------------------- Fp6 / Fp12 (src/fp6.rs:375-561, src/fp12.rs:264-649)
def _tower_identities(orc, level, a, b, c, one):
    T = lambda op, x, y=None: orc.tower(level, op, x, y)
    add = (lambda x, y: T("add", x, y)) if level == 6 else None
    for x in (a, b, c):
        assert eq(T("square", x), T("mul", x, x))
    if level == 6:
        lhs = T("mul", add(a, b), T("square", c))
        rhs = add(T("mul", T("mul", c, c), a), T("mul", T("mul", c, c), b))
        assert eq(lhs, rhs)
    assert eq(T("mul", T("invert", a), T("invert", b)), T("invert", T("mul", a, b)))
    assert eq(T("mul", T("invert", a), a), one)


def test_fp6_arithmetic(orc):
    k = "fp6.rs::test_arithmetic"
    a, b, c = cat(k, range(0, 6)), cat(k, range(6, 12)), cat(k, range(12, 18))
    one = np.zeros(36, np.uint64)
    one[:6] = orc.R_LIMBS
    _tower_identities(orc, 6, a, b, c, one)
    f = a
    for _ in range(6):
        f = orc.tower(6, "frobenius", f)
    assert eq(f, a) and not eq(orc.tower(6, "frobenius", a), a)


def test_fp12_arithmetic(orc):
    k = "fp12.rs::test_arithmetic"
    a, b, c = cat(k, range(0, 12)), cat(k, range(12, 24)), cat(k, range(24, 36))
    one = np.zeros(72, np.uint64)
    one[:6] = orc.R_LIMBS
    _tower_identities(orc, 12, a, b, c, one)
    f = a
    for _ in range(12):
        f = orc.tower(12, "frobenius", f)
    assert eq(f, a) and not eq(orc.tower(12, "frobenius", a), a)
    # sparse mul == dense mul with the sparse operand embedded (src/fp12.rs:116 vs :197)
    c0, c1, c4 = a[0:12], a[12:24], a[24:36]
    sparse = np.zeros(72, np.uint64)
    sparse[0:12], sparse[12:24], sparse[48:60] = c0, c1, c4
    assert eq(orc.fp12_mul_by_014(b, c0, c1, c4), orc.tower(12, "mul", b, sparse))


# ----------------------------------------------------------------------------- G1   (src/g1.rs:1262-1540)
def test_g1_kats(orc):
    G = orc.G1
    g = G.generator()
    k = "g1.rs::test_doubling"            # affine(2G) x, y
    xy, inf = G.to_affine(G.double(g))
    assert inf[0] == 0 and eq(xy, cat(k, [0, 1]))
    assert eq(G.double(G.identity()), G.identity())
    # degenerate addition (src/g1.rs:1372-1417): a = 4G, b = (a.x*beta^2, -a.y, a.z)
    k = "g1.rs::test_projective_addition"     # groups: z, z(again?) ... beta, x, y  -> take the last three
    beta, ex, ey = L(k, len(KAT[k]) - 3), L(k, len(KAT[k]) - 2), L(k, len(KAT[k]) - 1)
    beta2 = orc.tower(1, "square", beta)
    a = G.double(G.double(g))
    b = a.copy()
    b[0, 0:6] = orc.tower(1, "mul", a[0, 0:6], beta2)
    b[0, 6:12] = orc.tower(1, "neg", a[0, 6:12])
    xy, inf = G.to_affine(G.add(a, b))
    assert inf[0] == 0 and eq(xy, np.concatenate([ex, ey]))
    # the same through add_mixed (src/g1.rs:1493-1539): affine(a) + b
    axy, ainf = G.to_affine(a)
    xy2, _ = G.to_affine(G.add_mixed(b, axy, ainf))
    assert eq(xy2, xy)
    # identity handling
    assert eq(G.add(G.identity(), g), g) or eq(G.to_affine(G.add(G.identity(), g))[0], G.to_affine(g)[0])
    ixy, iinf = G.affine_identity()
    assert eq(G.add_mixed(g, ixy, iinf), g)           # src/g1.rs:751
    xy, inf = G.to_affine(G.identity())
    assert inf[0] == 1 and eq(xy, ixy)                # src/g1.rs:49-63


def test_g2_kats(orc):
    G = orc.G2
    g = G.generator()
    k = "g2.rs::test_doubling"
    xy, inf = G.to_affine(G.double(g))
    assert inf[0] == 0 and eq(xy, cat(k, [0, 1, 2, 3]))
    # [5]G + G chain consistency between add / add_mixed / double (src/g2.rs:1478-1806 style)
    g2 = G.double(g)
    g4 = G.double(g2)
    g5 = G.add(g4, g)
    gxy, ginf = G.to_affine(g)
    g5m = G.add_mixed(g4, gxy, ginf)
    assert eq(G.to_affine(g5)[0], G.to_affine(g5m)[0])
    g6a = G.add(g5, g)
    g6b = G.double(G.add(g2, g))
    assert eq(G.to_affine(g6a)[0], G.to_affine(g6b)[0])


# ----------------------------------------------------------------------------- .dat golden files (src/tests/mod.rs:3-76)
@pytest.mark.parametrize("k", [1, 2])
def test_dat_vectors(orc, k):
    """[i]G for i = 0..999 built by repeated `+ generator` must serialize to exactly the golden bytes,
    both compressed and uncompressed, and the goldens must deserialize back to the same points."""
    G = orc.G1 if k == 1 else orc.G2
    unc = DAT["g%d_uncompressed" % k].reshape(1000, 96 * k)
    cmp_ = DAT["g%d_compressed" % k].reshape(1000, 48 * k)
    g = G.generator()
    e = G.identity()
    pts = []
    for i in range(1000):
        pts.append(e)
        e = G.add(e, g)
    xy, inf = G.batch_normalize(np.concatenate(pts))
    xy1, inf1 = G.to_affine(np.concatenate(pts))
    assert eq(xy, xy1) and eq(inf, inf1)              # batch_normalize == to_affine (src/g1.rs:1690-1727)
    for i in range(1000):
        assert np.array_equal(G.to_uncompressed(xy[i], inf[i]), unc[i]), i
        assert np.array_equal(G.to_compressed(xy[i], inf[i]), cmp_[i]), i
    for i in list(range(0, 40)) + [999]:
        ok, p, pinf = G.from_uncompressed(unc[i])
        assert ok and pinf == inf[i] and eq(p, xy[i])
        ok, p, pinf = G.from_compressed(cmp_[i])
        assert ok and pinf == inf[i] and eq(p, xy[i])


# ----------------------------------------------------------------------------- independent bigint cross-check
@pytest.mark.parametrize("k", [1, 2])
def test_group_against_bigints(orc, k):
    G, E = (orc.G1, pyref.E1) if k == 1 else (orc.G2, pyref.E2)
    gen = pyref.G1_GEN if k == 1 else pyref.G2_GEN
    assert E.is_on_curve(gen)
    gxy, ginf = G.to_affine(G.generator())
    assert E.from_affine_limbs(gxy[0]) == gen         # pyref's textbook generator == reference limbs
    rng = np.random.default_rng(7 + k)
    scal = [0, 1, 2, pyref.Q - 1, pyref.Q - 2] + [int.from_bytes(rng.bytes(32), "little") % pyref.Q for _ in range(3)]
    sb = np.stack([np.frombuffer(s.to_bytes(32, "little"), np.uint8) for s in scal])
    out = G.mul(np.repeat(G.generator(), len(scal), 0), sb)
    xy, inf = G.to_affine(out)
    for i, s in enumerate(scal):
        exp_xy, exp_inf = E.to_affine_limbs(E.mul(gen, s))
        assert inf[i] == exp_inf and eq(xy[i], exp_xy), i


# ----------------------------------------------------------------------------- scalars (src/scalar.rs:1030-1040)
def test_scalar_from_bytes_wide(orc):
    k = "scalar.rs::test_from_bytes_wide_maximum"
    mont = np.array([int(x, 16) for x in KAT[k][0]], dtype=np.uint64)
    got = orc.scalar_from_wide(np.full(64, 0xff, np.uint8))
    assert np.array_equal(got, orc.scalar_to_bytes(mont))
    v = int.from_bytes(bytes([0xff] * 64), "little") % pyref.Q
    assert int.from_bytes(got[0].tobytes(), "little") == v
    rng = np.random.default_rng(3)
    w = np.frombuffer(rng.bytes(64 * 16), np.uint8).reshape(16, 64)
    got = orc.scalar_from_wide(w)
    for i in range(16):
        assert int.from_bytes(got[i].tobytes(), "little") == int.from_bytes(w[i].tobytes(), "little") % pyref.Q


# ----------------------------------------------------------------------------- pairings (src/pairings.rs:826-970, src/tests/mod.rs:78-231)
def _gens(orc):
    gxy, ginf = orc.G1.to_affine(orc.G1.generator())
    hxy, hinf = orc.G2.to_affine(orc.G2.generator())
    return gxy, ginf, hxy, hinf


def test_gt_generator_kat(orc):
    gxy, ginf, hxy, hinf = _gens(orc)
    e = orc.pairing(gxy, ginf, hxy, hinf)
    assert eq(e, cat("pairings.rs::generator", range(12)))                          # src/pairings.rs:827-832
    assert eq(e, cat("tests/mod.rs::test_pairing_result_against_relic", range(12)))   # src/tests/mod.rs:114-231
    # the RELIC author's canonical (non-Montgomery) hex dump of the same value (src/tests/mod.rs:81-95):
    # RELIC's e(G1,G2) is the textbook pairing; the crate returns its cube, so compare f^... only
    # on the limbs the reference itself asserts (above); the canonical dump is checked to be a valid Fp12.
    for h in KAT["tests/mod.rs::relic_canonical_hex"]:
        assert int(h, 16) < pyref.P
    # pairing == multi_miller_loop(prepared) . final_exponentiation  (src/tests/mod.rs:105-112)
    ml = orc.multi_miller_loop(gxy, ginf, hxy, hinf)
    assert eq(orc.final_exponentiation(ml), e)
    assert eq(orc.final_exponentiation(orc.miller_loop(gxy, ginf, hxy, hinf)), e)


def test_pairing_bilinearity_and_identities(orc):
    gxy, ginf, hxy, hinf = _gens(orc)
    a, b = 0x1234567 * 0x9abcdef % pyref.Q, (pyref.Q - 77)
    sb = lambda s: np.frombuffer(int(s).to_bytes(32, "little"), np.uint8)
    ag = orc.G1.to_affine(orc.G1.mul(orc.G1.generator(), sb(a)))
    bh = orc.G2.to_affine(orc.G2.mul(orc.G2.generator(), sb(b)))
    abg = orc.G1.to_affine(orc.G1.mul(orc.G1.generator(), sb(a * b % pyref.Q)))
    lhs = orc.pairing(ag[0], ag[1], bh[0], bh[1])
    rhs = orc.pairing(abg[0], abg[1], hxy, hinf)
    assert eq(lhs, rhs)                                                             # src/pairings.rs:835-855
    one = np.zeros(72, np.uint64)
    one[:6] = orc.R_LIMBS
    assert not eq(lhs, one)
    # unitarity (src/pairings.rs:858-867): e(P,Q) * e(-P,Q) == 1 and e(-P,Q) == conj(e(P,Q))
    ng = orc.G1.to_affine(np.concatenate([orc.G1.generator()[0, :6], orc.tower(1, "neg", orc.G1.generator()[0, 6:12])[0],
                                          orc.G1.generator()[0, 12:]]))
    p = orc.pairing(gxy, ginf, hxy, hinf)
    q = orc.pairing(ng[0], ng[1], hxy, hinf)
    assert eq(orc.tower(12, "mul", p, q), one) and eq(q, orc.tower(12, "conjugate", p))
    # identity on either side -> Gt identity (src/pairings.rs:944-970)
    i1, i1f = orc.G1.affine_identity()
    i2, i2f = orc.G2.affine_identity()
    assert eq(orc.pairing(i1, i1f, hxy, hinf), one)
    assert eq(orc.pairing(gxy, ginf, i2, i2f), one)
    assert eq(orc.miller_loop(i1, i1f, hxy, hinf), one)
    # multi_miller_loop with identity terms == product of the pairings (src/pairings.rs:869-921)
    pxy = np.concatenate([ag[0], i1, gxy, abg[0]])
    pinf = np.concatenate([ag[1], i1f, ginf, abg[1]])
    qxy = np.concatenate([bh[0], hxy, i2, hxy])
    qinf = np.concatenate([bh[1], hinf, i2f, hinf])
    prod = orc.final_exponentiation(orc.multi_miller_loop(pxy, pinf, qxy, qinf))
    each = orc.pairing(pxy, pinf, qxy, qinf)
    acc = one
    for i in range(4):
        acc = orc.tower(12, "mul", acc, each[i])
    assert eq(prod, acc)
    assert orc.g2_prepare(hxy).shape == (68, 36)                                    # src/pairings.rs:539


def test_msm_pippenger_matches_naive(orc):
    """CPU Pippenger (used for full-size GPU checks) == the reference-API MSM sum_i p_i*s_i."""
    rng = np.random.default_rng(11)
    for G in (orc.G1, orc.G2):
        n = 37
        t = np.frombuffer(rng.bytes(32 * n), np.uint8).reshape(n, 32).copy()
        t[:, 31] &= 0x3f
        pts = G.mul(np.repeat(G.generator(), n, 0), t, threads=4)
        xy, inf = G.batch_normalize(pts)
        s = np.frombuffer(rng.bytes(32 * n), np.uint8).reshape(n, 32).copy()
        s[:, 31] &= 0x3f
        s[0] = 0
        inf[3] = 1
        a = G.to_affine(G.msm_naive(xy, inf, s, threads=4))
        for c in (4, 7):
            b = G.to_affine(G.msm_pippenger(xy, inf, s, c=c, threads=4))
            assert eq(a[0], b[0]) and a[1][0] == b[1][0]


# ----------------------------------------------------------------------------- subgroup checks (src/g1.rs:1598-1623, src/g2.rs:1862-1907)
def test_is_torsion_free_kats(orc):
    k = "g1.rs::test_is_torsion_free"               # a curve point outside the q-order subgroup
    a = cat(k, [0, 1])
    assert orc.G1.checks(a)[0] == 1                  # on the curve, NOT torsion free
    gxy, ginf = orc.G1.to_affine(orc.G1.generator())
    ixy, iinf = orc.G1.affine_identity()
    assert orc.G1.checks(gxy, ginf)[0] == 3 and orc.G1.checks(ixy, iinf)[0] == 3
    k = "g2.rs::test_is_torsion_free"
    a = cat(k, [0, 1, 2, 3])
    assert orc.G2.checks(a)[0] == 1
    hxy, hinf = orc.G2.to_affine(orc.G2.generator())
    ixy, iinf = orc.G2.affine_identity()
    assert orc.G2.checks(hxy, hinf)[0] == 3 and orc.G2.checks(ixy, iinf)[0] == 3
    # every golden [i]G is in the subgroup
    rng = np.random.default_rng(2)
    t = rng.integers(0, 256, (6, 32), dtype=np.uint8)
    t[:, 31] &= 0x3f
    for G in (orc.G1, orc.G2):
        xy, inf = G.batch_normalize(G.mul(np.repeat(G.generator(), 6, 0), t))
        assert (G.checks(xy, inf) == 3).all()
