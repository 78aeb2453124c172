"""The DEVICE-SIDE source (bls12_381_b200/csrc/*.cuh) compiled for the host and checked against the oracle.

tests/emul/cuda_host_shim.h turns the CUDA qualifiers into no-ops and swaps the PTX carry-chain primitives of fp.cuh for
bit-exact C models, so the limb algorithms (even/odd-accumulator Montgomery product, wide product + REDC, dedicated
squaring, binary-GCD inverse), the tower, the curve formulas, the Miller loop and the final exponentiation that the
GPU executes are exercised by the CPU suite too.  The PTX->SASS path, launch geometry and warp cooperation are NOT
covered here — that is what the -m gpu tests are for.  Both Fp2-multiply variants the build uses are compiled.
"""
import ctypes as C

import numpy as np
import pytest

from tests import pyref, util
from tests.emul import build as emul_build


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Emul:
    def __init__(self, variant):
        self.lib = C.CDLL(emul_build.build(variant))
        self.lib.emul_tower_op.restype = C.c_int

    def tower(self, level, op, a, b=None):
        a = np.ascontiguousarray(a, np.uint64).reshape(-1, 6 * level)
        b = None if b is None else np.ascontiguousarray(b, np.uint64).reshape(-1, 6 * level)
        out = np.empty_like(a)
        rc = self.lib.emul_tower_op(level, op, _p(a), _p(b), _p(out), C.c_size_t(a.shape[0]))
        assert rc == 0
        return out

    def group(self, k, name, *arrs, n=None, extra=()):
        f = getattr(self.lib, "emul_g%d_%s" % (k, name))
        f(*[_p(a) if isinstance(a, np.ndarray) or a is None else a for a in arrs], *extra)


@pytest.fixture(scope="module", params=["default", "kcall", "lazy3"])
def em(request):
    return Emul(request.param)


OPS = dict(mul=0, add=1, sub=2, square=3, neg=4, invert=5, frobenius=6, conjugate=7, mul_by_nonresidue=8,
           cyclotomic_square=9, invert_fast=10)


def test_fp_limb_algorithms(em, orc):
    rng = np.random.default_rng(4100)
    a = np.concatenate([util.rand_fp(rng, 200), util.edge_fp(), util.edge_fp()[::-1]])
    b = np.concatenate([util.rand_fp(rng, 200), util.edge_fp(), util.edge_fp()])
    want_mul, want_sq = orc.tower(1, "mul", a, b), orc.tower(1, "square", a)
    for code in (0, 100, 101, 104, 105):      # interleaved product, its called copy, wide product + REDC, dual-stream
        assert np.array_equal(em.tower(1, code, a, b), want_mul)
    for code in (3, 102):           # dedicated squaring (78 + 156 IMAD), mul(a, a)
        assert np.array_equal(em.tower(1, code, a), want_sq)
    for name in ("add", "sub"):
        assert np.array_equal(em.tower(1, OPS[name], a, b), orc.tower(1, name, a, b))
    assert np.array_equal(em.tower(1, OPS["neg"], a), orc.tower(1, "neg", a))
    inv = orc.tower(1, "invert", a[:40])
    assert np.array_equal(em.tower(1, OPS["invert"], a[:40]), inv)
    assert np.array_equal(em.tower(1, OPS["invert_fast"], a), orc.tower(1, "invert", a))  # Kaliski binary GCD
    # independent of the oracle: Python big integers
    for i in range(0, 200, 17):
        x, y = pyref.from_mont(a[i]), pyref.from_mont(b[i])
        assert np.array_equal(em.tower(1, 0, a[i:i + 1], b[i:i + 1])[0], pyref.to_mont(x * y % pyref.P))


def test_tower(em, orc):
    rng = np.random.default_rng(4200)
    for level, n in ((2, 64), (6, 24), (12, 10)):
        a, b = util.rand_fp(rng, n, level), util.rand_fp(rng, n, level)
        a[0] = 0
        b[1] = 0
        ops2 = ("mul",) if level == 12 else ("mul", "add", "sub")
        for name in ops2:
            assert np.array_equal(em.tower(level, OPS[name], a, b), orc.tower(level, name, a, b)), (level, name)
        ops1 = {2: ("square", "neg", "invert", "conjugate", "mul_by_nonresidue"),
                6: ("square", "neg", "invert", "frobenius", "mul_by_nonresidue"),
                12: ("square", "invert", "frobenius", "conjugate")}[level]
        for name in ops1:
            assert np.array_equal(em.tower(level, OPS[name], a), orc.tower(level, name, a)), (level, name)
    a2, b2 = util.rand_fp(rng, 32, 2), util.rand_fp(rng, 32, 2)
    assert np.array_equal(em.tower(2, 100, a2, b2), orc.tower(2, "mul", a2, b2))   # inline Karatsuba
    assert np.array_equal(em.tower(2, 101, a2), orc.tower(2, "square", a2))
    assert np.array_equal(em.tower(2, 102, a2), orc.tower(2, "invert", a2))


@pytest.mark.parametrize("k", [1, 2])
def test_group_formulas(em, orc, k):
    G = orc.G1 if k == 1 else orc.G2
    rng = np.random.default_rng(4300 + k)
    n = 24
    pr, xy, inf = util.rand_points(orc, k, rng, n)
    qr, qxy, qinf = util.rand_points(orc, k, rng, n)
    pr = util.randomize_z(orc, k, rng, pr)
    # exceptional cases: identity operands, P + P, P + (-P)
    ident = G.identity(1)
    pr[0], qr[1] = ident, ident
    qr[2] = pr[2]
    qr[3] = pr[3]
    w = 6 * k
    qr[3, w:2 * w] = orc.tower(1 if k == 1 else 2, "neg", pr[3, w:2 * w])
    qinf[4] = 1
    out = np.empty_like(pr)
    em.group(k, "double", pr, out, extra=(C.c_size_t(n),))
    assert np.array_equal(out, G.double(pr))
    em.group(k, "add", pr, qr, out, extra=(C.c_size_t(n),))
    assert np.array_equal(out, G.add(pr, qr))
    em.group(k, "add_mixed", pr, qxy, qinf, out, extra=(C.c_size_t(n),))
    assert np.array_equal(out, G.add_mixed(pr, qxy, qinf))
    axy, ainf = np.empty_like(xy), np.empty_like(inf)
    em.group(k, "to_affine", pr, axy, ainf, extra=(C.c_size_t(n),))
    wxy, winf = G.to_affine(pr)
    assert np.array_equal(axy, wxy) and np.array_equal(ainf, winf)
    # the reference's double-and-add, limb-exact on raw (x, y, z)   (config 1 of BASELINE.json)
    s = util.rand_scalars(rng, 6)
    s[0] = 0
    s[1] = util.scalar_bytes(pyref.Q - 1)
    m = 6
    outm = np.empty_like(pr[:m])
    s32 = np.ascontiguousarray(s).view(np.uint32)
    em.group(k, "mul", np.ascontiguousarray(pr[5:5 + m]), s32, outm, extra=(C.c_size_t(m), 4))
    assert np.array_equal(outm, G.mul(pr[5:5 + m], s, threads=4))
    # XYZZ bucket accumulator == sum of the points (compared in affine; the representation differs by design)
    dup = np.concatenate([xy, xy[:3], qxy])
    dinf = np.concatenate([inf, inf[:3], qinf])
    neg = xy[5:6].copy()
    neg[0, w:] = orc.tower(1 if k == 1 else 2, "neg", xy[5:6, w:])
    dup = np.concatenate([dup, neg, xy[5:6]])       # ... + (-P5) + P5: passes through the identity-after-add case
    dinf = np.concatenate([dinf, [0, 0]]).astype(np.uint8)
    acc = np.empty((1, 18 * k), np.uint64)
    em.group(k, "xyzz_sum", np.ascontiguousarray(dup), np.ascontiguousarray(dinf), C.c_size_t(dup.shape[0]), acc)
    ones = np.zeros((dup.shape[0], 32), np.uint8)
    ones[:, 0] = 1
    want = G.msm_naive(dup, dinf, ones, threads=4)
    assert np.array_equal(G.to_affine(acc)[0], G.to_affine(want)[0])


def test_pairing_source(em, orc):
    rng = np.random.default_rng(4400)
    n = 6
    _, pxy, pinf = util.rand_points(orc, 1, rng, n)
    _, qxy, qinf = util.rand_points(orc, 2, rng, n)
    pinf[1] = 1
    qinf[2] = 1
    ml = np.empty((n, 72), np.uint64)
    em.lib.emul_miller_loop(_p(pxy), _p(pinf), _p(qxy), _p(qinf), C.c_size_t(n), _p(ml), 6)
    want = orc.miller_loop(pxy, pinf, qxy, qinf, threads=6)
    assert np.array_equal(ml, want)
    fe = np.empty_like(ml)
    em.lib.emul_final_exponentiation(_p(ml), C.c_size_t(n), _p(fe), 6)
    assert np.array_equal(fe, orc.final_exponentiation(want, threads=6))
    assert np.array_equal(fe, orc.pairing(pxy, pinf, qxy, qinf, threads=6))
    # e(G1, G2) == the Gt generator constant of src/pairings.rs:359-475 (golden KAT)
    import json, os
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))
    gt = np.concatenate([np.array([int(x, 16) for x in g], dtype=np.uint64) for g in kat["pairings.rs::generator"]])
    g1xy, _ = orc.G1.to_affine(orc.G1.generator())
    g2xy, _ = orc.G2.to_affine(orc.G2.generator())
    one = np.empty((1, 72), np.uint64)
    em.lib.emul_miller_loop(_p(g1xy), None, _p(g2xy), None, C.c_size_t(1), _p(one), 1)
    em.lib.emul_final_exponentiation(_p(one), C.c_size_t(1), _p(one), 1)
    assert np.array_equal(one, orc.pairing(g1xy, None, g2xy, None))
    assert np.array_equal(one.reshape(-1), gt)
    # G2Prepared coefficients and the prepared Miller loop
    co = np.empty((68, 36), np.uint64)
    em.lib.emul_g2_prepare(_p(qxy[0:1]), 0, _p(co))
    assert np.array_equal(co, orc.g2_prepare(qxy[0:1], 0))
    mp = np.empty((1, 72), np.uint64)
    em.lib.emul_miller_loop_prepared(_p(pxy[0:1]), 0, _p(co), 0, _p(mp))
    assert np.array_equal(mp, want[0:1])


def test_glv_decompose_source(em):
    rng = np.random.default_rng(4500)
    lam = 0xac45a4010001a40200000000ffffffff
    vals = [0, 1, pyref.Q - 1, lam, lam - 1, lam + 1, (pyref.Q + 1) // 2, (pyref.Q - 1) // 2] + \
           [int.from_bytes(rng.bytes(40), "little") % pyref.Q for _ in range(200)]
    s = np.stack([np.frombuffer(v.to_bytes(32, "little"), np.uint32) for v in vals])
    out = np.empty((len(vals), 10), np.uint32)
    em.lib.emul_glv_decompose(_p(np.ascontiguousarray(s)), C.c_size_t(len(vals)), _p(out))
    for v, o in zip(vals, out):
        k1 = sum(int(x) << (32 * i) for i, x in enumerate(o[:4]))
        k2 = sum(int(x) << (32 * i) for i, x in enumerate(o[4:8]))
        assert k1 < (1 << 127) and k2 < (1 << 127)
        k1 = -k1 if o[8] else k1
        k2 = -k2 if o[9] else k2
        assert (k1 + k2 * lam - v) % pyref.Q == 0
