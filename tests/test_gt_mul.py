"""`&Gt * &Scalar` (src/pairings.rs:296-323): the oracle against an independent Python tower model and against
bilinearity; the device source (gt.cuh) on the CPU harness against the oracle; the GPU test of the same (first hardware
run pending, like the other rows written after round 1's GPU budget: non-strict xfail)."""
import ctypes as C

import numpy as np
import pytest

from tests import pyref, pyref_tower as T, util
from tests.emul import build as emul_build


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _inputs(orc, n, seed):
    rng = np.random.default_rng(seed)
    _, pxy, pinf = util.rand_points(orc, 1, rng, n)
    _, qxy, qinf = util.rand_points(orc, 2, rng, n)
    g = orc.pairing(pxy, pinf, qxy, qinf, threads=8)
    s = util.rand_scalars(rng, n)
    s[0] = 0
    if n > 1:
        s[1] = util.scalar_bytes(1)
    if n > 2:
        s[2] = util.scalar_bytes(pyref.Q - 1)
    return rng, (pxy, pinf, qxy, qinf), g, s


def test_oracle_gt_mul(orc):
    rng, (pxy, pinf, qxy, qinf), g, s = _inputs(orc, 6, 12100)
    out = orc.gt_mul(g, s, threads=6)
    one = np.zeros(72, np.uint64)
    one[:6] = pyref.to_mont(1)
    assert np.array_equal(out[0], one) and np.array_equal(out[1], g[1])
    assert np.array_equal(out[2], orc.tower(12, "conjugate", g[2:3])[0])          # g^(q-1) = g^-1 = conj(g) in Gt
    # independent: plain square-and-multiply in the Python tower
    for i in (3, 4):
        e = int.from_bytes(s[i].tobytes(), "little")
        assert np.array_equal(out[i], T.f12_to_limbs(T.f12_pow(T.f12_from_limbs(g[i]), e)))
    # bilinearity: e([a]P, Q) == e(P, Q)^a   (src/pairings.rs:834-867 tests the same identity)
    pr = orc.G1.from_affine(pxy, pinf)
    axy, ainf = orc.G1.batch_normalize(orc.G1.mul(pr, s, threads=6))
    assert np.array_equal(orc.pairing(axy, ainf, qxy, qinf, threads=6), out)
    # not only for Gt members: an arbitrary Fp12
    x = util.rand_fp(rng, 1, 12)
    e = int.from_bytes(s[3].tobytes(), "little")
    assert np.array_equal(orc.gt_mul(x, s[3:4])[0], T.f12_to_limbs(T.f12_pow(T.f12_from_limbs(x[0]), e)))


@pytest.mark.parametrize("variant", ["default", "kcall"])
def test_device_source_gt_mul_on_cpu(orc, variant):
    lib = C.CDLL(emul_build.build(variant))
    rng, _, g, s = _inputs(orc, 5, 12200)
    g = np.concatenate([g, util.rand_fp(rng, 1, 12)])
    s = np.concatenate([s, util.rand_scalars(rng, 1)])
    out = np.empty_like(g)
    lib.emul_gt_mul(_p(g), _p(np.ascontiguousarray(s).view(np.uint32)), C.c_size_t(g.shape[0]), _p(out), 6)
    assert np.array_equal(out, orc.gt_mul(g, s, threads=6))


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_gpu_zz_gt_mul(orc):
    import bls12_381_b200
    eng = bls12_381_b200.Engine()
    try:
        rng, (pxy, pinf, qxy, qinf), g, s = _inputs(orc, 200, 12300)
        g = np.concatenate([g, util.rand_fp(rng, 3, 12)])
        s = np.concatenate([s, util.rand_scalars(rng, 3)])
        out = eng.gt_mul_batch(g, s)
        assert np.array_equal(out, orc.gt_mul(g, s, threads=8))
        # bilinearity through the GPU's own scalar multiplication and pairing
        pr = orc.G1.from_affine(pxy, pinf)
        axy, ainf = eng.batch_normalize(1, eng.mul_batch(1, pr, s[:200]))
        assert np.array_equal(eng.pairing_batch(axy, ainf, qxy, qinf), out[:200])
        assert eng.gt_mul_batch(g[:0], s[:0]).shape == (0, 72)
    finally:
        eng.close()
