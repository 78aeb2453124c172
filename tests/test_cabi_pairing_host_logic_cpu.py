"""The pairing units of the C ABI (capi_pairing.cu + pairing_v4.cu: launch wrappers, variant dispatch by the
tuning key pairing_variant, chunked two-stream schedule, product kernel with __syncthreads + dynamic shared memory,
prepared Miller loop) compiled with g++ against the mock CUDA runtime, every launch on the fiber scheduler, and driven
through the real Engine methods — against the oracle.  CPU only; complements tests/test_gpu_parity.py (v4, validated on
hardware)."""
import ctypes as C

import numpy as np
import pytest

from tests import util
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


@pytest.fixture(scope="module")
def pairs(orc):
    rng = np.random.default_rng(17100)
    n = 7
    _, pxy, pinf = util.rand_points(orc, 1, rng, n)
    _, qxy, qinf = util.rand_points(orc, 2, rng, n)
    pinf[2] = 1
    qinf[5] = 1
    return pxy, pinf, qxy, qinf


@pytest.mark.parametrize("variant", [4, 7])
def test_variants_through_the_c_abi(eng, orc, pairs, variant):
    pxy, pinf, qxy, qinf = pairs
    eng.set_tuning("pairing_variant", variant)
    try:
        ml = eng.miller_loop_batch(pxy, pinf, qxy, qinf)
        want_ml = orc.miller_loop(pxy, pinf, qxy, qinf, threads=7)
        assert np.array_equal(ml, want_ml)
        assert np.array_equal(eng.final_exponentiation_batch(ml), orc.final_exponentiation(want_ml, threads=7))
        assert np.array_equal(eng.pairing_batch(pxy, pinf, qxy, qinf), orc.pairing(pxy, pinf, qxy, qinf, threads=7))
    finally:
        eng.set_tuning("pairing_variant", 0)


def test_product_and_prepared_paths(eng, orc, pairs):
    pxy, pinf, qxy, qinf = pairs
    want = orc.multi_miller_loop(pxy, pinf, qxy, qinf)
    assert np.array_equal(eng.multi_miller_loop(pxy, pinf, qxy, qinf), want)          # k_fp12_product: barriers + smem
    co = eng.g2_prepare(qxy, qinf)
    assert np.array_equal(co[0].reshape(68, 36), orc.g2_prepare(qxy[0:1], int(qinf[0])))
    assert np.array_equal(eng.multi_miller_loop_prepared(pxy, pinf, co, qinf), want)
    assert np.array_equal(eng.multi_miller_loop(pxy[:1], pinf[:1], qxy[:1], qinf[:1]),
                          orc.multi_miller_loop(pxy[:1], pinf[:1], qxy[:1], qinf[:1]))


def test_tuning_key_validation(eng):
    from bls12_381_b200 import B200Error
    for bad in (3, 5, 8, -1):
        with pytest.raises(B200Error):
            eng.set_tuning("pairing_variant", bad)
    with pytest.raises(B200Error):
        eng.set_tuning("no_such_key", 1)


def test_chunked_schedule(eng, orc):
    """pairing_dev's chunked two-stream path (taken above sm_count * 256 + 2048 pairs; the mock ctx has sm_count = 1):
    chunk boundaries, 64-alignment of the chunk size, the ragged last chunk"""
    rng = np.random.default_rng(17200)
    n = 2400
    _, pxy, pinf = util.rand_points(orc, 1, rng, 40)
    _, qxy, qinf = util.rand_points(orc, 2, rng, 40)
    idx = rng.integers(0, 40, n)
    jdx = rng.integers(0, 40, n)
    P, PI, Qx, QI = pxy[idx], pinf[idx].copy(), qxy[jdx], qinf[jdx].copy()
    PI[700] = 1
    QI[2399] = 1
    eng.set_tuning("pairing_chunks", 3)
    eng.set_tuning("pairing_variant", 4)        # the chunked schedule belongs to the one-thread-per-pairing kernels
    try:
        got = eng.pairing_batch(P, PI, Qx, QI)
    finally:
        eng.set_tuning("pairing_chunks", 4)
        eng.set_tuning("pairing_variant", 0)
    assert np.array_equal(got, orc.pairing(P, PI, Qx, QI, threads=8))


def test_six_lane_kernels_ragged_batches(eng, orc):
    """pairing_variant 7 (pairing_coop.cu): batch sizes that are not multiples of the 5 pairs a warp takes per turn, several
    warps per block, more warps than work, identities on either side; Miller value, final exponentiation and Gt limb-exact"""
    rng = np.random.default_rng(17300)
    _, pxy, pinf = util.rand_points(orc, 1, rng, 23)
    _, qxy, qinf = util.rand_points(orc, 2, rng, 23)
    pinf[4] = 1
    qinf[9] = 1
    pinf[22], qinf[22] = 1, 1
    for warps, n in ((1, 1), (2, 6), (3, 23), (12, 11)):
        eng.set_tuning("coop_warps", warps)
        try:
            a = (pxy[:n], pinf[:n], qxy[:n], qinf[:n])
            ml = orc.miller_loop(*a, threads=8)
            assert np.array_equal(eng.miller_loop_batch(*a), ml)
            assert np.array_equal(eng.final_exponentiation_batch(ml), orc.final_exponentiation(ml, threads=8))
            assert np.array_equal(eng.pairing_batch(*a), orc.pairing(*a, threads=8))
        finally:
            eng.set_tuning("coop_warps", 12)


def test_shared_squaring_product_mode(eng, orc):
    """multi_miller_loop with ONE squaring per bit for all terms (src/pairings.rs:554-603) on the six-lane kernels: a single
    product over n terms (chunked over groups + k_coop_product fold; the mock ctx has one SM, so 24 terms already take the
    chunked route) and batches of small products (Groth16 shape), identities skipped, with and without final exponentiation"""
    rng = np.random.default_rng(17500)
    n = 24
    _, pxy, pinf = util.rand_points(orc, 1, rng, n)
    _, qxy, qinf = util.rand_points(orc, 2, rng, n)
    pinf[2] = 1
    qinf[5] = 1
    pinf[7], qinf[7] = 1, 1
    one = np.zeros(72, np.uint64)
    one[:6] = orc.R_LIMBS
    eng.set_tuning("coop_warps", 2)
    try:
        for m in (0, 1, 2, 9, 24):
            got = eng.multi_miller_loop(pxy[:m], pinf[:m], qxy[:m], qinf[:m]).reshape(-1)
            want = orc.multi_miller_loop(pxy[:m], pinf[:m], qxy[:m], qinf[:m]).reshape(-1) if m else one
            assert np.array_equal(got, want), m
        for terms in (1, 3, 4):
            npr = n // terms
            a = tuple(x[:npr * terms] for x in (pxy, pinf, qxy, qinf))
            want = np.concatenate([orc.multi_miller_loop(*(x[i * terms:(i + 1) * terms] for x in a)).reshape(1, 72)
                                   for i in range(npr)])
            assert np.array_equal(eng.pairing_product_batch(*a, terms, final_exp=False), want)
            assert np.array_equal(eng.pairing_product_batch(*a, terms, final_exp=True), orc.final_exponentiation(want, threads=4))
    finally:
        eng.set_tuning("coop_warps", 12)


def test_six_lane_kernels_chunked_two_stream_schedule(eng, orc):
    """full pairings of a batch larger than two waves of groups (one SM x 1 warp x 5 pairs on the mock): coop_chunks chunks
    alternating between the two streams of the ctx, coefficients of all chunks in one arena block"""
    rng = np.random.default_rng(17301)
    _, pxy, pinf = util.rand_points(orc, 1, rng, 13)
    _, qxy, qinf = util.rand_points(orc, 2, rng, 13)
    pinf[3] = 1
    qinf[11] = 1
    eng.set_tuning("coop_warps", 1)
    eng.set_tuning("coop_chunks", 3)
    try:
        got = eng.pairing_batch(pxy, pinf, qxy, qinf)
    finally:
        eng.set_tuning("coop_warps", 12)
        eng.set_tuning("coop_chunks", 3)
    assert np.array_equal(got, orc.pairing(pxy, pinf, qxy, qinf, threads=8))


@pytest.mark.parametrize("prepare_max", [0, 5000])
def test_g2_prepare_both_kernels_on_mock(eng, orc, prepare_max):
    """G2Prepared coefficients from the one-thread-per-Q kernel and from the six-lanes-per-Q kernel (k_coop_g2_prepare: shared-
    memory board, per-group __syncwarp on the fiber scheduler), limb for limb against the oracle; identity = generator's coefficients"""
    rng = np.random.default_rng(17400)
    _, qxy, qinf = util.rand_points(orc, 2, rng, 7)
    qinf[3] = 1
    eng.set_tuning("coop_prepare_max", prepare_max)
    try:
        co = eng.g2_prepare(qxy, qinf)
    finally:
        eng.set_tuning("coop_prepare_max", 5000)
    for i in range(7):
        assert np.array_equal(co[i], orc.g2_prepare(qxy[i], qinf[i])), i
