"""The MSM unit (capi_msm.cu: digit extraction, counting sort, bucket scheduling, bucket accumulation with its prefetch
pipeline, giant-bucket split, bucket reduction, warp-cooperative Horner, GLV, batched-affine levels, window sharding) and
its Fp2 multiply (row-alternated lazy reduction, B200_FP2_LAZY3) on the CPU: compiled with g++ against the mock CUDA runtime, every
kernel on the fiber scheduler (match.any, shuffles, atomics, shared memory, barriers), driven through the real C ABI and
the real Engine class — against the oracle.  Sizes are tiny (one emulated MSM costs ~2 s, mostly the 256-doubling Horner
chain on the fiber scheduler); the hardware-validated tests at real sizes are tests/test_gpu_parity.py and
tests/test_gpu_fullsize.py.  Purpose: a regression net for MSM edits made without a GPU."""
import ctypes as C

import numpy as np
import pytest

import tests.test_gpu_parity as GP
from tests import pyref, util
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
            assert lib.b200_ctx_create(-1, C.byref(h)) == 0
            self.h = h

    e = MockEngine()
    yield e
    e.close()


@pytest.fixture(scope="module")
def data(orc):
    rng = np.random.default_rng(18100)
    out = {}
    for k in (1, 2):
        _, xy, inf = util.rand_points(orc, k, rng, 64)
        out[k] = (xy, inf, util.rand_scalars(rng, 64))
    return out


def _edge_set(orc, k, xy, inf, s, n):
    xy2, inf2, s2 = xy[:n].copy(), inf[:n].copy(), s[:n].copy()
    w = 6 * k
    s2[0] = 0
    s2[1] = util.scalar_bytes(pyref.Q - 1)
    s2[2] = util.scalar_bytes(1)
    inf2[3] = 1
    xy2[5], s2[5] = xy2[4], s2[4]                    # P + P in a bucket
    xy2[7] = xy2[6]
    xy2[7, w:] = orc.tower(k, "neg", xy2[6, w:])
    s2[7] = s2[6]                                    # P + (-P) in a bucket
    s2[8:n] = s2[8]                                  # one crowded bucket per window
    return xy2, inf2, s2


def test_g1_msm_sizes_and_edges(eng, orc, data):
    xy, inf, s = data[1]
    G = orc.G1
    assert G.to_affine(eng.msm(1, xy[:0], None, s[:0]))[1][0] == 1                       # n = 0 -> identity
    GP._msm_case(eng, orc, 1, xy[:1], inf[:1], s[:1], cs=(0, 7))
    GP._msm_case(eng, orc, 1, xy[:37], inf[:37], s[:37], cs=(0,))
    GP._msm_case(eng, orc, 1, *_edge_set(orc, 1, xy, inf, s, 64), cs=(0, 8))
    both = np.concatenate([xy[:4], xy[:4]])
    both[4:, 6:] = orc.tower(1, "neg", xy[:4, 6:])
    assert G.to_affine(eng.msm(1, both, None, np.concatenate([s[:4], s[:4]])))[1][0] == 1   # everything cancels


@pytest.mark.parametrize("mode", [4, 2, 3])
def test_g2_msm_and_bucket_kernel_variants(eng, orc, data, mode):
    xy, inf, s = data[2]
    eng.set_tuning("g2_acc_blocks", mode)
    try:
        GP._msm_case(eng, orc, 2, *_edge_set(orc, 2, xy, inf, s, 24), cs=(6,))
        if mode == 4:
            GP._msm_case(eng, orc, 2, xy[:5], inf[:5], s[:5], cs=(0,))
    finally:
        eng.set_tuning("g2_acc_blocks", 4)


def test_glv_prefetch_and_affine_levels(eng, orc, data):
    xy, inf, s = data[1]
    case = _edge_set(orc, 1, xy, inf, s, 40)
    for key, val, back in (("g1_glv", 1, 0), ("g1_prefetch", 0, 1), ("msm_affine_levels", 2, 0)):
        eng.set_tuning(key, val)
        try:
            GP._msm_case(eng, orc, 1, *case, cs=(5,))
        finally:
            eng.set_tuning(key, back)


def test_window_sharding_and_sum(eng, orc, data):
    """b200_g1_msm_shard_dev for both shards of a 2-way split + b200_g1_sum_dev == the full MSM ("device" pointers are host
    pointers under the mock runtime)"""
    xy, inf, s = data[1]
    n = 30
    parts = np.zeros((2, 18), np.uint64)
    for shard in (0, 1):
        rc = eng.lib.b200_g1_msm_shard_dev(eng.h, xy.ctypes.data, inf.ctypes.data, s.ctypes.data, n, shard, 2,
                                           parts[shard:shard + 1].ctypes.data)
        assert rc == 0
    out = np.zeros((1, 18), np.uint64)
    assert eng.lib.b200_g1_sum_dev(eng.h, parts.ctypes.data, 2, out.ctypes.data) == 0
    want = orc.G1.msm_naive(xy[:n], inf[:n], s[:n], threads=4)
    assert np.array_equal(orc.G1.to_affine(out)[0], orc.G1.to_affine(want)[0])


def test_non_canonical_scalar_is_rejected(eng, orc, data):
    """the ABI takes Scalar::to_bytes() (canonical, < q): a raw 32-byte string >= q makes the call fail with B200_EINVAL
    instead of returning a wrong point (the signed-window recoding drops the carry out of the top window)"""
    from bls12_381_b200 import B200Error
    xy, inf, s = data[1]
    q = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    for bad in (q, q + 5, (1 << 256) - 1):
        s2 = s[:9].copy()
        s2[4] = np.frombuffer(bad.to_bytes(32, "little"), np.uint8)
        with pytest.raises(B200Error):
            eng.msm(1, xy[:9], inf[:9], s2)
    s2 = s[:9].copy()
    s2[4] = np.frombuffer((q - 1).to_bytes(32, "little"), np.uint8)          # the largest canonical scalar is fine
    GP._msm_case(eng, orc, 1, xy[:9], inf[:9], s2, cs=(0,))


def test_cooperative_bucket_reduction(eng, orc, data):
    """msm_reduce = 1 (k_msm_reduce_coop / k_msm_fold_coop: six lanes per bucket chunk, fixed-length offset multiply, folds),
    the default on hardware, on the fiber scheduler: one G1 case with a window wide enough for a fold level (the G2
    instantiation of the same templates runs in the -m gpu suite)"""
    eng.set_tuning("msm_reduce", 2)
    try:
        for k in (1,):
            xy, inf, s = data[k]
            GP._msm_case(eng, orc, k, xy[:20], inf[:20], s[:20], cs=(9,))        # 2^8 buckets: 16 chunks -> one fold level
    finally:
        eng.set_tuning("msm_reduce", 0)

