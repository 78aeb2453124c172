"""The bodies of the hardware-validated parity tests (tests/test_gpu_parity.py) replayed on the CPU: the real C ABI units
compiled with g++ against the mock CUDA runtime (tests/emul/: same device source, bit-exact PTX carry models, fiber
scheduler for cooperative kernels) behind the real Engine class.  MSM has its own file (tests/test_msm_on_mock_cpu.py); the golden-file serialization test (1000 warp-cooperative
scalar multiplications to build its inputs) is too slow on the fiber scheduler — tests/test_cabi_host_logic_cpu.py has a
smaller one — and the G2Prepared test needs torch CUDA tensors.  Purpose: a regression net for edits made without a GPU —
these tests were green on a B200 in round 1; if one turns red here, the edit changed behaviour."""
import ctypes as C

import numpy as np
import pytest

import tests.test_gpu_parity as GP
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
            rc = lib.b200_ctx_create(-1, C.byref(h))       # the REAL b200_ctx_create of capi_basic.cu, on the mock runtime
            assert rc == 0
            self.h = h

    e = MockEngine()
    yield e
    e.close()


@pytest.mark.parametrize("op", ["mul", "add", "sub", "square", "neg", "invert"])
def test_fp_ops(eng, orc, op):
    GP.test_fp_ops(eng, orc, op)


def test_fp_kats(eng):
    GP.test_fp_kat_on_gpu(eng)


@pytest.mark.parametrize("level,ops,n", [
    (2, ["mul", "add", "sub", "square", "neg", "invert", "frobenius", "conjugate", "mul_by_nonresidue"], 60),
    (6, ["mul", "add", "sub", "square", "neg", "invert", "frobenius", "mul_by_nonresidue"], 30),
    (12, ["mul", "square", "invert", "frobenius", "conjugate", "cyclotomic_square"], 12)])
def test_tower_ops(eng, orc, level, ops, n):
    GP.test_tower_ops(eng, orc, level, ops, n)


@pytest.mark.parametrize("k", [1, 2])
def test_group_ops(eng, orc, k):
    GP.test_group_ops(eng, orc, k)


@pytest.mark.parametrize("k,n", [(1, 6), (2, 4)])
def test_mul_batch_config1(eng, orc, k, n):
    GP.test_mul_batch_config1(eng, orc, k, n)        # small n: the warp-per-item kernel (shuffles on the fiber scheduler)


def test_mul_batch_three_items_per_warp(eng, orc):
    """the six-lane group kernel with several items per warp (config 1 runs two per warp on the GPU): 7 items, 3 per warp"""
    import tests.util as util
    rng = np.random.default_rng(4711)
    pr, _, _ = util.rand_points(orc, 1, rng, 7)
    p = util.randomize_z(orc, 1, rng, pr)
    s = util.rand_scalars(rng, 7)
    s[0] = 0
    p[4] = orc.G1.identity()
    eng.set_tuning("mul_groups", 3)
    try:
        got = eng.mul_batch(1, p, s)
    finally:
        eng.set_tuning("mul_groups", 0)
    assert np.array_equal(got, orc.G1.mul(p, s, threads=4))


def test_pairing_and_gt_kat(eng, orc):
    GP.test_gt_generator_kat_on_gpu(eng, orc)


@pytest.mark.parametrize("k", [1, 2])
def test_deserialization_rejects(eng, orc, k):
    GP.test_deserialization_rejects(eng, orc, k)


@pytest.mark.parametrize("k", [1, 2])
def test_subgroup_checks(eng, orc, k):
    GP.test_subgroup_checks(eng, orc, k)


def test_fp_invert_fast(eng, orc):
    GP.test_fp_invert_fast(eng, orc)


def test_imad_peak_and_timing_records(eng):
    v, ms = eng.imad_peak(iters=2)
    assert v > 0 and ms >= 0
    eng.set_timing(True)
    eng.tower(1, "mul", np.zeros((4, 6), np.uint64), np.zeros((4, 6), np.uint64))
    rec = eng.get_timing()
    eng.set_timing(False)
    assert len(rec) == 1 and rec[0][0].startswith("k_tower_op")
