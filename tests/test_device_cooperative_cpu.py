"""Warp-cooperative device code on the CPU harness: tests/emul/fiber_warp.h runs every CUDA thread of a block as a
user-level fiber and implements __shfl_sync / __ballot_sync / __syncwarp / __syncthreads as counting barriers, so the
lane-parallel group operations of curve_warp.cuh (the config-1 warp kernel and the MSM Horner tail use them) are checked
against the oracle without a GPU.  A self-test pins the scheduler itself (shared memory + barriers, shuffles, ballot,
an early-exiting warp)."""
import ctypes as C

import numpy as np
import pytest

from tests import pyref, util
from tests.emul import build as emul_build


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def lib():
    return C.CDLL(emul_build.build("default"))


@pytest.mark.parametrize("block", [32, 64, 96, 256])
def test_fiber_scheduler_selftest(lib, block):
    rng = np.random.default_rng(14100 + block)
    x = rng.integers(-1000, 1000, block).astype(np.int32)
    bs = np.zeros(1, np.int32)
    ws = np.zeros(block // 32, np.int32)
    bl = np.zeros(block // 32, np.uint32)
    lib.emul_fiber_selftest(block, _p(x), block, _p(bs), _p(ws), _p(bl))
    live = np.ones(block, bool)
    if block > 64:
        live[32:64] = False                     # warp 1 exits before any collective
    assert bs[0] == x[live].sum()
    for w in range(block // 32):
        if not live[32 * w]:
            continue
        assert ws[w] == x[32 * w:32 * w + 32].sum()
        assert bl[w] == sum(int(v & 1) << l for l, v in enumerate(x[32 * w:32 * w + 32]))


@pytest.mark.parametrize("k", [1, 2])
def test_warp_cooperative_scalar_mul(lib, orc, k):
    """capi_basic.cu::k_mul_batch_warp's schedule (warp_double / warp_add of curve_warp.cuh) == the reference's
    double-and-add, limb-exact on raw (x, y, z)"""
    G = orc.G1 if k == 1 else orc.G2
    rng = np.random.default_rng(14200 + k)
    n = 5 if k == 1 else 3                      # 5 items: two 128-thread blocks, the second one partly empty
    pr, _, _ = util.rand_points(orc, k, rng, n)
    pr = util.randomize_z(orc, k, rng, pr)
    s = util.rand_scalars(rng, n)
    s[0] = util.scalar_bytes(pyref.Q - 1)
    if n > 3:
        s[3] = 0
        pr[4] = G.identity(1)
    out = np.empty_like(pr)
    getattr(lib, "emul_g%d_mul_warp" % k)(_p(pr), _p(np.ascontiguousarray(s).view(np.uint32)), _p(out), C.c_size_t(n))
    assert np.array_equal(out, G.mul(pr, s, threads=4))
