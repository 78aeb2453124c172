"""GPU test (-m gpu): the EXPERIMENTAL second build of the MSM unit (capi_msm_lazy3.cu, row-alternated lazy Fp2 multiply;
entry points b200x_lazy3_*, not part of the public header) gives the same G2 MSM result as the default build and as the
oracle.  CPU-validated arithmetic (tests/test_device_source_cpu.py variant "lazy3"); first hardware run pending ->
non-strict xfail.  Nothing in the product calls these entry points by default."""
import ctypes as C

import numpy as np
import pytest

from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900),
              pytest.mark.xfail(strict=False, reason="first hardware run pending (round-1 GPU budget exhausted)")]


def test_lazy3_g2_msm_agrees(orc):
    import bls12_381_b200
    eng = bls12_381_b200.Engine()
    try:
        f = eng.lib.b200x_lazy3_g2_msm
        f.argtypes = [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p]
        f.restype = C.c_int
        rng = np.random.default_rng(16100)
        for n in (1, 33, 3000):
            _, xy, inf = util.rand_points(orc, 2, rng, n)
            s = util.rand_scalars(rng, n)
            if n > 2:
                inf[1] = 1
                s[2] = 0
            out = np.empty((1, 36), np.uint64)
            assert f(eng.h, xy.ctypes.data, inf.ctypes.data, s.ctypes.data, n, out.ctypes.data) == 0
            base = eng.msm(2, xy, inf, s)
            assert np.array_equal(orc.G2.to_affine(out)[0], orc.G2.to_affine(base)[0])
            if n <= 33:
                assert np.array_equal(orc.G2.to_affine(out)[0], orc.G2.to_affine(orc.G2.msm_naive(xy, inf, s, threads=8))[0])
    finally:
        eng.close()
