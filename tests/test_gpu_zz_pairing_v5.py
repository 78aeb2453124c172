"""GPU parity test (-m gpu) of the EXPERIMENTAL pairing kernels (pairing_v5.cu / pairing_v6.cu: dual- / triple-stream Fp2 multiply, tuning
key pairing_variant = 5 / 6) against the oracle and against the default kernels.  CPU-validated (tests/test_device_source_cpu.py,
variants "kdual" / "ktriple"); first hardware run pending -> non-strict xfail.  The default (pairing_variant = 4) is untouched."""
import numpy as np
import pytest

from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900),
              pytest.mark.xfail(strict=False, reason="first hardware run pending (round-1 GPU budget exhausted)")]


@pytest.mark.parametrize("variant", [5, 6])
def test_pairing_variant_parity(orc, variant):
    import bls12_381_b200
    eng = bls12_381_b200.Engine()
    try:
        rng = np.random.default_rng(15100)
        n = 700
        _, pxy, pinf = util.rand_points(orc, 1, rng, n)
        _, qxy, qinf = util.rand_points(orc, 2, rng, n)
        pinf[5] = 1
        qinf[9] = 1
        base = eng.pairing_batch(pxy, pinf, qxy, qinf)
        ml4 = eng.miller_loop_batch(pxy, pinf, qxy, qinf)
        eng.set_tuning("pairing_variant", variant)
        got = eng.pairing_batch(pxy, pinf, qxy, qinf)
        ml5 = eng.miller_loop_batch(pxy, pinf, qxy, qinf)
        fe5 = eng.final_exponentiation_batch(ml5)
        eng.set_tuning("pairing_variant", 4)
        assert np.array_equal(got, base) and np.array_equal(ml5, ml4) and np.array_equal(fe5, base)
        assert np.array_equal(got[:64], orc.pairing(pxy[:64], pinf[:64], qxy[:64], qinf[:64], threads=8))
        with pytest.raises(bls12_381_b200.B200Error):
            eng.set_tuning("pairing_variant", 7)
    finally:
        eng.close()
