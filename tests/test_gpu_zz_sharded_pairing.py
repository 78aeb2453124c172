"""GPU test (-m gpu) of bls12_381_b200.sharding.ShardedPairingProduct on one rank (the multi-rank exchange is covered by
the gloo tests in tests/test_sharding_cpu.py; the device entry points it composes are validated in
tests/test_gpu_parity.py).  Hardware-validated in round 2."""
import numpy as np
import pytest

from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def test_pairing_product_one_rank(orc):
    import torch
    import bls12_381_b200
    from bls12_381_b200.sharding import ShardedPairingProduct
    eng = bls12_381_b200.Engine()
    try:
        rng = np.random.default_rng(13100)
        n = 37
        _, pxy, pinf = util.rand_points(orc, 1, rng, n)
        _, qxy, qinf = util.rand_points(orc, 2, rng, n)
        pinf[3] = 1
        qinf[n - 1] = 1
        dev = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).cuda()
        out = torch.zeros((1, 72), dtype=torch.int64, device="cuda")
        parts = torch.zeros((1, 72), dtype=torch.int64, device="cuda")
        scratch = torch.zeros((n, 72), dtype=torch.int64, device="cuda")
        sp = ShardedPairingProduct(eng)
        dp, dpi, dq, dqi = dev(pxy), dev(pinf), dev(qxy), dev(qinf)
        torch.cuda.synchronize()              # the engine works on its own non-blocking stream
        sp.multi_miller_loop(dp, dpi, dq, dqi, n, out, parts, scratch)
        torch.cuda.synchronize()
        want = orc.multi_miller_loop(pxy, pinf, qxy, qinf)
        assert np.array_equal(out.cpu().numpy().view(np.uint64), want)
        assert np.array_equal(eng.multi_miller_loop(pxy, pinf, qxy, qinf), want)      # the single-call entry point agrees
        sp.multi_miller_loop(dp, dpi, dq, dqi, n, out, parts, scratch, final_exp=True)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint64), orc.final_exponentiation(want))
    finally:
        eng.close()
