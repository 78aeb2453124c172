"""GPU tests (-m gpu) of the multi-GPU layer inside the library (csrc/capi_multi.cu): the single-process form
(b200_multi_*: one ctx + NCCL communicator + host thread per device) on however many GPUs the box has (1 works too: the
same code path without the collective).  The one-process-per-GPU form (b200_ctx_comm_init + b200_g1_msm_sharded_dev) is
exercised by bench.py under torchrun and by tools/multi_gpu_check.py."""
import numpy as np
import pytest

from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.mark.parametrize("mode", ["points", "window"])
@pytest.mark.parametrize("k", [1, 2])
def test_multi_engine_msm(orc, k, mode):
    import torch
    import bls12_381_b200
    G = orc.G1 if k == 1 else orc.G2
    ngpu = min(torch.cuda.device_count(), 8)
    m = bls12_381_b200.MultiEngine(ngpu, mode=mode)
    try:
        assert m.gpus == ngpu
        rng = np.random.default_rng(19000 + k)
        for n in (1, 37, 3000):
            _, xy, inf = util.rand_points(orc, k, rng, n)
            s = util.rand_scalars(rng, n)
            if n > 4:
                inf[3] = 1
                s[4] = 0
            got = G.to_affine(m.msm(k, xy, inf, s))
            want = G.to_affine(G.msm_naive(xy, inf, s, threads=8) if n <= 37 else G.msm_pippenger(xy, inf, s, c=8, threads=8))
            assert np.array_equal(got[0], want[0]) and got[1][0] == want[1][0], (n, ngpu)
        # n = 0 -> identity
        z = m.msm(k, np.zeros((0, 12 * k), np.uint64), None, np.zeros((0, 32), np.uint8))
        assert G.to_affine(z)[1][0] == 1
    finally:
        m.close()


def test_comm_api_single_rank(orc):
    """b200_comm_unique_id / b200_ctx_comm_init with world = 1 and the sharded entry point on one rank"""
    import torch
    import bls12_381_b200
    eng = bls12_381_b200.Engine()
    try:
        uid = eng.comm_unique_id()
        assert len(uid) == 128
        eng.comm_init(uid, 0, 1)
        assert eng.comm_world == 1
        rng = np.random.default_rng(19100)
        n = 500
        _, xy, inf = util.rand_points(orc, 1, rng, n)
        s = util.rand_scalars(rng, n)
        dev = torch.device("cuda", eng.device)
        t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(dev)
        out = torch.empty((1, 18), dtype=torch.int64, device=dev)
        for mode in ("points", "window"):
            eng.msm_sharded_dev(1, t(xy), t(inf), t(s), n, out, mode=mode)
            got = orc.G1.to_affine(out.cpu().numpy().view(np.uint64))
            want = orc.G1.to_affine(orc.G1.msm_pippenger(xy, inf, s, c=8, threads=8))
            assert np.array_equal(got[0], want[0])
        eng.comm_destroy()
    finally:
        eng.close()
