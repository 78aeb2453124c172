"""GPU parity at BASELINE.json's FULL sizes (-m gpu).  Inputs are generated on the GPU with the config-1 kernel
(itself limb-exact against the oracle at small sizes, tests/test_gpu_parity.py); the results are checked
(a) directly against the CPU oracle's Pippenger (cross-validated against the reference-API MSM in
tests/test_oracle_golden.py) and (b) through size-independent properties: additivity over a split of the point
range, additivity in the scalars (linearity), window-shard partials summing to the whole, bilinearity samples."""
import numpy as np
import pytest
import torch

from tests import pyref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import bls12_381_b200
    e = bls12_381_b200.Engine()
    yield e
    e.close()


def _points(eng, k, n, seed):
    from bls12_381_b200 import constants_host as ch
    dev = torch.device("cuda", eng.device)
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    t[:, 31] &= 0x3f
    g = torch.from_numpy(np.tile(ch.generator_projective(k), (n, 1))).to(dev)
    pr = torch.empty_like(g)
    eng.mul_batch_dev(k, g, torch.from_numpy(t).to(dev), pr, n)
    xy = torch.empty((n, 12 * k), dtype=torch.int64, device=dev)
    inf = torch.empty(n, dtype=torch.uint8, device=dev)
    eng.batch_normalize_dev(k, pr, n, xy, inf)
    s = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    s[:, 31] &= 0x3f
    return xy, inf, torch.from_numpy(s).to(dev), s


def _affine(eng, k, projs):
    """canonical affine limbs of a few projective results (device tensors (1,18k))"""
    both = torch.cat(projs).contiguous()
    m = both.shape[0]
    axy = torch.empty((m, 12 * k), dtype=torch.int64, device=both.device)
    ainf = torch.empty(m, dtype=torch.uint8, device=both.device)
    eng.batch_normalize_dev(k, both, m, axy, ainf)
    return axy.cpu().numpy().view(np.uint64), ainf.cpu().numpy()


@pytest.mark.parametrize("k,log2n", [(1, 20), (2, 20)])
def test_msm_full_size(eng, orc, k, log2n):
    n = 1 << log2n
    G = orc.G1 if k == 1 else orc.G2
    xy, inf, sc, s_host = _points(eng, k, n, 4242 + k)
    dev = xy.device
    out = [torch.empty((1, 18 * k), dtype=torch.int64, device=dev) for _ in range(8)]
    eng.msm_dev(k, xy, inf, sc, n, out[0])
    # (b1) additivity over a split of the point range
    h = n // 2 + 12345
    eng.msm_dev(k, xy[:h], inf[:h], sc[:h], h, out[1])
    eng.msm_dev(k, xy[h:], inf[h:], sc[h:], n - h, out[2])
    eng.sum_dev(k, torch.cat([out[1], out[2]]).contiguous(), 2, out[3])
    # (b2) the window-shard partials of 3 shards add up to the whole (what the multi-GPU path exchanges)
    parts = torch.empty((3, 18 * k), dtype=torch.int64, device=dev)
    for r in range(3):
        eng.msm_dev(k, xy, inf, sc, n, parts[r:r + 1], shard=r, n_shards=3)
    eng.sum_dev(k, parts, 3, out[4])
    # (b3) linearity in the scalars: MSM(s) + MSM(s2) == MSM(s + s2 mod q) on a 2^16 slice
    m = 1 << 16
    rng = np.random.default_rng(9)
    s2 = rng.integers(0, 256, (m, 32), dtype=np.uint8)
    s2[:, 31] &= 0x3f
    ssum = np.empty((m, 32), np.uint8)
    for i in range(m):
        v = (int.from_bytes(s_host[i].tobytes(), "little") + int.from_bytes(s2[i].tobytes(), "little")) % pyref.Q
        ssum[i] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
    eng.msm_dev(k, xy[:m], inf[:m], sc[:m], m, out[5])
    eng.msm_dev(k, xy[:m], inf[:m], torch.from_numpy(s2).to(dev), m, out[6])
    eng.sum_dev(k, torch.cat([out[5], out[6]]).contiguous(), 2, out[5])
    eng.msm_dev(k, xy[:m], inf[:m], torch.from_numpy(ssum).to(dev), m, out[6])
    a, ai = _affine(eng, k, [out[0], out[3], out[4], out[5], out[6]])
    assert np.array_equal(a[0], a[1]) and ai[0] == ai[1] == 0, "split-range additivity"
    assert np.array_equal(a[0], a[2]) and ai[2] == 0, "window-shard partials"
    assert np.array_equal(a[3], a[4]) and ai[3] == ai[4], "linearity in the scalars"
    # (a) the same inputs through the CPU oracle's Pippenger, bit-exact on the affine result
    threads = min(32, orc.hardware_threads())
    exp = G.to_affine(G.msm_pippenger(xy.cpu().numpy().view(np.uint64), inf.cpu().numpy(), s_host, c=16, threads=threads))
    assert np.array_equal(a[0], exp[0][0]) and ai[0] == exp[1][0], "full-size MSM differs from the oracle"


def test_pairing_full_size(eng, orc):
    n = 1 << 16
    pxy, pinf, _, _ = _points(eng, 1, n, 777)
    qxy, qinf, _, _ = _points(eng, 2, n, 778)
    dev = pxy.device
    gt = torch.empty((n, 72), dtype=torch.int64, device=dev)
    eng.set_tuning("pairing_variant", 7)                               # six lanes per pairing at the full 2^16
    eng.pairing_batch_dev(pxy, pinf, qxy, qinf, n, gt)
    # ALL 2^16 pairs against the oracle, limb-exact (the oracle runs on the host cores: ~65536 x 0.7 ms / threads)
    threads = min(64, orc.hardware_threads())
    exp = orc.pairing(pxy.cpu().numpy().view(np.uint64), pinf.cpu().numpy(), qxy.cpu().numpy().view(np.uint64),
                      qinf.cpu().numpy(), threads=threads)
    assert np.array_equal(gt.cpu().numpy().view(np.uint64), exp)
    # the one-thread-per-pairing kernels (pairing_variant 4; what the default picks at this size): chunked (4 chunks on two
    # streams) == unchunked == variant 7 == default
    eng.set_tuning("pairing_variant", 4)
    try:
        gt4 = torch.empty_like(gt)
        eng.pairing_batch_dev(pxy, pinf, qxy, qinf, n, gt4)
        assert torch.equal(gt, gt4)
        eng.set_tuning("pairing_chunks", 1)
        eng.pairing_batch_dev(pxy, pinf, qxy, qinf, n, gt4)
        eng.set_tuning("pairing_chunks", 4)
        assert torch.equal(gt, gt4)
        eng.set_tuning("pairing_variant", 0)
        eng.pairing_batch_dev(pxy, pinf, qxy, qinf, n, gt4)
        assert torch.equal(gt, gt4)
    finally:
        eng.set_tuning("pairing_variant", 0)
    # product mode over the full batch: prod_i ML(p_i, q_i) then ONE final exponentiation == prod_i Gt_i on a slice
    m = 256
    ml = torch.empty((m, 72), dtype=torch.int64, device=dev)
    eng.miller_loop_batch_dev(pxy[:m], pinf[:m], qxy[:m], qinf[:m], m, ml)
    prod = torch.empty((1, 72), dtype=torch.int64, device=dev)
    eng.fp12_product_dev(ml, m, prod)
    fe = torch.empty_like(prod)
    eng.final_exponentiation_batch_dev(prod, 1, fe)
    prod_gt = torch.empty((1, 72), dtype=torch.int64, device=dev)
    eng.fp12_product_dev(gt[:m].contiguous(), m, prod_gt)
    assert torch.equal(fe, prod_gt)
