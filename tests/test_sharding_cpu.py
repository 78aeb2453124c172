"""world_size-2 gloo test (CPU) of the multi-GPU host logic in bls12_381_b200/sharding.py: window-sharded and
point-sharded MSM partials, one all_gather, local combine == the full MSM.  The per-rank device calls are
replaced by a stand-in engine built on the CPU oracle (test infrastructure), so only the sharding /
collective / combine logic of the product is under test here; the CUDA kernels are covered by -m gpu."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C = 16


class OracleEngine:
    """msm_dev / sum_dev with the oracle; partial of a shard = MSM over the scalars masked to its windows"""

    def __init__(self):
        from oracle import pyoracle
        self.o = pyoracle

    def msm_dev(self, k, xy, inf, s, n, out, shard=0, n_shards=1):
        G = self.o.G1 if k == 1 else self.o.G2
        sb = s.numpy()[:n].copy()
        if n_shards > 1:
            bits = np.unpackbits(sb, axis=1, bitorder="little")
            keep = np.zeros(256, np.uint8)
            for w in range(shard, (256 + C - 1) // C, n_shards):
                keep[w * C:(w + 1) * C] = 1
            sb = np.packbits(bits * keep, axis=1, bitorder="little")
        r = G.msm_pippenger(xy.numpy().view(np.uint64)[:n], None if inf is None else inf.numpy()[:n], sb, c=8, threads=2)
        out.copy_(torch.from_numpy(r.view(np.int64)))

    def sum_dev(self, k, parts, n, out):
        G = self.o.G1 if k == 1 else self.o.G2
        acc = G.identity()
        p = parts.numpy().view(np.uint64)
        for i in range(n):
            acc = G.add(acc, p[i:i + 1])
        out.copy_(torch.from_numpy(acc.view(np.int64)))


    # ---- pairing product stand-ins
    def miller_loop_batch_dev(self, p, pinf, q, qinf, n, out):
        r = self.o.miller_loop(p.numpy().view(np.uint64)[:n], None if pinf is None else pinf.numpy()[:n],
                               q.numpy().view(np.uint64)[:n], None if qinf is None else qinf.numpy()[:n], threads=2)
        out[:n].copy_(torch.from_numpy(r.view(np.int64)))

    def fp12_product_dev(self, f, n, out):
        a = f.numpy().view(np.uint64)
        acc = a[0:1].copy()
        for i in range(1, n):
            acc = self.o.tower(12, "mul", acc, a[i:i + 1])
        out.copy_(torch.from_numpy(acc.view(np.int64)))

    def final_exponentiation_batch_dev(self, f, n, out):
        r = self.o.final_exponentiation(f.numpy().view(np.uint64)[:n])
        out[:n].copy_(torch.from_numpy(r.view(np.int64)))


def _pair_worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bls12_381_b200.sharding import ShardedPairingProduct
    from oracle import pyoracle as o
    rng = np.random.default_rng(77)                       # replicated inputs
    t = rng.integers(0, 256, (2 * n, 32), dtype=np.uint8)
    t[:, 31] &= 0x3f
    pxy, pinf = o.G1.batch_normalize(o.G1.mul(np.repeat(o.G1.generator(), n, 0), t[:n]))
    qxy, qinf = o.G2.batch_normalize(o.G2.mul(np.repeat(o.G2.generator(), n, 0), t[n:]))
    if n > 2:
        pinf[1] = 1                                       # identity terms are skipped (src/pairings.rs:566-569)
        qinf[n - 1] = 1
    T = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a)
    out = torch.zeros((1, 72), dtype=torch.int64)
    parts = torch.zeros((world, 72), dtype=torch.int64)
    scratch = torch.zeros(((n + world - 1) // world, 72), dtype=torch.int64)
    sp = ShardedPairingProduct(OracleEngine(), dist=dist)
    sp.multi_miller_loop(T(pxy), T(pinf), T(qxy), T(qinf), n, out, parts, scratch)
    want = o.multi_miller_loop(pxy, pinf, qxy, qinf)       # the reference's shared-squaring loop, restated
    ok = bool(np.array_equal(out.numpy().view(np.uint64), want))
    sp.multi_miller_loop(T(pxy), T(pinf), T(qxy), T(qinf), n, out, parts, scratch, final_exp=True)
    ok &= bool(np.array_equal(out.numpy().view(np.uint64), o.final_exponentiation(want)))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 5), (3, 2)])
def test_sharded_pairing_product_gloo(orc, world, n):
    """SURVEY §8(e) "pairing product": terms sharded by index, one all_gather of 576-byte partials, local product, one
    final exponentiation — equal to the single-process multi_miller_loop; (3, 2) leaves one rank without terms"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + (os.getpid() % 40) + world
    procs = [ctx.Process(target=_pair_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def _worker(rank, world, port, mode, k, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bls12_381_b200.sharding import ShardedMSM, index_range, windows_of
    from oracle import pyoracle as o
    G = o.G1 if k == 1 else o.G2
    rng = np.random.default_rng(99)                       # same inputs on every rank (replicated, like the GPU path)
    n = 23
    t = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    t[:, 31] &= 0x3f
    s = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    s[:, 31] &= 0x3f
    xy, inf = G.batch_normalize(G.mul(np.repeat(G.generator(), n, 0), t))
    inf[5] = 1
    txy, tinf, ts = torch.from_numpy(xy.view(np.int64)), torch.from_numpy(inf), torch.from_numpy(s)
    out = torch.zeros((1, 18 * k), dtype=torch.int64)
    parts = torch.zeros((world, 18 * k), dtype=torch.int64)
    ShardedMSM(OracleEngine(), k, dist=dist, mode=mode).msm(txy, tinf, ts, n, out, parts)
    full = G.to_affine(G.msm_naive(xy, inf, s))
    got = G.to_affine(out.numpy().view(np.uint64))
    ok = bool(np.array_equal(full[0], got[0]) and full[1][0] == got[1][0])
    # partition helpers: every window / index exactly once
    wins = sorted(sum((windows_of(r, world, C) for r in range(world)), []))
    ok &= wins == list(range(16))
    spans = [index_range(n, r, world) for r in range(world)]
    ok &= spans[0][0] == 0 and spans[-1][1] == n and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,k", [("window", 1), ("points", 1), ("window", 2)])
def test_sharded_msm_two_ranks_gloo(orc, mode, k):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300) + (0 if mode == "window" else 1) + 2 * k
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, k, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
