"""Per-shard time of a sharded 2^20 G1 MSM on ONE GPU (what each rank of an N-GPU run executes, without the NCCL gather):

    python tools/bench_shard.py [n_shards] [key=value ...]          e.g.  python tools/bench_shard.py 8 msm_tail_groups=0

Prints, for window sharding, the time of every shard (the step of the N-GPU run is the slowest one + gather + combine) and,
for point-range sharding, the time of one n / n_shards slice with all windows."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bls12_381_b200 as b  # noqa: E402
from bls12_381_b200 import constants_host as ch

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = 1 << 20
eng = b.Engine()
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    eng.set_tuning(k, int(v))
dev = torch.device("cuda", eng.device)
rng = np.random.default_rng(5)
t = rng.integers(0, 256, (n, 32), dtype=np.uint8); t[:, 31] &= 0x3f
g = torch.from_numpy(np.tile(ch.generator_projective(1), (n, 1))).to(dev)
pr = torch.empty_like(g)
eng.mul_batch_dev(1, g, torch.from_numpy(t).to(dev), pr, n)
xy = torch.empty((n, 12), dtype=torch.int64, device=dev); inf = torch.empty(n, dtype=torch.uint8, device=dev)
eng.batch_normalize_dev(1, pr, n, xy, inf)
s = rng.integers(0, 256, (n, 32), dtype=np.uint8); s[:, 31] &= 0x3f
sc = torch.from_numpy(s).to(dev)
out = torch.empty((1, 18), dtype=torch.int64, device=dev)
st = torch.cuda.ExternalStream(eng.stream, device=dev)


def timed(fn):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10


res = [timed(lambda r=r: eng.msm_dev(1, xy, inf, sc, n, out, shard=r, n_shards=ns)) for r in range(ns)]
m = n // ns
pt = timed(lambda: eng.msm_dev(1, xy[:m], inf[:m], sc[:m], m, out))
print("n_shards", ns, sys.argv[2:], "window shards ms:", ["%.3f" % x for x in res], "max %.3f" % max(res), "| point-range slice %.3f" % pt)
