"""One window shard of a 2^20 G1 MSM, a few times (for an ncu launch list):  python tools/shard_once.py <shard> <n_shards> [key=value ...]"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bls12_381_b200 as b  # noqa: E402
from bls12_381_b200 import constants_host as ch  # noqa: E402

shard, ns = int(sys.argv[1]), int(sys.argv[2])
n = 1 << 20
eng = b.Engine()
for kv in sys.argv[3:]:
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
for _ in range(3):
    eng.msm_dev(1, xy, inf, sc, n, out, shard=shard, n_shards=ns)
torch.cuda.synchronize()
eng.close()
