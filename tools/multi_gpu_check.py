#!/usr/bin/env python
"""torchrun check of the one-process-per-GPU path inside the library (b200_ctx_comm_init + b200_g{1,2}_msm_sharded_dev):
every rank builds the same inputs, runs the collective MSM in both sharding modes and compares with the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import bls12_381_b200
    from oracle import pyoracle as orc
    from tests import util
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    eng = bls12_381_b200.Engine(lr, stream=stream.cuda_stream)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    eng.comm_init(uid.cpu().numpy().tobytes(), rank, world)
    ok = True
    for k in (1, 2):
        G = orc.G1 if k == 1 else orc.G2
        rng = np.random.default_rng(20000 + k)              # same inputs on every rank
        n = 5000
        _, xy, inf = util.rand_points(orc, k, rng, n)
        s = util.rand_scalars(rng, n)
        inf[7] = 1
        t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(dev)
        want = G.to_affine(G.msm_pippenger(xy, inf, s, c=8, threads=4))
        out = torch.empty((1, 18 * k), dtype=torch.int64, device=dev)
        for mode in ("points", "window"):
            eng.msm_sharded_dev(k, t(xy), t(inf), t(s), n, out, mode=mode)
            got = G.to_affine(out.cpu().numpy().view(np.uint64))
            good = bool(np.array_equal(got[0], want[0]) and got[1][0] == want[1][0])
            ok &= good
            print("rank %d G%d %s-sharded over %d ranks: %s" % (rank, k, mode, world, "ok" if good else "MISMATCH"), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    dist.barrier()
    eng.close()
    dist.destroy_process_group()
    if rank == 0:
        print("multi_gpu_check:", "PASS" if int(flag.item()) == 1 else "FAIL", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
