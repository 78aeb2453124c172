"""Times the default G2 MSM against the EXPERIMENTAL second build of the MSM unit (capi_msm_lazy3.cu: row-alternated lazy
Fp2 multiply) on the same device-resident inputs, after checking that both give the same point.  Not part of bench.py.

    python tools/bench_g2_msm_variant.py [--log-n 20] [--steps 5] [--warmup 2]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    import torch
    import bls12_381_b200 as b
    import bench as B                                   # reuse bench.py's seeded input generator
    eng = b.Engine()
    n = 1 << a.log_n
    dev = torch.device("cuda", eng.device)
    # the same recipe as bench.py: P_i = [t_i]G2 with the parity-tested config-1 kernel + batch_normalize, random scalars
    from bls12_381_b200 import constants_host as ch
    t = torch.from_numpy(B.rand_scalars(0xB1512381 * 7, n).copy()).to(dev)
    g = torch.from_numpy(np.tile(ch.generator_projective(2), (n, 1))).to(dev)
    pr = torch.empty_like(g)
    torch.cuda.synchronize()
    eng.mul_batch_dev(2, g, t, pr, n)
    xy = torch.empty((n, 24), dtype=torch.int64, device=dev)
    inf = torch.empty(n, dtype=torch.uint8, device=dev)
    eng.batch_normalize_dev(2, pr, n, xy, inf)
    s = torch.from_numpy(B.rand_scalars(0xB1512381 * 13, n).copy()).to(dev)
    del g, t, pr
    torch.cuda.synchronize()
    lazy3 = eng.lib.b200x_lazy3_g2_msm_dev
    lazy3.argtypes = [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p]
    lazy3.restype = C.c_int
    out0 = torch.zeros((1, 36), dtype=torch.int64, device=dev)
    out1 = torch.zeros((1, 36), dtype=torch.int64, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None

    def run_default():
        eng.msm_dev(2, xy, inf, s, n, out0)

    def run_lazy3():
        assert lazy3(eng.h, p(xy), p(inf), p(s), n, p(out1)) == 0

    run_default()
    run_lazy3()
    torch.cuda.synchronize()
    a0, _ = eng.batch_normalize(2, out0.cpu().numpy().view(np.uint64))
    a1, _ = eng.batch_normalize(2, out1.cpu().numpy().view(np.uint64))
    assert np.array_equal(a0, a1), "the two builds disagree"
    res = {}
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    for name, fn in (("default", run_default), ("lazy3", run_lazy3), ("default_again", run_default)):
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(a.steps):
                fn()
            e1.record(stream)
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"workload": "g2_msm_2^%d" % a.log_n, "ms_per_step": res, "steps": a.steps, "warmup": a.warmup}))
    eng.close()
    os._exit(0)


if __name__ == "__main__":
    main()
