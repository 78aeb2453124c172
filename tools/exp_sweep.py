#!/usr/bin/env python
"""Round-2 tuning sweeps on one GPU (not a bench line: device-timed kernels, per-kernel CUDA-event records of the library).

    python tools/exp_sweep.py pairing  [--log2n 16]     pairing batch: variant x coop_warps
    python tools/exp_sweep.py mul                        scalar-multiplication batch: n x mul_groups
    python tools/exp_sweep.py products                   products of pairings (shared squaring) against n independent loops
Prints one JSON object per measurement."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def timed(torch, stream, fn, reps, flush):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps)]
    with torch.cuda.stream(stream):
        fn()
        stream.synchronize()
        ms = []
        for r in range(reps):
            flush.fill_(r)
            ev[2 * r].record(stream)
            fn()
            ev[2 * r + 1].record(stream)
        stream.synchronize()
        for r in range(reps):
            ms.append(ev[2 * r].elapsed_time(ev[2 * r + 1]))
    return ms


def kernel_ms(eng, fn, stream):
    eng.set_timing(True)
    fn()
    stream.synchronize()
    rec = eng.get_timing()
    eng.set_timing(False)
    out = {}
    for name, ms in rec:
        out[name] = out.get(name, 0.0) + ms
    return {k: round(v, 3) for k, v in out.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["pairing", "mul", "products"])
    ap.add_argument("--log2n", type=int, default=16)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--warps", default="12")
    ap.add_argument("--variants", default="7,4")
    a = ap.parse_args()
    import torch
    import bls12_381_b200
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    eng = bls12_381_b200.Engine(0, stream=stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    class A:  # the slice of bench.Bench that gen_points needs
        pass
    bb = A()
    bb.torch, bb.eng, bb.dev = torch, eng, dev
    gen = lambda kk, count, sd: B.Bench.gen_points(bb, kk, count, sd)  # noqa: E731
    with torch.cuda.stream(stream):
        if a.what == "pairing":
            n = 1 << a.log2n
            pxy, pinf, _ = gen(1, n, 77)
            qxy, qinf, _ = gen(2, n, 78)
            out = torch.empty((n, 72), dtype=torch.int64, device=dev)
            stream.synchronize()
            fn = lambda: eng.pairing_batch_dev(pxy, pinf, qxy, qinf, n, out)  # noqa: E731
            ref = None
            for var in [int(v) for v in a.variants.split(",")]:
                eng.set_tuning("pairing_variant", var)
                for w in ([int(x) for x in a.warps.split(",")] if var == 7 else [0]):
                    if var == 7:
                        eng.set_tuning("coop_warps", w)
                    ms = timed(torch, stream, fn, a.reps, flush)
                    km = kernel_ms(eng, fn, stream)
                    h = int(out.sum().item()) & 0xffffffff
                    if ref is None:
                        ref = h
                    print(json.dumps({"what": "pairing", "n": n, "variant": var, "coop_warps": w,
                                      "ms": [round(x, 3) for x in ms], "kernels": km, "same_result": h == ref}), flush=True)
        elif a.what == "products":
            # Groth16 / BLS batch-verification shape: P products of T pairs each, one Gt per product
            for T, P in ((3, 4096), (4, 4096), (3, 16384)):
                n = T * P
                pxy, pinf, _ = gen(1, n, 177 + T)
                qxy, qinf, _ = gen(2, n, 178 + T)
                out = torch.empty((P, 72), dtype=torch.int64, device=dev)
                ml = torch.empty((n, 72), dtype=torch.int64, device=dev)
                stream.synchronize()
                shared = lambda: eng.pairing_product_batch_dev(pxy, pinf, qxy, qinf, T, P, out)  # noqa: E731

                def loops():   # without the product mode: n independent Miller loops, then P final exponentiations (the T - 1
                    eng.miller_loop_batch_dev(pxy, pinf, qxy, qinf, n, ml)       # Fp12 products per item are left out: favours
                    eng.final_exponentiation_batch_dev(ml, P, out)               # this baseline)
                ms_s = timed(torch, stream, shared, a.reps, flush)
                ms_l = timed(torch, stream, loops, a.reps, flush)
                print(json.dumps({"what": "pairing products", "terms": T, "products": P, "shared_squaring_ms": [round(x, 3) for x in ms_s],
                                  "n_loops_ms": [round(x, 3) for x in ms_l], "speedup": round(min(ms_l) / min(ms_s), 3),
                                  "kernels": kernel_ms(eng, shared, stream)}), flush=True)
            # one product of 2^16 terms -> one MillerLoopResult
            n = 1 << 16
            pxy, pinf, _ = gen(1, n, 277)
            qxy, qinf, _ = gen(2, n, 278)
            one = torch.empty((1, 72), dtype=torch.int64, device=dev)
            stream.synchronize()
            fn = lambda: eng.multi_miller_loop_dev(pxy, pinf, qxy, qinf, n, one)  # noqa: E731
            res = {}
            for var in (0, 4):
                eng.set_tuning("pairing_variant", var)
                res[var] = timed(torch, stream, fn, a.reps, flush)
                res["h%d" % var] = int(one.sum().item())
            eng.set_tuning("pairing_variant", 0)
            print(json.dumps({"what": "multi_miller_loop, one product", "n": n, "shared_squaring_ms": [round(x, 3) for x in res[0]],
                              "n_loops_plus_product_ms": [round(x, 3) for x in res[4]], "speedup": round(min(res[4]) / min(res[0]), 3),
                              "same_result": res["h0"] == res["h4"]}), flush=True)
        else:
            from bls12_381_b200 import constants_host as ch
            for k in (1, 2):
                for n in ([1024, 2048, 4096, 8192, 16384, 32768] if k == 1 else [1024, 4096]):
                    t = torch.from_numpy(B.rand_scalars(5 + n, n).copy()).to(dev)
                    g = torch.from_numpy(np.tile(ch.generator_projective(k), (n, 1))).to(dev)
                    _, _, pr = gen(k, n, 99 + n)          # random projective points as bases
                    o = torch.empty_like(g)
                    fn = lambda: eng.mul_batch_dev(k, pr, t, o, n)  # noqa: E731
                    ref = None
                    for mg in (6, -1, 1, 2, 3, 4, 5, 0):
                        if mg == 6 and n > 8192:
                            continue
                        eng.set_tuning("mul_groups", mg)
                        ms = timed(torch, stream, fn, a.reps, flush)
                        h = int(o.sum().item()) & 0xffffffff
                        if ref is None:
                            ref = h
                        print(json.dumps({"what": "mul", "k": k, "n": n, "mul_groups": mg, "ms": [round(x, 4) for x in ms],
                                          "same_result": h == ref}), flush=True)
                    eng.set_tuning("mul_groups", 0)
    flush = None
    torch.cuda.synchronize()
    eng.close()


if __name__ == "__main__":
    main()
