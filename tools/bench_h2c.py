"""Standalone measurement of batched hash_to_curve — NOT part of bench.py's contract.

    python tools/bench_h2c.py [--log-n 16] [--steps 5] [--warmup 2] [--msg-len 32]

One JSON line per group (G1, G2): messages/s through the host-pointer entry point (the messages live on the host: H2D of the
messages and D2H of the projective points are inside the timed region), per-kernel device times from the library's event
records, the integer roofline fraction of the map kernel (model: 1 200 FpM per G1 point, 7 000 per G2 point — DESIGN.md
§4.6 — x 300 multiply-adds) and the oracle's time on a sample.  Parity with the oracle is checked on the sample first."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
MODEL_FPM = {1: 1200.0, 2: 7000.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--msg-len", type=int, default=32)
    a = ap.parse_args()
    import bls12_381_b200 as b
    from oracle import pyoracle as orc
    eng = b.Engine()
    n = 1 << a.log_n
    rng = np.random.default_rng(0xB1512381 + 7)
    raw = rng.bytes(a.msg_len * n)
    msgs = [raw[i * a.msg_len:(i + 1) * a.msg_len] for i in range(n)]
    peak, _ = eng.imad_peak(2000)
    for k in (1, 2):
        dst = b"QUUX-V01-CS02-with-BLS12381G%d_XMD:SHA-256_SSWU_RO_" % k
        ns = 256
        t0 = time.perf_counter()
        want = orc.hash_to_curve(k, msgs[:ns], dst)
        cpu_s = time.perf_counter() - t0
        assert np.array_equal(eng.hash_to_curve(k, msgs[:ns], dst), want), "parity with the oracle failed"
        for _ in range(a.warmup):
            eng.hash_to_curve(k, msgs, dst)
        eng.set_timing(True)
        l0 = eng.launches
        t0 = time.perf_counter()
        for _ in range(a.steps):
            eng.hash_to_curve(k, msgs, dst)
        wall = (time.perf_counter() - t0) / a.steps
        launches = eng.launches - l0
        per = {}
        for name, ms in eng.get_timing():
            per[name] = per.get(name, 0.0) + ms / a.steps
        eng.set_timing(False)
        kmap = max(per.items(), key=lambda kv: kv[1])
        achieved = MODEL_FPM[k] * 300 * n / (kmap[1] * 1e-3)
        print(json.dumps({"metric": "G%d hash_to_curve messages/s" % k, "value": n / wall, "unit": "messages/s", "n_gpus": 1,
                          "steps": a.steps, "warmup": a.warmup, "ms_per_step": wall * 1e3, "higher_is_better": True,
                          "data": "synthetic", "config": {"workload": "hash_to_curve_g%d_2^%d" % (k, a.log_n), "msg_len": a.msg_len,
                                                          "timing": "host wall clock around the host-pointer call (H2D of messages + kernels + D2H of points)"},
                          "gpu_launches": launches // a.steps, "kernel_ms": per,
                          "roofline": {"bound": "int", "kernel": kmap[0], "achieved": achieved, "peak": peak, "unit": "IMAD32/s",
                                       "frac": achieved / peak, "model": "%d FpM per point x 300" % MODEL_FPM[k]},
                          "cpu_baseline": {"value": ns / cpu_s, "unit": "messages/s", "cores": 1, "kind": "port",
                                           "sample": "%d messages, oracle, one thread" % ns}}))
    eng.close()


if __name__ == "__main__":
    main()
