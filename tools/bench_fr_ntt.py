"""Standalone measurement of the scalar-field NTT — NOT part of bench.py's contract.

    python tools/bench_fr_ntt.py [--log-n 24] [--steps 10] [--warmup 3] [--cpu-log-n 18]

Prints one JSON line: NTT butterflies/s with the data resident in HBM (device-pointer entry point, CUDA events on the
ctx stream), the end-to-end figure through the host-buffer entry point, the integer roofline fraction (one butterfly =
one Fr multiplication = 128 32x32+64 multiply-adds, against b200_imad_peak measured in the same process) and the CPU
oracle's O(n log n) transform timed beside it on a smaller size.  Correctness is checked first (round trip at the
timed size and parity with the oracle at --cpu-log-n)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-log-n", type=int, default=18)
    a = ap.parse_args()
    import torch
    import bls12_381_b200 as b
    from oracle import pyoracle as orc
    eng = b.Engine()
    rng = np.random.default_rng(0xB1512381 + 6)
    n = 1 << a.log_n
    host = np.ascontiguousarray(np.frombuffer(rng.bytes(32 * n), np.uint64).reshape(n, 4) & np.uint64(0x0fffffffffffffff))
    # parity at a size the CPU finishes quickly, round trip at the timed size
    small = host[:1 << a.cpu_log_n]
    t0 = time.perf_counter()
    want = orc.fr_ntt(small, threads=orc.hardware_threads())
    cpu_s = time.perf_counter() - t0
    assert np.array_equal(eng.fr_ntt(small), want), "parity with the oracle failed"
    dev = torch.from_numpy(host.view(np.int64)).cuda()
    out = torch.empty_like(dev)
    back = torch.empty_like(dev)
    torch.cuda.synchronize()                       # the engine works on its own non-blocking stream
    eng.fr_ntt_dev(dev, a.log_n, out)
    eng.fr_ntt_dev(out, a.log_n, back, inverse=True)
    torch.cuda.synchronize()
    assert torch.equal(back, dev), "round trip failed"
    stream = torch.cuda.ExternalStream(eng.stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(a.warmup):
        eng.fr_ntt_dev(dev, a.log_n, out)
    torch.cuda.synchronize()
    l0 = eng.launches
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(a.steps):
            eng.fr_ntt_dev(dev, a.log_n, out)      # 512 MiB in + 512 MiB out at 2^24: larger than L2, no flush needed
        ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / a.steps
    launches = (eng.launches - l0) // a.steps
    t0 = time.perf_counter()
    for _ in range(max(1, a.steps // 3)):
        eng.fr_ntt(host)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / max(1, a.steps // 3)
    peak, _ = eng.imad_peak()
    bf = (n // 2) * a.log_n
    print(json.dumps({
        "metric": "Fr NTT butterflies/s", "value": bf / ms * 1e3, "unit": "butterflies/s", "n_gpus": 1, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "dtype": "u32x8 (255-bit Montgomery)",
        "data": "synthetic", "config": {"workload": "fr_ntt_2^%d" % a.log_n, "l2": "inputs larger than L2"},
        "gpu_launches": int(launches),
        "e2e": {"value": bf / e2e_ms * 1e3, "unit": "butterflies/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 32 * n},
        "roofline": {"bound": "int", "achieved": bf * 128 / ms * 1e3, "peak": peak, "unit": "IMAD32/s",
                     "frac": bf * 128 / ms * 1e3 / peak, "traffic": None,
                     "hbm_GBps": 2 * 32 * n * launches / ms / 1e6},
        "cpu_baseline": {"value": (1 << a.cpu_log_n) // 2 * a.cpu_log_n / cpu_s, "unit": "butterflies/s",
                         "cores": orc.hardware_threads(), "kind": "port", "sample": "one 2^%d transform" % a.cpu_log_n},
    }))
    eng.close()
    os._exit(0)


if __name__ == "__main__":
    main()
