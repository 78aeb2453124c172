#!/usr/bin/env python
"""bench.py — the BASELINE.json metric on B200: G1/G2 MSM point-scalar-muls/sec and pairings/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload g1_msm|g2_msm|pairing|g1_mul]
                    [--log2n L] [--impl b200|reference] [--shard window|points]

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload = BASELINE
configs[1]: G1 Pippenger MSM, 2^20 random subgroup points x random Scalars, on 1 GPU.  Rank 0 prints ONE
JSON line.  `value` = device-resident throughput (inputs already in HBM, CUDA events on the engine's
stream, L2 flushed between steps); `e2e` = the same call through the host-pointer C ABI (pinned host
buffers, H2D/D2H inside the timed region); `roofline` = the dominant kernel against the measured
IMAD.WIDE peak of this GPU (the path is integer-ALU bound, not HBM bound — HBM GB/s is reported
beside it); `cpu_baseline` = the oracle port of the reference's own path timed on the host cores.

--impl reference times that CPU path alone (the Rust reference cannot be built here: no rustc/cargo; the
oracle is its op-for-op C++ restatement, kind "port").
N>1 (torchrun): the MSM is window-sharded as north_star asks — every rank holds all points/scalars,
computes the partial sum of its windows, one NCCL all-gather of the 144-byte partials, local combine.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

Q = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
SEED_BASE = 0xB1512381            # seed = 0xB1512381 + config_id
CONFIG_ID = {"g1_mul": 1, "g1_msm": 2, "g2_msm": 3, "pairing": 4}
# SURVEY §8d fixed algorithmic-work model (reference formula costs), in 32x32+64 multiply-adds per unit
IMAD_PER_FPM = 300
MODEL_FPM = {"g1_mul": 5100.0, "g1_msm": 200.0, "g2_msm": 600.0, "pairing": 16020.0}
UNIT = {"g1_mul": "G1 scalar-muls/s", "g1_msm": "G1 MSM point-scalar-muls/s", "g2_msm": "G2 MSM point-scalar-muls/s",
        "pairing": "pairings/s"}
DOMINANT = {"g1_mul": "k_mul_batch_warp", "g1_msm": "k_msm_accumulate", "g2_msm": "k_msm_accumulate",
            "pairing": "k_final_exp"}


def rand_scalars(seed, n):
    """n canonical 32-byte LE scalars: 64 random bytes reduced mod q (mirrors Scalar::random ->
    from_bytes_wide, src/scalar.rs:646-650, :300-331)."""
    rng = np.random.default_rng(seed)
    if n > (1 << 21):
        # very large batches (config 5): 254 uniform random bits (< q), vectorised — statistically equivalent for
        # the bucket distribution; the exact from_bytes_wide reduction below is a Python big-int loop
        out = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        out[:, 31] &= 0x3f
        return out
    raw = rng.bytes(64 * n)
    out = bytearray(32 * n)
    for i in range(n):
        v = int.from_bytes(raw[64 * i:64 * i + 64], "little") % Q
        out[32 * i:32 * i + 32] = v.to_bytes(32, "little")
    return np.frombuffer(bytes(out), np.uint8).reshape(n, 32)


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in o.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            self._halt.wait(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(float(r[0])) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [int(float(r[1])) for r in self.rows if r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_reference(workload, n_sample, seed, threads):
    """the reference's own CPU path (oracle port) on a bounded sample; returns (units, seconds)"""
    from oracle import pyoracle as orc
    orc.build()
    rng = np.random.default_rng(seed)
    s = rng.integers(0, 256, (n_sample, 32), dtype=np.uint8)
    s[:, 31] &= 0x3f
    t = rng.integers(0, 256, (min(n_sample, 64), 32), dtype=np.uint8)
    t[:, 31] &= 0x3f
    reps = (n_sample + t.shape[0] - 1) // t.shape[0]
    if workload in ("g1_msm", "g1_mul", "g2_msm"):
        G = orc.G1 if workload != "g2_msm" else orc.G2
        base = G.mul(np.repeat(G.generator(), t.shape[0], 0), t, threads=threads)
        xy, inf = G.batch_normalize(base)
        xy = np.tile(xy, (reps, 1))[:n_sample]
        inf = np.tile(inf, reps)[:n_sample]
        t0 = time.perf_counter()
        if workload == "g1_mul":
            G.mul(G.from_affine(xy, inf), s, threads=threads)
        else:
            G.msm_naive(xy, inf, s, threads=threads)            # sum_i p_i * s_i, SURVEY §3.2
        return n_sample, time.perf_counter() - t0
    b1 = orc.G1.batch_normalize(orc.G1.mul(np.repeat(orc.G1.generator(), t.shape[0], 0), t, threads=threads))
    b2 = orc.G2.batch_normalize(orc.G2.mul(np.repeat(orc.G2.generator(), t.shape[0], 0), t, threads=threads))
    pxy, pinf = np.tile(b1[0], (reps, 1))[:n_sample], np.tile(b1[1], reps)[:n_sample]
    qxy, qinf = np.tile(b2[0], (reps, 1))[:n_sample], np.tile(b2[1], reps)[:n_sample]
    t0 = time.perf_counter()
    orc.pairing(pxy, pinf, qxy, qinf, threads=threads)
    return n_sample, time.perf_counter() - t0


def cpu_sample_size(workload, cores):
    # ~10-30 s of CPU work in total: reference cost/unit ~ FpM * 45 ns
    per_unit = {"g1_mul": 5100, "g1_msm": 5100, "g2_msm": 17085, "pairing": 16020}[workload] * 45e-9
    n = int(15.0 / per_unit)
    if os.environ.get("B200_BENCH_CPU_SAMPLE"):        # tests shrink the sample (tests/test_bench_cli.py)
        return int(os.environ["B200_BENCH_CPU_SAMPLE"])
    return max(64, min(n, 1 << 16))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="g1_msm", choices=list(CONFIG_ID))
    ap.add_argument("--log2n", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shard", default="window", choices=["window", "points"])
    ap.add_argument("--window", type=int, default=0, help="MSM window bits (0 = auto)")
    ap.add_argument("--tune", action="append", default=[], help="key=value tuning knob (b200_ctx_set_tuning)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    wl = a.workload
    log2n = a.log2n if a.log2n is not None else {"g1_mul": 10, "g1_msm": 20, "g2_msm": 20, "pairing": 16}[wl]
    n = 1 << log2n
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    seed = SEED_BASE + CONFIG_ID[wl]
    cfg_name = {"g1_mul": "%d x G1Projective scalar-mul batch", "g1_msm": "G1 Pippenger MSM, 2^%d random points x random Scalars",
                "g2_msm": "G2 Pippenger MSM, 2^%d random points x random Scalars",
                "pairing": "batched pairing: 2^%d (G1Affine,G2Affine) pairs -> Gt"}[wl] % (n if wl == "g1_mul" else log2n)

    # ------------------------------------------------------------------ reference arm (CPU path of the reference)
    if a.impl == "reference":
        if rank != 0:
            return
        from oracle import pyoracle as orc
        orc.build()
        cores = orc.hardware_threads()
        ns = cpu_sample_size(wl, cores)
        times = []
        for i in range(a.warmup + a.steps):
            units, sec = cpu_reference(wl, ns, seed + i, cores)
            if i >= a.warmup:
                times.append(sec)
        tot = sum(times)
        val = ns * len(times) / tot
        line = {"impl": "reference", "metric": UNIT[wl].replace("/s", "") + " per second", "value": val, "unit": UNIT[wl],
                "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot / len(times),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64 limbs (381-bit Montgomery)",
                "data": "synthetic", "config": {"workload": cfg_name, "sample": "%d units per step" % ns},
                "cpu_baseline": {"value": val, "unit": UNIT[wl], "cores": cores, "kind": "port",
                                 "sample": "%d-unit slice per step of the reference's constant-time path (oracle C++ port; "
                                           "the Rust crate cannot be built in this image), all %d host threads" % (ns, cores)},
                "e2e": {"value": val, "unit": UNIT[wl], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import bls12_381_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = bls12_381_b200.Engine(local_rank)
    if a.window:
        eng.set_msm_window(a.window)
    for kv in a.tune:
        key, val = kv.split("=")
        eng.set_tuning(key, int(val))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    k = 2 if wl == "g2_msm" else 1
    AFFW, PROJW = 12 * k, 18 * k

    def gen_points(kk, count, sd):
        """[t_i]G on the GPU with the config-1 kernel + batch_normalize (both parity-tested)"""
        from bls12_381_b200 import constants_host as ch
        t = torch.from_numpy(rand_scalars(sd, count).copy()).to(dev)
        g = torch.from_numpy(np.tile(ch.generator_projective(kk), (count, 1))).to(dev)
        pr = torch.empty_like(g)
        eng.mul_batch_dev(kk, g, t, pr, count)
        xy = torch.empty((count, 12 * kk), dtype=torch.int64, device=dev)
        inf = torch.empty(count, dtype=torch.uint8, device=dev)
        eng.batch_normalize_dev(kk, pr, count, xy, inf)
        return xy, inf, pr

    # per-rank share of the batch for the embarrassingly parallel workloads
    if wl in ("pairing", "g1_mul"):
        n_local = n // world
        off = rank * n_local
    else:
        n_local = n
        off = 0
    t_gen = time.perf_counter()
    if wl == "pairing":
        pxy, pinf, _ = gen_points(1, n_local, seed * 7 + rank)
        qxy, qinf, _ = gen_points(2, n_local, seed * 11 + rank)
        out = torch.empty((n_local, 72), dtype=torch.int64, device=dev)
    else:
        xy, inf, pr = gen_points(k, n_local, seed * 7 + (rank if wl == "g1_mul" else 0))
        sc_host = rand_scalars(seed * 13 + (rank if wl == "g1_mul" else 0), n_local)
        sc = torch.from_numpy(sc_host.copy()).to(dev)
        out = torch.empty((max(n_local, 1) if wl == "g1_mul" else 1, PROJW), dtype=torch.int64, device=dev)
        parts = torch.empty((world, PROJW), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    from bls12_381_b200.sharding import ShardedMSM
    sharded = ShardedMSM(eng, k, dist=dist, stream=stream, mode=a.shard) if wl in ("g1_msm", "g2_msm") else None

    def step_device():
        if wl == "g1_mul":
            eng.mul_batch_dev(1, pr, sc, out, n_local)
        elif wl == "pairing":
            eng.pairing_batch_dev(pxy, pinf, qxy, qinf, n_local, out)
        else:
            sharded.msm(xy, inf, sc, n_local, out, parts)            # bls12_381_b200/sharding.py

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(a.warmup):
        step_device()
    barrier()
    verified = None
    if world > 1 and sharded is not None:
        # outside the timed region: the sharded result must be the same group element as the one-GPU MSM
        full = torch.empty_like(out)
        eng.msm_dev(k, xy, inf, sc, n_local, full)
        both = torch.cat([out, full]).contiguous()
        axy = torch.empty((2, AFFW), dtype=torch.int64, device=dev)
        ainf = torch.empty(2, dtype=torch.uint8, device=dev)
        eng.batch_normalize_dev(k, both, 2, axy, ainf)
        verified = bool(torch.equal(axy[0], axy[1]) and ainf[0] == ainf[1])
        if not verified:
            raise SystemExit("bench.py: sharded MSM result differs from the single-GPU result on rank %d" % rank)
    eng.set_timing(True)
    launches0 = eng.launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for i in range(a.steps):
        flush.zero_()                                                # L2 flush, outside the timed events
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ev[i][0].record(stream)
        step_device()
        ev[i][1].record(stream)
    barrier()
    clocks = sampler.stop()
    launches = eng.launches - launches0
    timing = eng.get_timing()
    eng.set_timing(False)
    ms_steps = [e0.elapsed_time(e1) for e0, e1 in ev]
    total_ms = sum(ms_steps)
    if dist is not None:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / a.steps
    units_per_step = n                                              # whole-job units (all ranks together)
    value = units_per_step / (ms_per_step * 1e-3)

    # ------------------------------------------------------------------ e2e through the host-pointer C ABI
    e2e = None
    if not a.no_e2e and world > 1:
        # N > 1: every rank feeds its own GPU from pinned host memory inside the timed region (H2D of everything the rank
        # needs), runs its share through the public API, and rank-locally reads the result back (D2H); max over ranks.
        if wl == "pairing":
            hp = [torch.empty(x.shape, dtype=x.dtype).pin_memory().copy_(x) for x in (pxy, pinf, qxy, qinf)]
            h2d = sum(x.numel() * x.element_size() for x in hp) * world
            d2h = n_local * 576 * world

            def step_host():
                eng.pairing_batch(hp[0].numpy().view(np.uint64), hp[1].numpy(), hp[2].numpy().view(np.uint64), hp[3].numpy())
        elif wl == "g1_mul":
            hpr = torch.empty(pr.shape, dtype=pr.dtype).pin_memory().copy_(pr)
            hsc = torch.empty(sc.shape, dtype=sc.dtype).pin_memory().copy_(sc)
            h2d, d2h = (hpr.numel() * 8 + hsc.numel()) * world, hpr.numel() * 8 * world

            def step_host():
                eng.mul_batch(1, hpr.numpy().view(np.uint64), hsc.numpy())
        else:
            hxy = torch.empty(xy.shape, dtype=xy.dtype).pin_memory().copy_(xy)
            hinf = torch.empty(inf.shape, dtype=inf.dtype).pin_memory().copy_(inf)
            hsc = torch.empty(sc.shape, dtype=sc.dtype).pin_memory().copy_(sc)
            dxy, dinf, dsc = torch.empty_like(xy), torch.empty_like(inf), torch.empty_like(sc)
            hres = torch.empty((1, PROJW), dtype=torch.int64).pin_memory()
            h2d, d2h = (hxy.numel() * 8 + hinf.numel() + hsc.numel()) * world, PROJW * 8 * world

            def step_host():
                # copies on torch's own stream (pinned -> device), fenced before the engine's stream touches the data
                dxy.copy_(hxy, non_blocking=True)
                dinf.copy_(hinf, non_blocking=True)
                dsc.copy_(hsc, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                sharded.msm(dxy, dinf, dsc, n_local, out, parts)     # window-sharded MSM + all_gather + combine (synchronous)
                hres.copy_(out)
        for _ in range(2):
            step_host()
        barrier()
        ne = max(3, min(a.steps, 10))
        t0 = time.perf_counter()
        for _ in range(ne):
            step_host()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / ne
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
        e2e = {"value": n / (e2e_ms * 1e-3), "unit": UNIT[wl], "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e2e_ms, "timing": "host wall clock per rank (pinned H2D + public API + D2H), max over ranks, %d steps" % ne}
    if not a.no_e2e and world == 1:
        if wl == "pairing":
            hp = [torch.empty(x.shape, dtype=x.dtype).pin_memory().copy_(x) for x in (pxy, pinf, qxy, qinf)]
            hout = torch.empty((n_local, 72), dtype=torch.int64).pin_memory()
            h2d = sum(x.numel() * x.element_size() for x in hp)
            d2h = hout.numel() * 8

            def step_host():
                eng.pairing_batch(hp[0].numpy().view(np.uint64), hp[1].numpy(), hp[2].numpy().view(np.uint64), hp[3].numpy())
        else:
            hxy = torch.empty(xy.shape, dtype=xy.dtype).pin_memory().copy_(xy)
            hinf = torch.empty(inf.shape, dtype=inf.dtype).pin_memory().copy_(inf)
            hsc = torch.empty(sc.shape, dtype=sc.dtype).pin_memory().copy_(sc)
            hpr = torch.empty(pr.shape, dtype=pr.dtype).pin_memory().copy_(pr) if wl == "g1_mul" else None
            if wl == "g1_mul":
                h2d, d2h = hpr.numel() * 8 + hsc.numel(), hpr.numel() * 8

                def step_host():
                    eng.mul_batch(1, hpr.numpy().view(np.uint64), hsc.numpy())
            else:
                h2d, d2h = hxy.numel() * 8 + hinf.numel() + hsc.numel(), PROJW * 8

                def step_host():
                    eng.msm(k, hxy.numpy().view(np.uint64), hinf.numpy(), hsc.numpy())
        for _ in range(2):
            step_host()
        torch.cuda.synchronize()
        ne = max(3, min(a.steps, 10))
        t0 = time.perf_counter()
        for _ in range(ne):
            step_host()                                             # synchronous: returns after the D2H copy
        e2e_ms = (time.perf_counter() - t0) * 1e3 / ne
        e2e = {"value": n / (e2e_ms * 1e-3), "unit": UNIT[wl], "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e2e_ms, "timing": "host wall clock around the synchronous C-ABI call, %d steps" % ne}

    # ------------------------------------------------------------------ roofline of the dominant kernel
    roof = None
    if rank == 0:
        peak, peak_ms = eng.imad_peak(3000)
        per_kernel = {}
        for name, ms in timing:
            base = name.split("<")[0].strip("( ")
            if base.startswith("k_msm_accumulate"):
                base = "k_msm_accumulate"           # G1 / G2-register / G2-shared-memory variants of the bucket kernel
            per_kernel.setdefault(base, []).append(ms)
        dom = DOMINANT[wl]
        share = {kname: sum(v) for kname, v in per_kernel.items()}
        tot_k = sum(share.values()) or 1.0
        if dom in per_kernel:
            launches_dom = len(per_kernel[dom])
            # a step may launch the dominant kernel several times (one per window group): its algorithmic work per
            # STEP over its summed duration per STEP
            avg_ms = sum(per_kernel[dom]) / a.steps
            # algorithmic work of that kernel per launch, SURVEY §8d cost sheet (FpM x 300 IMAD32):
            if wl in ("g1_msm", "g2_msm"):
                c = a.window or 16
                nwin = (256 + c - 1) // c
                nwin_local = len(range(rank, nwin, world)) if (world > 1 and a.shard == "window") else nwin
                n_eff = n if (world == 1 or a.shard == "window") else n // world
                fpm = (11 if k == 1 else 33) * n_eff * nwin_local      # SURVEY 8d model: one complete mixed add per term per window
                fpm_exec = (10 if k == 1 else 30) * n_eff * nwin_local  # what the kernel executes: XYZZ madd, 8M+2S
            elif wl == "g1_mul":
                fpm = fpm_exec = 5100.0 * n_local
            else:
                fpm = fpm_exec = 9104.0 * n_local                        # final exponentiation kernel
            achieved = fpm * IMAD_PER_FPM / (avg_ms * 1e-3)
            hbm_peak = None
            try:
                hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            except Exception:
                hbm_peak = 6650.0
            alg_bytes = {"g1_msm": (96 + 4) * n * 16, "g2_msm": (192 + 4) * n * 16, "g1_mul": (144 * 2 + 32) * n_local,
                         "pairing": 576 * 2 * n_local}[wl]
            roof = {"bound": "int (IMAD.WIDE.U32 pipe; SURVEY 8d: not hbm, not tensor)", "kernel": dom,
                    "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "T IMAD32/s", "frac": achieved / peak,
                    "executed_frac": fpm_exec * IMAD_PER_FPM / (avg_ms * 1e-3) / peak,
                    "peak_source": "b200_imad_peak microbenchmark run live on this GPU (%.2f ms, dependent-free IMAD.WIDE.U32)" % peak_ms,
                    "model": "SURVEY 8d cost sheet x 300 IMAD32 per FpM", "kernel_ms_per_step": avg_ms, "kernel_launches_per_step": launches_dom / a.steps,
                    "kernel_share_of_step": share[dom] / tot_k,
                    # dram__bytes_read+write of this kernel per step from the ncu --set full capture under profiles/
                    # (497 MB for the 6-window launch of the 2^20 G1 MSM, scaled to the 16 windows of a step)
                    "traffic": (497.3e6 * 16 / 6 if (wl == "g1_msm" and log2n == 20 and world == 1) else None),
                    "traffic_source": "profiles/r01_ncu_full_k_msm_accumulate_g1_final.txt",
                    "hbm": {"achieved_gbs": alg_bytes / (avg_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                            "frac": alg_bytes / (avg_ms * 1e-3) / 1e9 / hbm_peak, "of": "measured"},
                    "kernel_ms": {kname: sum(v) / a.steps for kname, v in per_kernel.items()}}
            roof["model_frac_whole_step"] = MODEL_FPM[wl] * IMAD_PER_FPM * units_per_step / world / (ms_per_step * 1e-3) / peak

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import pyoracle as orc
        orc.build()
        cores = orc.hardware_threads()
        ns = cpu_sample_size(wl, cores)
        units, sec = cpu_reference(wl, ns, seed, cores)
        cpu = {"value": units / sec, "unit": UNIT[wl], "cores": cores, "kind": "port",
               "sample": "%d-unit slice of the same workload through the reference's constant-time path "
                         "(oracle C++ port), %d host threads, %.1f s wall" % (ns, cores, sec)}

    if rank == 0:
        line = {"metric": UNIT[wl].replace("/s", "") + " per second", "value": value, "unit": UNIT[wl], "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None,
                "dtype": "u32x12 limbs (381-bit Montgomery, integer)", "data": "synthetic",
                "config": {"workload": cfg_name, "n": n, "sharding": ("none" if world == 1 else
                           ("by pair/item index, no collective" if wl in ("pairing", "g1_mul") else
                            a.shard + "-sharded, one NCCL all-gather of partial sums")),
                           "l2": "256 MiB buffer written between timed steps (L2 flush)",
                           "input_generation_s": t_gen, "seed": hex(seed), "sharded_result_equals_single_gpu": verified},
                "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line))
    sys.stdout.flush()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    # leave without running interpreter-exit destructors: torch's allocators may still hold events tied to the engine's
    # (now destroyed) stream, and a teardown-order abort must not turn a finished measurement into a failed run
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
