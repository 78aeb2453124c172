#!/usr/bin/env python
"""bench.py — the BASELINE.json metric on B200: G1/G2 MSM point-scalar-muls/sec and pairings/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|g1_msm|g2_msm|pairing|g1_mul]
                    [--log2n L] [--impl b200|reference] [--shard points|window]

A "step" is one pass of the hot path over one batch of synthetic input.  The headline (`metric`, `value`, `e2e`,
`roofline`, `cpu_baseline`) is BASELINE configs[1]: G1 Pippenger MSM, 2^20 random subgroup points x random Scalars.
With the default `--workload all` the SAME JSON line also carries, under `configs`, the other BASELINE configurations
measured the same way with fewer steps: `g1_mul_1024` (configs[0]), `g2_msm_2p20` (configs[2]), `pairing_2p16`
(configs[3]) and, when launched on 8 GPUs, `g1_msm_2p24` (configs[4]).  Rank 0 prints ONE JSON line.

  value     device-resident throughput (inputs already in HBM, CUDA events on the engine's stream, L2 flushed between steps)
  e2e       the same call through the host-pointer C ABI (pinned host buffers, H2D / D2H inside the timed region)
  roofline  the dominant kernel against the IMAD.WIDE peak measured live on this GPU (the path is
            integer-ALU bound, not HBM bound; HBM GB/s is reported beside it); `traffic` from the committed ncu capture
  cpu_baseline  the oracle port of the reference's own constant-time path on the host cores: all usable cores AND one thread

--impl reference times that CPU path alone (the Rust reference cannot be built here: no rustc/cargo; the oracle is its
op-for-op C++ restatement, kind "port").
N > 1 (torchrun, one rank per GPU): the library owns the NCCL communicator (b200_ctx_comm_init; torch.distributed only
ships the 128-byte id and does the barriers).  MSM: b200_g1_msm_sharded_dev = shard + ncclAllGather of the 144-byte
partials + combine on one stream.  Pairing / scalar-mul batches shard by index with no collective.
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
DEFAULT_LOG2N = {"g1_mul": 10, "g1_msm": 20, "g2_msm": 20, "pairing": 16}
# SURVEY §8d fixed algorithmic-work model (reference formula costs), in 32x32+64 multiply-adds per unit
IMAD_PER_FPM = 300
UNIT = {"g1_mul": "G1 scalar-muls/s", "g1_msm": "G1 MSM point-scalar-muls/s", "g2_msm": "G2 MSM point-scalar-muls/s",
        "pairing": "pairings/s"}
DOMINANT = {"g1_mul": "k_mul_batch_grp", "g1_msm": "k_msm_accumulate", "g2_msm": "k_msm_accumulate",
            "pairing": "k_coop_pairing"}


def model_fpm(wl, log2n):
    """SURVEY §8d cost sheet, FpM per unit: Pippenger model c = 16, W = 16, one complete mixed add per term and window
    plus 2 (2^c - 1) complete adds per window for the bucket reduction"""
    if wl == "g1_mul":
        return 5100.0
    if wl == "pairing":
        return 16020.0
    per_add, per_term = (12, 11) if wl == "g1_msm" else (36, 33)
    return 16.0 * per_term + 16 * 2 * 65535 * per_add / float(1 << log2n)


def _reduce_chunk(raw):
    out = bytearray(len(raw) // 2)
    for i in range(len(raw) // 64):
        out[32 * i:32 * i + 32] = (int.from_bytes(raw[64 * i:64 * i + 64], "little") % Q).to_bytes(32, "little")
    return bytes(out)


def rand_scalars(seed, n):
    """n canonical 32-byte LE scalars: 64 random bytes reduced mod q (mirrors Scalar::random -> from_bytes_wide,
    src/scalar.rs:646-650, :300-331).  Exact for every n; large n is reduced by a process pool."""
    rng = np.random.default_rng(seed)
    raw = rng.bytes(64 * n)
    if n <= (1 << 17):
        return np.frombuffer(_reduce_chunk(raw), np.uint8).reshape(n, 32)
    from concurrent.futures import ProcessPoolExecutor
    step = 64 << 14
    nproc = max(1, min(16, len(os.sched_getaffinity(0))))
    with ProcessPoolExecutor(nproc) as ex:
        parts = list(ex.map(_reduce_chunk, [raw[o:o + step] for o in range(0, len(raw), step)]))
    return np.frombuffer(b"".join(parts), np.uint8).reshape(n, 32)


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


# ---------------------------------------------------------------------------------------------------- CPU arm
def host_cores():
    """what this process may actually use: scheduler affinity and the cgroup CPU quota, next to the hardware thread count"""
    aff = len(os.sched_getaffinity(0))
    quota = None
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(per)
    except Exception:
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / per
        except Exception:
            pass
    used = aff if quota is None else max(1, min(aff, int(quota + 0.5)))
    return {"hardware_threads": os.cpu_count(), "affinity": aff, "cgroup_quota": quota, "used": used}


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
    gt = min(threads, 8)
    if workload in ("g1_msm", "g1_mul", "g2_msm"):
        G = orc.G1 if workload != "g2_msm" else orc.G2
        base = G.mul(np.repeat(G.generator(), t.shape[0], 0), t, threads=gt)
        xy, inf = G.batch_normalize(base)
        xy = np.tile(xy, (reps, 1))[:n_sample]
        inf = np.tile(inf, reps)[:n_sample]
        t0 = time.perf_counter()
        if workload == "g1_mul":
            G.mul(G.from_affine(xy, inf), s, threads=threads)
        else:
            G.msm_naive(xy, inf, s, threads=threads)            # sum_i p_i * s_i, SURVEY §3.2
        return n_sample, time.perf_counter() - t0
    b1 = orc.G1.batch_normalize(orc.G1.mul(np.repeat(orc.G1.generator(), t.shape[0], 0), t, threads=gt))
    b2 = orc.G2.batch_normalize(orc.G2.mul(np.repeat(orc.G2.generator(), t.shape[0], 0), t, threads=gt))
    pxy, pinf = np.tile(b1[0], (reps, 1))[:n_sample], np.tile(b1[1], reps)[:n_sample]
    qxy, qinf = np.tile(b2[0], (reps, 1))[:n_sample], np.tile(b2[1], reps)[:n_sample]
    t0 = time.perf_counter()
    orc.pairing(pxy, pinf, qxy, qinf, threads=threads)
    return n_sample, time.perf_counter() - t0


CPU_FPM = {"g1_mul": 5100, "g1_msm": 5100, "g2_msm": 17085, "pairing": 16020}   # reference cost per unit


def cpu_sample_size(workload, threads, seconds):
    """units that take about `seconds` of wall time on `threads` threads at ~45 ns per FpM"""
    if os.environ.get("B200_BENCH_CPU_SAMPLE"):        # tests shrink the sample (tests/test_bench_cli.py)
        return int(os.environ["B200_BENCH_CPU_SAMPLE"])
    n = int(seconds * threads / (CPU_FPM[workload] * 45e-9))
    return max(threads, min(n, 1 << 16))


def cpu_baseline(workload, seed, seconds_all=8.0, seconds_one=2.0):
    """all usable cores and ONE thread (the reference's own execution model), same oracle port, bounded samples"""
    cores = host_cores()
    th = cores["used"]
    n_all = cpu_sample_size(workload, th, seconds_all)
    u, sec = cpu_reference(workload, n_all, seed, th)
    n_one = max(4, cpu_sample_size(workload, 1, seconds_one))
    u1, sec1 = cpu_reference(workload, n_one, seed + 1, 1)
    v, v1 = u / sec, u1 / sec1
    return {"value": v, "unit": UNIT[workload], "cores": th, "kind": "port",
            "sample": "%d-unit slice of the same workload through the reference's constant-time path (oracle C++ port, "
                      "-O3 -march=native, dedicated squaring), %d host threads, %.1f s wall" % (n_all, th, sec),
            "single_thread": {"value": v1, "unit": UNIT[workload], "sample": "%d units, %.1f s wall" % (n_one, sec1),
                              "ns_per_fpm": 1e9 / (v1 * CPU_FPM[workload])},
            "effective_cores": v / v1, "host": cores}


def reference_arm(a, wl, log2n, cfg_name, seed):
    from oracle import pyoracle as orc
    orc.build()
    cores = host_cores()
    th = cores["used"]
    ns = cpu_sample_size(wl, th, 4.0)
    times = []
    for i in range(a.warmup + a.steps):
        units, sec = cpu_reference(wl, ns, seed + i, th)
        if i >= a.warmup:
            times.append(sec)
    tot = sum(times)
    val = ns * len(times) / tot
    n1 = max(4, cpu_sample_size(wl, 1, 2.0))
    u1, s1 = cpu_reference(wl, n1, seed + 1000, 1)
    v1 = u1 / s1
    line = {"impl": "reference", "metric": UNIT[wl].replace("/s", "") + " per second", "value": val, "unit": UNIT[wl],
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot / len(times),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64 limbs (381-bit Montgomery)",
            "data": "synthetic", "config": {"workload": cfg_name, "sample": "%d units per step" % ns},
            "cpu_baseline": {"value": val, "unit": UNIT[wl], "cores": th, "kind": "port",
                             "sample": "%d-unit slice per step of the reference's constant-time path (oracle C++ port; the Rust "
                                       "crate cannot be built in this image), all %d usable host threads" % (ns, th),
                             "single_thread": {"value": v1, "unit": UNIT[wl], "ns_per_fpm": 1e9 / (v1 * CPU_FPM[wl])},
                             "effective_cores": val / v1, "host": cores},
            "e2e": {"value": val, "unit": UNIT[wl], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- B200 arm
def load_traffic():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


class Bench:
    def __init__(self, a):
        import torch
        import bls12_381_b200
        self.torch, self.a = torch, a
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        # the stream is torch's: the library enqueues on it (b200_ctx_create_on_stream) and never destroys it, so events,
        # copies and the library's kernels share one stream whose lifetime torch manages
        self.stream = torch.cuda.Stream(device=self.dev)
        self.eng = bls12_381_b200.Engine(self.local_rank, stream=self.stream.cuda_stream)
        if a.window:
            self.eng.set_msm_window(a.window)
        for kv in a.tune:
            key, val = kv.split("=")
            self.eng.set_tuning(key, int(val))
        if self.world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device=self.dev)
            if self.rank == 0:
                uid.copy_(torch.frombuffer(bytearray(self.eng.comm_unique_id()), dtype=torch.uint8))
            self.dist.broadcast(uid, 0)
            self.eng.comm_init(bytes(uid.cpu().numpy().tobytes()), self.rank, self.world)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)      # > 126 MB L2
        self.peak = self.peak_ms = None
        self.peak_lohi = None
        if self.rank == 0:
            self.peak, self.peak_ms = self.eng.imad_peak(3000)
            self.peak_lohi, _ = self.eng.imad_peak(3000, mode=1)     # the same products as separate IMAD + IMAD.HI.U32
        self.traffic = load_traffic()

    def close(self):
        torch = self.torch
        self.flush = None
        torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            torch.cuda.synchronize()
        self.eng.close()                     # destroys the NCCL communicator, side streams and scratch of the ctx
        if self.dist is not None:
            self.dist.destroy_process_group()
        self.stream = None
        torch.cuda.synchronize()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def gen_points(self, kk, count, sd):
        """[t_i]G on the GPU with the config-1 kernel + batch_normalize (both parity-tested)"""
        torch = self.torch
        from bls12_381_b200 import constants_host as ch
        t = torch.from_numpy(rand_scalars(sd, count).copy()).to(self.dev)
        g = torch.from_numpy(np.tile(ch.generator_projective(kk), (count, 1))).to(self.dev)
        pr = torch.empty_like(g)
        self.eng.mul_batch_dev(kk, g, t, pr, count)
        xy = torch.empty((count, 12 * kk), dtype=torch.int64, device=self.dev)
        inf = torch.empty(count, dtype=torch.uint8, device=self.dev)
        self.eng.batch_normalize_dev(kk, pr, count, xy, inf)
        return xy, inf, pr

    def run(self, wl, log2n, steps, warmup, with_e2e=True, with_cpu=True, cpu_seconds=8.0):
        torch, eng, a, dev = self.torch, self.eng, self.a, self.dev
        rank, world, dist, stream = self.rank, self.world, self.dist, self.stream
        n = 1 << log2n
        seed = SEED_BASE + CONFIG_ID[wl] + (3 if (wl == "g1_msm" and log2n == 24) else 0)   # config_id 5 = G1 MSM 2^24
        k = 2 if wl == "g2_msm" else 1
        AFFW, PROJW = 12 * k, 18 * k
        # --shard auto (default): the device-resident arm shards by WINDOW (every rank holds all points: fastest per shard, 3.05 vs
        # 3.67 ms at 8 GPUs), the end-to-end arm by POINT RANGE (a rank uploads only its slice: 17 MB instead of 135 MB per GPU)
        # (2^24, config 5: point ranges for both arms — a rank then only generates and holds its own 2^21-point slice)
        mode = ("window" if log2n <= 22 else "points") if a.shard == "auto" else a.shard
        mode_e2e = "points" if a.shard == "auto" else a.shard
        from bls12_381_b200.sharding import index_range
        t_gen = time.perf_counter()
        if wl in ("pairing", "g1_mul"):
            lo, hi = index_range(n, rank, world)
            n_local = hi - lo
        else:
            lo, hi = (index_range(n, rank, world) if (world > 1 and mode == "points") else (0, n))
            n_local = n
        if wl == "pairing":
            pxy, pinf, _ = self.gen_points(1, n_local, seed * 7 + rank)
            qxy, qinf, _ = self.gen_points(2, n_local, seed * 11 + rank)
            out = torch.empty((n_local, 72), dtype=torch.int64, device=dev)
        elif wl == "g1_mul":
            xy, inf, pr = self.gen_points(k, n_local, seed * 7 + rank)
            sc = torch.from_numpy(rand_scalars(seed * 13 + rank, n_local).copy()).to(dev)
            out = torch.empty((max(n_local, 1), PROJW), dtype=torch.int64, device=dev)
        else:
            if world > 1 and mode == "points":
                # point-range sharding: a rank only ever reads its slice, so it only generates its slice (the buffers keep
                # the full length: b200_g1_msm_sharded_dev takes the whole arrays)
                xy = torch.zeros((n, AFFW), dtype=torch.int64, device=dev)
                inf = torch.zeros(n, dtype=torch.uint8, device=dev)
                sc = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
                sxy, sinf, _ = self.gen_points(k, hi - lo, seed * 7 + 1000 * (rank + 1))
                xy[lo:hi].copy_(sxy)
                inf[lo:hi].copy_(sinf)
                sc[lo:hi].copy_(torch.from_numpy(rand_scalars(seed * 13 + 1000 * (rank + 1), hi - lo).copy()))
                del sxy, sinf
            else:
                xy, inf, _ = self.gen_points(k, n, seed * 7)
                sc = torch.from_numpy(rand_scalars(seed * 13, n).copy()).to(dev)
            out = torch.empty((1, PROJW), dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        t_gen = time.perf_counter() - t_gen

        def step_device():
            if wl == "g1_mul":
                eng.mul_batch_dev(1, pr, sc, out, n_local)
            elif wl == "pairing":
                eng.pairing_batch_dev(pxy, pinf, qxy, qinf, n_local, out)
            elif world == 1:
                eng.msm_dev(k, xy, inf, sc, n, out)
            else:
                eng.msm_sharded_dev(k, xy, inf, sc, n, out, mode=mode)     # shard + ncclAllGather + combine in the library

        for _ in range(warmup):
            step_device()
        self.barrier()
        verified = None
        if world > 1 and wl in ("g1_msm", "g2_msm") and mode == "window" and log2n <= 22:
            # outside the timed region: the sharded result must be the same group element as the one-GPU MSM
            full = torch.empty_like(out)
            eng.msm_dev(k, xy, inf, sc, n, full)
            both = torch.cat([out, full]).contiguous()
            axy = torch.empty((2, AFFW), dtype=torch.int64, device=dev)
            ainf = torch.empty(2, dtype=torch.uint8, device=dev)
            eng.batch_normalize_dev(k, both, 2, axy, ainf)
            verified = bool(torch.equal(axy[0], axy[1]) and ainf[0] == ainf[1])
            if not verified:
                raise SystemExit("bench.py: sharded MSM result differs from the single-GPU result on rank %d" % rank)
        if world > 1 and wl in ("g1_msm", "g2_msm"):
            # every rank must hold the same combined point
            ref = out.clone()
            dist.broadcast(ref, 0)
            if not torch.equal(ref, out):
                raise SystemExit("bench.py: ranks disagree on the combined MSM result (rank %d)" % rank)
            verified = True if verified is None else verified
        split_checked = None
        if world > 1 and wl in ("g1_msm", "g2_msm") and (log2n > 22 or os.environ.get("B200_BENCH_FORCE_SPLIT_CHECK")):
            # sizes with no single-GPU / oracle comparison (config 5, 2^24): a size-independent property instead — the MSM over all
            # points must equal MSM(first half of every rank's range) + MSM(second half), computed by the same sharded call with the
            # other half flagged as identities (different bucket populations, same group element).  Outside the timed region.
            try:
                r_lo, r_hi = index_range(n, rank, world)
                mid = (r_lo + r_hi) // 2
                parts = torch.empty((2, PROJW), dtype=torch.int64, device=dev)
                for half, (a0, a1) in enumerate(((mid, r_hi), (r_lo, mid))):       # flag [a0, a1) as identities
                    inf_h = inf.clone()
                    inf_h[a0:a1] = 1
                    if mode == "window":                                            # every rank holds all points: flag the same halves everywhere
                        for rr in range(world):
                            q_lo, q_hi = index_range(n, rr, world)
                            q_mid = (q_lo + q_hi) // 2
                            b0, b1 = ((q_mid, q_hi), (q_lo, q_mid))[half]
                            inf_h[b0:b1] = 1
                    tmp = torch.empty((1, PROJW), dtype=torch.int64, device=dev)
                    eng.msm_sharded_dev(k, xy, inf_h, sc, n, tmp, mode=mode)
                    parts[half].copy_(tmp[0])
                total = torch.empty((1, PROJW), dtype=torch.int64, device=dev)
                eng.sum_dev(k, parts, 2, total)
                both = torch.cat([out, total]).contiguous()
                axy = torch.empty((2, AFFW), dtype=torch.int64, device=dev)
                ainf = torch.empty(2, dtype=torch.uint8, device=dev)
                eng.batch_normalize_dev(k, both, 2, axy, ainf)
                torch.cuda.synchronize()
                split_checked = bool(torch.equal(axy[0], axy[1]) and ainf[0] == ainf[1] and int(ainf[0]) == 0)
            except Exception as e:                                                   # a broken CHECK must not hide the measurement
                split_checked = "not run: %s" % (str(e)[:120],)
            if split_checked is False:
                raise SystemExit("bench.py: MSM(all) != MSM(first halves) + MSM(second halves) on rank %d" % rank)
            # re-run the real call so that `out` and the warm state are those of the timed configuration
            step_device()
            self.barrier()
        eng.set_timing(True)
        launches0 = eng.launches
        sampler = ClockSampler(self.local_rank)
        sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            self.flush.zero_()                                           # L2 flush, outside the timed events
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
                torch.cuda.synchronize()
            ev[i][0].record(stream)
            step_device()
            ev[i][1].record(stream)
        self.barrier()
        clocks = sampler.stop()
        launches = eng.launches - launches0
        timing = eng.get_timing()
        eng.set_timing(False)
        total_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
        del ev
        if dist is not None:
            tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            total_ms = float(tt.item())
        ms_per_step = total_ms / steps
        value = n / (ms_per_step * 1e-3)                                 # whole-job units (all ranks together)

        # -------------------------------------------------------------- e2e through the host-pointer C ABI
        e2e = None
        if with_e2e:
            pin = lambda x: torch.empty(x.shape, dtype=x.dtype).pin_memory().copy_(x)
            if wl == "pairing":
                hp = [pin(x) for x in (pxy, pinf, qxy, qinf)]
                h2d = sum(x.numel() * x.element_size() for x in hp) * world
                d2h = n_local * 576 * world

                hgt = torch.empty((n_local, 72), dtype=torch.int64).pin_memory()       # caller-owned pinned result buffer
                hgt_np = hgt.numpy().view(np.uint64)

                def step_host():
                    eng.pairing_batch(hp[0].numpy().view(np.uint64), hp[1].numpy(), hp[2].numpy().view(np.uint64), hp[3].numpy(),
                                      out=hgt_np)
            elif wl == "g1_mul":
                hpr, hsc = pin(pr), pin(sc)
                h2d, d2h = (hpr.numel() * 8 + hsc.numel()) * world, hpr.numel() * 8 * world

                def step_host():
                    eng.mul_batch(1, hpr.numpy().view(np.uint64), hsc.numpy())
            elif world == 1:
                hxy, hinf, hsc = pin(xy), pin(inf), pin(sc)
                h2d, d2h = hxy.numel() * 8 + hinf.numel() + hsc.numel(), PROJW * 8

                def step_host():
                    eng.msm(k, hxy.numpy().view(np.uint64), hinf.numpy(), hsc.numpy())
            else:
                # every rank feeds its own GPU from pinned host memory inside the timed region (point-range sharding: only its
                # slice; window sharding: everything), runs the collective MSM of the library, reads the result back
                if mode_e2e == "points":
                    lo, hi = index_range(n, rank, world)
                hxy, hinf, hsc = pin(xy[lo:hi]), pin(inf[lo:hi]), pin(sc[lo:hi])
                hres = torch.empty((1, PROJW), dtype=torch.int64).pin_memory()
                h2d = (AFFW * 8 + 33) * (n if mode_e2e == "points" else n * world)
                d2h = PROJW * 8 * world

                def step_host():
                    with torch.cuda.stream(stream):
                        xy[lo:hi].copy_(hxy, non_blocking=True)
                        inf[lo:hi].copy_(hinf, non_blocking=True)
                        sc[lo:hi].copy_(hsc, non_blocking=True)
                    eng.msm_sharded_dev(k, xy, inf, sc, n, out, mode=mode_e2e)
                    with torch.cuda.stream(stream):
                        hres.copy_(out, non_blocking=True)
                    stream.synchronize()
            for _ in range(2):
                step_host()
            self.barrier()
            ne = max(3, min(steps, 10))
            t0 = time.perf_counter()
            for _ in range(ne):
                step_host()                                               # synchronous: returns after the D2H copy
            torch.cuda.synchronize()
            e2e_ms = (time.perf_counter() - t0) * 1e3 / ne
            if dist is not None:
                tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                e2e_ms = float(tt.item())
            e2e = {"value": n / (e2e_ms * 1e-3), "unit": UNIT[wl], "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": e2e_ms,
                   "timing": "host wall clock around the synchronous public call (pinned H2D + kernels + D2H)%s, %d steps"
                             % (", max over ranks" if world > 1 else "", ne)}

        # -------------------------------------------------------------- roofline of the dominant kernel
        roof = None
        if rank == 0:
            peak = self.peak
            per_kernel = {}
            for name, ms in timing:
                base = name.split("<")[0].strip("( ")
                if base.startswith("k_msm_accumulate"):
                    base = "k_msm_accumulate"           # G1 / G2-register / G2-shared-memory variants of the bucket kernel
                per_kernel.setdefault(base, []).append(ms)
            dom = DOMINANT[wl]
            if wl == "g1_mul" and dom not in per_kernel:      # forced shapes (--tune mul_groups=6 / -1)
                dom = next((c for c in ("k_mul_batch_warp", "k_mul_batch") if c in per_kernel), dom)
            chunked_v4 = False
            if wl == "pairing" and dom not in per_kernel:
                # one-thread-per-pairing kernels (what the default picks above 28 672 pairs): Miller loop and final exponentiation
                # run as 4 chunks on two streams, so their event times overlap — the roofline is taken over the whole step
                dom, chunked_v4 = "k_final_exp", True
            ksum = {kname: sum(v) for kname, v in per_kernel.items()}
            tot_k = sum(ksum.values()) or 1.0
            if dom in per_kernel:
                avg_ms = sum(per_kernel[dom]) / steps
                if wl in ("g1_msm", "g2_msm"):
                    c = a.window or 16
                    nwin = (256 + c - 1) // c
                    if world > 1 and mode == "window":
                        n_eff, nwin_local = n, len(range(rank, nwin, world))
                    elif world > 1:
                        n_eff, nwin_local = hi - lo, nwin
                    else:
                        n_eff, nwin_local = n, nwin
                    fpm = (11 if k == 1 else 33) * n_eff * nwin_local       # SURVEY 8d model: one complete mixed add per term and window
                    fpm_exec = (10 if k == 1 else 30) * n_eff * nwin_local  # what the kernel executes: XYZZ madd, 8M+2S
                    alg_bytes = (96 * k + 4) * n_eff * nwin_local
                elif wl == "g1_mul":
                    fpm = fpm_exec = 5100.0 * n_local
                    alg_bytes = (144 * 2 + 32) * n_local
                elif dom == "k_coop_pairing":
                    # the six-lane kernels: G2 line coefficients (k_g2_prepare) + Miller loop + final exponentiation together are
                    # the model's 16 020 FpM per pairing, so the three launches are timed together
                    fpm = fpm_exec = 16020.0 * n_local
                    alg_bytes = (96 + 192 + 576) * n_local
                    avg_ms += (sum(per_kernel.get("k_g2_prepare", [])) + sum(per_kernel.get("k_coop_g2_prepare", []))) / steps
                    if avg_ms > 1.02 * ms_per_step:      # chunks on two streams: the launches overlap, their event times add up to
                        avg_ms, chunked_v4 = ms_per_step, True   # more than the step — the three kernels ARE the step: time that
                else:
                    fpm = fpm_exec = 16020.0 * n_local                      # k_miller_loop + k_final_exp together (whole step)
                    alg_bytes = (96 + 192 + 576) * n_local
                    avg_ms = ms_per_step
                achieved = fpm * IMAD_PER_FPM / (avg_ms * 1e-3)
                try:
                    hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
                    hbm_of = "measured"
                except Exception:
                    hbm_peak, hbm_of = 6650.0, "fallback"
                tr = self.traffic.get("%s_2p%d" % (wl, log2n)) if world == 1 else None
                roof = {"bound": "int (IMAD.WIDE.U32 pipe; SURVEY 8d: not hbm, not tensor)",
                        "kernel": ("k_g2_prepare + k_coop_pairing<Miller> + k_coop_pairing<final exp> (chunks on two streams; whole step)"
                                   if chunked_v4 and dom == "k_coop_pairing" else
                                   "k_miller_loop + k_final_exp (4 chunks on two streams; whole step)" if chunked_v4 else dom),
                        "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "T IMAD32/s", "frac": achieved / peak,
                        "executed_frac": fpm_exec * IMAD_PER_FPM / (avg_ms * 1e-3) / peak,
                        "peak_source": "b200_imad_peak microbenchmark run live on this GPU (%.2f ms, dependent-free IMAD.WIDE.U32; "
                                       "SASS + ncu of the microbenchmark: profiles/r02_imad_peak.txt)" % self.peak_ms,
                        "peak_detail": {"imad_wide_u32_per_s": peak, "imad_lo_plus_imad_hi_pairs_per_s": self.peak_lohi,
                                        "per_clock_per_sm_at_sampled_clock": (peak / (148 * clocks["sm_mhz"] * 1e6) if clocks.get("sm_mhz") else None),
                                        "sm_mhz": clocks.get("sm_mhz"),
                                        "note": "one 32x32+64 multiply-add = one IMAD.WIDE.U32 (a carry-linked mad.lo.cc/madc.hi pair is "
                                                "fused into it by ptxas); as separate IMAD + IMAD.HI.U32 it costs two multiplier issues"},
                        "model": "SURVEY 8d cost sheet x 300 IMAD32 per FpM", "kernel_ms_per_step": avg_ms,
                        "kernel_launches_per_step": len(per_kernel[dom]) / steps,
                        # share of the SUMMED kernel time (kernels of an MSM overlap on three streams), not of the step
                        "kernel_share_of_summed_kernel_time": ksum[dom] / tot_k,
                        # dram__bytes_read + dram__bytes_write of this kernel per step, from the ncu --set full capture
                        "traffic": (tr or {}).get("dram_bytes_per_step"), "traffic_source": (tr or {}).get("source"),
                        "hbm": {"achieved_gbs": alg_bytes / (avg_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                "frac": alg_bytes / (avg_ms * 1e-3) / 1e9 / hbm_peak, "of": hbm_of},
                        "kernel_ms": {kname: sum(v) / steps for kname, v in per_kernel.items()}}
                if wl == "g2_msm":
                    roof["frac_note"] = ("the SURVEY 8d model prices an Fp2 multiplication at 3 FpM = 900 multiply-adds (the "
                                         "reference's Karatsuba); the kernel's lazy-reduction Fp2 multiply executes 744, so the "
                                         "model fraction can exceed 1 — executed_frac counts 10 Fp2 products x 744")
                    roof["executed_frac"] = 10 * 744.0 * n_eff * nwin_local / (avg_ms * 1e-3) / peak
                roof["model_frac_whole_step"] = (model_fpm(wl, log2n) * IMAD_PER_FPM * n / world / (ms_per_step * 1e-3) / peak)

        cpu = None
        if rank == 0 and world == 1 and with_cpu:
            cpu = cpu_baseline(wl, seed, seconds_all=cpu_seconds)
        cfg_name = {"g1_mul": "%d x G1Projective scalar-mul batch", "g1_msm": "G1 Pippenger MSM, 2^%d random points x random Scalars",
                    "g2_msm": "G2 Pippenger MSM, 2^%d random points x random Scalars",
                    "pairing": "batched pairing: 2^%d (G1Affine,G2Affine) pairs -> Gt"}[wl] % (n if wl == "g1_mul" else log2n)
        res = {"value": value, "unit": UNIT[wl], "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
               "config": {"workload": cfg_name, "n": n,
                          "sharding": ("none" if world == 1 else
                                       ("by pair/item index, no collective" if wl in ("pairing", "g1_mul") else
                                        mode + "-sharded (device-resident arm) / " + mode_e2e + "-sharded (e2e arm) inside the library: "
                                        "shard + one ncclAllGather of the partial sums + combine on one stream")),
                          "l2": "256 MiB buffer written between timed steps (L2 flush)",
                          "input_generation_s": t_gen, "seed": hex(seed), "sharded_result_checked": verified,
                          "split_sum_checked": split_checked},
               "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu}
        return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="all", choices=["all"] + list(CONFIG_ID))
    ap.add_argument("--log2n", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shard", default="auto", choices=["auto", "window", "points"])
    ap.add_argument("--window", type=int, default=0, help="MSM window bits (0 = auto)")
    ap.add_argument("--tune", action="append", default=[], help="key=value tuning knob (b200_ctx_set_tuning)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    everything = a.workload == "all"
    wl = "g1_msm" if everything else a.workload
    log2n = a.log2n if a.log2n is not None else DEFAULT_LOG2N[wl]
    rank = int(os.environ.get("RANK", "0"))
    seed = SEED_BASE + CONFIG_ID[wl]
    cfg_name = {"g1_mul": "%d x G1Projective scalar-mul batch", "g1_msm": "G1 Pippenger MSM, 2^%d random points x random Scalars",
                "g2_msm": "G2 Pippenger MSM, 2^%d random points x random Scalars",
                "pairing": "batched pairing: 2^%d (G1Affine,G2Affine) pairs -> Gt"}[wl] % ((1 << log2n) if wl == "g1_mul" else log2n)

    if a.impl == "reference":
        if rank == 0:
            reference_arm(a, wl, log2n, cfg_name, seed)
        return

    b = Bench(a)
    try:
        head = b.run(wl, log2n, a.steps, a.warmup, with_e2e=not a.no_e2e, with_cpu=not a.no_cpu_baseline)
        configs = {}
        if everything:
            ks = max(3, min(a.steps, 5))
            for name, w2, l2 in (("g1_mul_1024", "g1_mul", 10), ("g2_msm_2p20", "g2_msm", 20), ("pairing_2p16", "pairing", 16)):
                configs[name] = b.run(w2, l2, ks, 3, with_e2e=not a.no_e2e, with_cpu=not a.no_cpu_baseline, cpu_seconds=3.0)
            if b.world == 8:
                configs["g1_msm_2p24"] = b.run("g1_msm", 24, 3, 3, with_e2e=False, with_cpu=False)
        if b.rank == 0:
            line = {"metric": UNIT[wl].replace("/s", "") + " per second", "value": head["value"], "unit": UNIT[wl],
                    "n_gpus": b.world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": head["ms_per_step"],
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": "u32x12 limbs (381-bit Montgomery, integer)", "data": "synthetic", "config": head["config"],
                    "gpu_launches": head["gpu_launches"] + sum(c["gpu_launches"] for c in configs.values()),
                    "gpu_launches_headline": head["gpu_launches"], "clocks": head["clocks"], "e2e": head["e2e"],
                    "roofline": head["roofline"], "cpu_baseline": head["cpu_baseline"]}
            if configs:
                line["configs"] = configs
            print(json.dumps(line))
            sys.stdout.flush()
    finally:
        b.close()


if __name__ == "__main__":
    main()
