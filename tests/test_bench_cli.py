"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the CPU path of the
reference (oracle port), prints ONE JSON line with the keys the driver reads, and non-zero ranks stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra=None):
    env = dict(os.environ, B200_BENCH_CPU_SAMPLE="192")
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=env,
                          timeout=600)


def test_reference_arm_json_contract(orc):
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    assert d["unit"] == "G1 MSM point-scalar-muls/s" and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_are_silent(orc):
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2"], {"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_pairing_workload(orc):
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "pairing"], {"B200_BENCH_CPU_SAMPLE": "16"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["unit"] == "pairings/s" and d["value"] > 0
