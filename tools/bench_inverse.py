"""Times Fermat vs binary-GCD inversion on the GPU (kernel time via the ctx's per-launch events)."""
import numpy as np
import bls12_381_b200 as b
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
eng = b.Engine()
rng = np.random.default_rng(1)
n = 1 << 18
a = rng.integers(0, 1 << 62, (n, 6), dtype=np.uint64)
a[:, 5] &= (1 << 60) - 1                                # < p
for op in ("invert", "invert_fast", "invert", "invert_fast", "mul"):
    eng.set_timing(True)
    eng.tower(1, op, a, a if op == "mul" else None)
    t = eng.get_timing()
    print(op, ["%.3f ms" % ms for _, ms in t], "-> %.1f ns/elem" % (1e6 * t[0][1] / n))
