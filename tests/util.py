"""Shared input generators for the parity tests (seeded; built from the CPU oracle + Python bigints)."""
import numpy as np

from tests import pyref


def rand_fp(rng, n, width=1):
    """n x (6*width) uint64 limbs, each Fp uniformly random, Montgomery form, canonical."""
    out = np.empty((n, 6 * width), np.uint64)
    for i in range(n):
        for k in range(width):
            out[i, 6 * k:6 * k + 6] = pyref.int_to_limbs(int.from_bytes(rng.bytes(48), "little") % pyref.P)
    return out


def edge_fp():
    vals = [0, 1, 2, pyref.P - 1, pyref.P - 2, pyref.R, (pyref.P - pyref.R) % pyref.P, (1 << 380), (pyref.P + 1) // 2]
    return np.stack([pyref.int_to_limbs(v) for v in vals])


def rand_scalars(rng, n):
    """canonical 32-byte LE scalars < q, like Scalar::random -> from_bytes_wide -> to_bytes"""
    out = np.empty((n, 32), np.uint8)
    for i in range(n):
        v = int.from_bytes(rng.bytes(64), "little") % pyref.Q
        out[i] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
    return out


def scalar_bytes(v):
    return np.frombuffer(int(v % pyref.Q).to_bytes(32, "little"), np.uint8).copy()


def rand_points(orc, k, rng, n, threads=8):
    """n random subgroup points [t_i]G as (projective, affine xy, inf)"""
    G = orc.G1 if k == 1 else orc.G2
    t = rand_scalars(rng, n)
    pr = G.mul(np.repeat(G.generator(), n, 0), t, threads=threads)
    xy, inf = G.batch_normalize(pr)
    return pr, xy, inf


def randomize_z(orc, k, rng, pr):
    """same points, random non-unit z: (x*l, y*l, z*l)"""
    n = pr.shape[0]
    lam = rand_fp(rng, n, k)
    level = 1 if k == 1 else 2
    w = 6 * k
    out = pr.copy()
    for c in range(3):
        out[:, c * w:(c + 1) * w] = orc.tower(level, "mul", pr[:, c * w:(c + 1) * w], lam)
    return out
