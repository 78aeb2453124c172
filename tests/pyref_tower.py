"""Independent Python big-integer model of the BLS12-381 extension tower (no Montgomery form, plain schoolbook
formulas — nothing shared with oracle/ or the CUDA code):  Fp2 = Fp[u]/(u^2+1),  Fp6 = Fp2[v]/(v^3-(u+1)),
Fp12 = Fp6[w]/(w^2-v).  Elements are nested tuples of ints.  Used to cross-check the oracle's tower arithmetic and
the exponent of its final exponentiation (f -> f^(3(p^12-1)/r))."""
from tests.pyref import P, Q, from_mont, to_mont
import numpy as np

XI = (1, 1)  # u + 1


def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
F2_ZERO, F2_ONE = (0, 0), (1, 0)


def f6_add(a, b): return tuple(f2_add(x, y) for x, y in zip(a, b))
def f6_sub(a, b): return tuple(f2_sub(x, y) for x, y in zip(a, b))


def f6_mul(a, b):
    # schoolbook in v with v^3 = xi
    c = [F2_ZERO] * 5
    for i in range(3):
        for j in range(3):
            c[i + j] = f2_add(c[i + j], f2_mul(a[i], b[j]))
    return (f2_add(c[0], f2_mul(XI, c[3])), f2_add(c[1], f2_mul(XI, c[4])), c[2])


def f6_mul_by_v(a):  # multiply by v
    return (f2_mul(XI, a[2]), a[0], a[1])


F6_ZERO = (F2_ZERO, F2_ZERO, F2_ZERO)
F6_ONE = (F2_ONE, F2_ZERO, F2_ZERO)


def f12_mul(a, b):
    # (a0 + a1 w)(b0 + b1 w) = a0 b0 + v a1 b1 + (a0 b1 + a1 b0) w
    return (f6_add(f6_mul(a[0], b[0]), f6_mul_by_v(f6_mul(a[1], b[1]))), f6_add(f6_mul(a[0], b[1]), f6_mul(a[1], b[0])))


F12_ONE = (F6_ONE, F6_ZERO)


def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_mul(a, a)
        e >>= 1
    return r


def f12_from_limbs(l):
    l = np.asarray(l, dtype=np.uint64).reshape(12, 6)
    v = [from_mont(x) for x in l]
    f2 = [(v[2 * i], v[2 * i + 1]) for i in range(6)]
    return ((f2[0], f2[1], f2[2]), (f2[3], f2[4], f2[5]))


def f12_to_limbs(a):
    out = []
    for c6 in a:
        for c2 in c6:
            out.append(to_mont(c2[0]))
            out.append(to_mont(c2[1]))
    return np.concatenate(out)


def f6_from_limbs(l):
    l = np.asarray(l, dtype=np.uint64).reshape(6, 6)
    v = [from_mont(x) for x in l]
    return tuple((v[2 * i], v[2 * i + 1]) for i in range(3))


def f6_to_limbs(a):
    return np.concatenate([np.concatenate([to_mont(c[0]), to_mont(c[1])]) for c in a])


FINAL_EXP = 3 * (P ** 12 - 1) // Q   # what src/pairings.rs:134-176 computes (the cube of the textbook exponent)
assert (P ** 12 - 1) % Q == 0
