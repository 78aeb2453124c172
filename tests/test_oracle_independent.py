"""Independent (Python big-integer, textbook formulas) checks of the oracle's extension tower and of the EXPONENT
of its final exponentiation — the oracle restates the reference's algorithms; this file shares nothing with them."""
import numpy as np

from tests import pyref, pyref_tower as T, util


def eq(a, b):
    return np.array_equal(np.asarray(a, np.uint64).reshape(-1), np.asarray(b, np.uint64).reshape(-1))


def test_fp6_fp12_mul_against_python_tower(orc):
    rng = np.random.default_rng(3100)
    a6, b6 = util.rand_fp(rng, 4, 6), util.rand_fp(rng, 4, 6)
    got = orc.tower(6, "mul", a6, b6)
    for i in range(4):
        assert eq(got[i], T.f6_to_limbs(T.f6_mul(T.f6_from_limbs(a6[i]), T.f6_from_limbs(b6[i]))))
    a12, b12 = util.rand_fp(rng, 3, 12), util.rand_fp(rng, 3, 12)
    got = orc.tower(12, "mul", a12, b12)
    sq = orc.tower(12, "square", a12)
    for i in range(3):
        x, y = T.f12_from_limbs(a12[i]), T.f12_from_limbs(b12[i])
        assert eq(got[i], T.f12_to_limbs(T.f12_mul(x, y)))
        assert eq(sq[i], T.f12_to_limbs(T.f12_mul(x, x)))
    # frobenius == x -> x^p on Fp12 (src/fp12.rs:145-171): check on one element by plain exponentiation
    x = T.f12_from_limbs(a12[0])
    assert eq(orc.tower(12, "frobenius", a12[:1]), T.f12_to_limbs(T.f12_pow(x, pyref.P)))
    # invert
    inv = orc.tower(12, "invert", a12[:1])
    assert T.f12_mul(T.f12_from_limbs(inv[0]), x) == T.F12_ONE


def test_final_exponentiation_exponent(orc):
    """MillerLoopResult::final_exponentiation == f^(3 (p^12 - 1) / r)  (src/pairings.rs:134-176), on the
    Miller-loop value of a random pair, by plain square-and-multiply in the Python tower"""
    rng = np.random.default_rng(3200)
    _, pxy, pinf = util.rand_points(orc, 1, rng, 1, threads=1)
    _, qxy, qinf = util.rand_points(orc, 2, rng, 1, threads=1)
    ml = orc.miller_loop(pxy, pinf, qxy, qinf)
    fe = orc.final_exponentiation(ml)
    f = T.f12_from_limbs(ml[0])
    assert eq(fe[0], T.f12_to_limbs(T.f12_pow(f, T.FINAL_EXP)))
    # and the result has order dividing r (it lies in Gt): Gt^r == 1
    assert T.f12_pow(T.f12_from_limbs(fe[0]), pyref.Q) == T.F12_ONE
