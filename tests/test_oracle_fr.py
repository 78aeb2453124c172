"""The oracle's scalar-field (Fr) arithmetic and NTT, pinned by the reference's own tests in src/scalar.rs (replayed
on the oracle with the literals extracted into tests/golden/kat.json) and by Python big integers (independent)."""
import json
import os

import numpy as np
import pytest

from tests import pyref

Q = pyref.Q
R = (1 << 256) % Q
KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


def L(key, i=None):
    v = KAT["scalar.rs::" + key]
    v = v if i is None else v[i]
    return np.array([int(x, 16) for x in v], dtype=np.uint64).reshape(1, 4)


def raw(v):
    return np.array([(v >> (64 * i)) & (2 ** 64 - 1) for i in range(4)], dtype=np.uint64).reshape(1, 4)


def to_mont(v):
    return raw(v * R % Q)


def from_mont(l):
    return sum(int(x) << (64 * i) for i, x in enumerate(np.asarray(l).reshape(-1))) * pow(R, -1, Q) % Q


def eq(a, b):
    return np.array_equal(np.asarray(a, np.uint64).reshape(-1), np.asarray(b, np.uint64).reshape(-1))


def test_constants(orc):  # src/scalar.rs:787-817 and the constant definitions :76-222
    assert sum(int(x) << (64 * i) for i, x in enumerate(L("const_MODULUS")[0])) == Q
    assert eq(orc.fr_const("one"), L("const_R")) and from_mont(L("const_R")) == 1
    assert eq(orc.fr_const("two_inv"), L("const_TWO_INV"))
    assert eq(orc.fr_const("root_of_unity"), L("const_ROOT_OF_UNITY"))
    assert eq(orc.fr_const("root_of_unity_inv"), L("const_ROOT_OF_UNITY_INV"))
    assert eq(orc.fr_const("generator"), L("const_GENERATOR")) and from_mont(L("const_GENERATOR")) == 7
    one = orc.fr_const("one")
    assert eq(orc.fr_op("mul", to_mont(2), L("const_TWO_INV")), one)
    assert eq(orc.fr_op("mul", L("const_ROOT_OF_UNITY"), L("const_ROOT_OF_UNITY_INV")), one)
    assert eq(orc.fr_pow(L("const_ROOT_OF_UNITY"), raw(1 << 32)), one)
    assert not eq(orc.fr_pow(L("const_ROOT_OF_UNITY"), raw(1 << 31)), one)
    t = (Q - 1) >> 32
    assert eq(orc.fr_pow(L("const_DELTA"), raw(t)), one)
    # ROOT_OF_UNITY = GENERATOR^t (doc comment :193-199)
    assert eq(orc.fr_pow(L("const_GENERATOR"), raw(t)), L("const_ROOT_OF_UNITY"))


def test_addition_negation_subtraction(orc):  # src/scalar.rs:1058-1105
    big = L("const_LARGEST")
    assert eq(orc.fr_op("add", big, big), L("test_addition", 0))
    assert eq(orc.fr_op("add", big, raw(1)), raw(0))
    assert eq(orc.fr_op("neg", big), raw(1))
    assert eq(orc.fr_op("neg", raw(0)), raw(0))
    assert eq(orc.fr_op("neg", raw(1)), big)
    assert eq(orc.fr_op("sub", big, big), raw(0))
    assert eq(orc.fr_op("sub", raw(0), big), orc.fr_op("sub", L("const_MODULUS"), big))
    a = orc.fr_op("mul", L("test_double", 0), L("const_R2"))  # from_raw (:335-337)
    assert eq(orc.fr_op("double", a), orc.fr_op("add", a, a))   # :1256-1266


def test_multiplication_squaring_by_double_and_add(orc):  # src/scalar.rs:1107-1163
    cur = L("const_LARGEST")
    for _ in range(100):
        prod = orc.fr_op("mul", cur, cur)
        assert eq(orc.fr_op("square", cur), prod)
        bits = int.from_bytes(orc.scalar_to_bytes(cur)[0].tobytes(), "little")
        acc = raw(0)
        for i in range(255, -1, -1):
            acc = orc.fr_op("add", acc, acc)
            if (bits >> i) & 1:
                acc = orc.fr_op("add", acc, cur)
        assert eq(acc, prod)
        cur = orc.fr_op("add", cur, L("const_LARGEST"))


def test_inversion(orc):  # src/scalar.rs:1165-1208
    one = orc.fr_const("one")
    assert eq(orc.fr_op("invert", raw(0)), raw(0))
    assert eq(orc.fr_op("invert", one), one)
    m1 = orc.fr_op("neg", one)
    assert eq(orc.fr_op("invert", m1), m1)
    tmp = L("const_R2")
    for _ in range(100):
        assert eq(orc.fr_op("mul", orc.fr_op("invert", tmp), tmp), one)
        assert from_mont(orc.fr_op("invert", tmp)) == pow(from_mont(tmp), -1, Q)
        tmp = orc.fr_op("add", tmp, L("const_R2"))


def test_to_from_bytes(orc):  # src/scalar.rs:863-968, :1238-1254
    enc = [np.array(b, np.uint8) for b in KAT["scalar.rs::test_to_bytes_bytes"]]
    one = orc.fr_const("one")
    vals = [raw(0), one, L("const_R2"), orc.fr_op("neg", one)]
    for v, e in zip(vals, enc):
        assert np.array_equal(orc.scalar_to_bytes(v)[0], e)
    dec = [np.array(b, np.uint8) for b in KAT["scalar.rs::test_from_bytes_bytes"]]
    want_ok = [1, 1, 1, 1, 0, 0, 0, 0]
    got, ok = orc.fr_from_bytes(np.stack(dec))
    assert list(ok) == want_ok
    for g, v in zip(got[:3], vals[:3]):
        assert eq(g, v)
    assert eq(got[3], vals[3])
    # from_raw (:335): val * R2; from_raw(MODULUS) == 0; from_raw([1,0,0,0]) == R; from_raw(2^256 - 1) == from_raw(...)
    r2 = L("const_R2")
    assert eq(orc.fr_op("mul", L("const_MODULUS"), r2), raw(0))
    assert eq(orc.fr_op("mul", raw(1), r2), one)
    assert from_mont(orc.fr_op("mul", raw(2 ** 256 - 1), r2)) == (2 ** 256 - 1) % Q


def test_random_against_python(orc):
    rng = np.random.default_rng(5100)
    va = [int.from_bytes(rng.bytes(40), "little") % Q for _ in range(64)] + [0, 1, Q - 1]
    vb = [int.from_bytes(rng.bytes(40), "little") % Q for _ in range(64)] + [Q - 1, Q - 1, Q - 1]
    a, b = np.concatenate([to_mont(v) for v in va]), np.concatenate([to_mont(v) for v in vb])
    for name, f in (("mul", lambda x, y: x * y), ("add", lambda x, y: x + y), ("sub", lambda x, y: x - y)):
        got = orc.fr_op(name, a, b)
        assert [from_mont(g) for g in got] == [f(x, y) % Q for x, y in zip(va, vb)]
    assert [from_mont(g) for g in orc.fr_op("neg", a)] == [(-x) % Q for x in va]
    assert [from_mont(g) for g in orc.fr_op("square", a)] == [x * x % Q for x in va]


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5])
def test_ntt_definition(orc, log_n):
    """orc.fr_dft_naive restates out[k] = sum_j a[j] w^(jk); checked against Python integers, then the O(n log n)
    transform against it (forward, inverse, coset)"""
    n = 1 << log_n
    rng = np.random.default_rng(5200 + log_n)
    vals = [int.from_bytes(rng.bytes(40), "little") % Q for _ in range(n)]
    a = np.concatenate([to_mont(v) for v in vals])
    w = pow(from_mont(L("const_ROOT_OF_UNITY")), 1 << (32 - log_n), Q)
    want = [sum(vals[j] * pow(w, j * k, Q) for j in range(n)) % Q for k in range(n)]
    assert [from_mont(x) for x in orc.fr_dft_naive(a)] == want
    assert [from_mont(x) for x in orc.fr_ntt(a)] == want
    wantc = [sum(vals[j] * pow(7, j, Q) * pow(w, j * k, Q) for j in range(n)) % Q for k in range(n)]
    assert [from_mont(x) for x in orc.fr_ntt(a, coset=True)] == wantc
    assert [from_mont(x) for x in orc.fr_dft_naive(a, coset=True)] == wantc
    for coset in (False, True):
        assert eq(orc.fr_ntt(orc.fr_ntt(a, coset=coset), inverse=True, coset=coset), a)
        assert eq(orc.fr_dft_naive(a, inverse=True, coset=coset), orc.fr_ntt(a, inverse=True, coset=coset))


def test_ntt_fast_vs_naive_medium(orc):
    rng = np.random.default_rng(5300)
    n = 1 << 9
    a = np.concatenate([to_mont(int.from_bytes(rng.bytes(40), "little") % Q) for _ in range(n)])
    for inverse in (False, True):
        for coset in (False, True):
            assert eq(orc.fr_ntt(a, inverse=inverse, coset=coset, threads=4), orc.fr_dft_naive(a, inverse=inverse, coset=coset))
    # convolution theorem on a larger size: NTT(a) * NTT(b) == NTT(a (*) b) for polynomials of degree < n/2
    n = 1 << 12
    va = [int.from_bytes(rng.bytes(40), "little") % Q for _ in range(8)]
    vb = [int.from_bytes(rng.bytes(40), "little") % Q for _ in range(8)]
    a = np.zeros((n, 4), np.uint64)
    b = np.zeros((n, 4), np.uint64)
    a[:8] = np.concatenate([to_mont(v) for v in va])
    b[n // 2 - 8:n // 2] = np.concatenate([to_mont(v) for v in vb])
    prod = orc.fr_ntt(orc.fr_op("mul", orc.fr_ntt(a, threads=4), orc.fr_ntt(b, threads=4)), inverse=True, threads=4)
    c = [0] * n
    for i, x in enumerate(va):
        for j, y in enumerate(vb):
            c[i + n // 2 - 8 + j] = (c[i + n // 2 - 8 + j] + x * y) % Q
    assert [from_mont(x) for x in prod] == c
