"""GPU parity tests (-m gpu) for the scalar-field row: b200_fr_op / to_bytes / from_bytes /
b200_fr_ntt through the C ABI against the oracle, bit-exact, plus size-independent properties at large n.

STATUS: hardware-validated (round 2, first GPU call: all tests passed on a B200); the same device source is also
covered on the CPU harness (tests/test_device_fr_cpu.py)."""
import numpy as np
import pytest

from tests.test_oracle_fr import Q, to_mont, raw, L

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.fixture(scope="module")
def eng():
    import bls12_381_b200
    e = bls12_381_b200.Engine()
    yield e
    e.close()


def rand_fr(rng, n):
    """n canonical elements: uniformly random 252-bit limb patterns (< q), cheap to make for large n"""
    return np.ascontiguousarray(np.frombuffer(rng.bytes(32 * n), np.uint64).reshape(n, 4) & np.uint64(0x0fffffffffffffff))


def edge_fr():
    vals = [0, 1, 2, Q - 1, Q - 2, (Q - 1) // 2, (Q + 1) // 2, 1 << 254, (1 << 32) - 1, 1 << 32, Q - (1 << 32)]
    return np.concatenate([raw(v) for v in vals])


@pytest.mark.parametrize("op", ["mul", "add", "sub", "square", "neg", "double", "invert"])
def test_fr_ops(eng, orc, op):
    rng = np.random.default_rng(7100)
    e = edge_fr()
    a = np.concatenate([rand_fr(rng, 4096), np.repeat(e, len(e), 0)])
    b = np.concatenate([rand_fr(rng, 4096), np.tile(e, (len(e), 1))])
    if op == "invert":
        a = np.ascontiguousarray(a[-400:])
    bb = b if op in ("mul", "add", "sub") else None
    assert np.array_equal(eng.fr_op(op, a, bb), orc.fr_op(op, a, bb, threads=8))


def test_fr_kats_on_gpu(eng):
    """the literals of src/scalar.rs's own tests, straight on the GPU (:1058-1105, :863-968)"""
    big = L("const_LARGEST")
    assert np.array_equal(eng.fr_op("add", big, big), L("test_addition", 0))
    assert np.array_equal(eng.fr_op("add", big, raw(1)), raw(0))
    assert np.array_equal(eng.fr_op("neg", big), raw(1))
    assert np.array_equal(eng.fr_op("sub", raw(0), big), raw(1))
    one = L("const_R")
    assert np.array_equal(eng.fr_op("mul", L("const_ROOT_OF_UNITY"), L("const_ROOT_OF_UNITY_INV")), one)
    assert np.array_equal(eng.fr_op("mul", to_mont(2), L("const_TWO_INV")), one)
    assert np.array_equal(eng.fr_op("invert", L("const_ROOT_OF_UNITY")), L("const_ROOT_OF_UNITY_INV"))
    import json, os
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))
    enc = np.array(kat["scalar.rs::test_to_bytes_bytes"], np.uint8)
    vals = np.concatenate([raw(0), one, L("const_R2"), eng.fr_op("neg", one)])
    assert np.array_equal(eng.fr_to_bytes(vals), enc)
    dec = np.array(kat["scalar.rs::test_from_bytes_bytes"], np.uint8)
    got, ok = eng.fr_from_bytes(dec)
    assert list(ok) == [1, 1, 1, 1, 0, 0, 0, 0]
    assert np.array_equal(got[:4], vals) and not got[4:].any()


def test_fr_bytes_roundtrip_and_msm_scalars(eng, orc):
    rng = np.random.default_rng(7200)
    a = rand_fr(rng, 5000)
    by = eng.fr_to_bytes(a)
    assert np.array_equal(by, orc.scalar_to_bytes(a))
    back, ok = eng.fr_from_bytes(by)
    assert ok.all() and np.array_equal(back, a)
    # Montgomery scalars -> to_bytes -> MSM: the caller-side step in front of every MSM
    from tests import util
    _, xy, inf = util.rand_points(orc, 1, rng, 64)
    got = eng.msm(1, xy, inf, by[:64])
    assert np.array_equal(orc.G1.to_affine(got)[0], orc.G1.to_affine(orc.G1.msm_naive(xy, inf, by[:64], threads=8))[0])


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 6, 7, 9, 12, 13, 16])
def test_ntt_parity(eng, orc, log_n):
    rng = np.random.default_rng(7300 + log_n)
    a = rand_fr(rng, 1 << log_n)
    for inverse in (False, True):
        for coset in (False, True):
            got = eng.fr_ntt(a, inverse=inverse, coset=coset)
            assert np.array_equal(got, orc.fr_ntt(a, inverse=inverse, coset=coset, threads=8)), (inverse, coset)
    if log_n <= 6:
        assert np.array_equal(eng.fr_ntt(a), orc.fr_dft_naive(a))   # the O(n^2) definition


def test_ntt_table_cache_switches_sizes(eng, orc):
    rng = np.random.default_rng(7400)
    for log_n in (10, 4, 10, 11, 4):
        a = rand_fr(rng, 1 << log_n)
        assert np.array_equal(eng.fr_ntt(a, coset=True), orc.fr_ntt(a, coset=True, threads=8))


def test_ntt_large_properties(eng, orc):
    """2^20: round trip, linearity, and a sample of outputs against the definition evaluated by Horner"""
    rng = np.random.default_rng(7500)
    log_n = 20
    n = 1 << log_n
    a, b = rand_fr(rng, n), rand_fr(rng, n)
    fa = eng.fr_ntt(a)
    for coset in (False, True):
        f = eng.fr_ntt(a, coset=coset)
        assert np.array_equal(eng.fr_ntt(f, inverse=True, coset=coset), a)
    assert np.array_equal(eng.fr_ntt(eng.fr_op("add", a, b)), eng.fr_op("add", fa, eng.fr_ntt(b)))
    assert np.array_equal(fa, orc.fr_ntt(a, threads=8))
    # out[k] = sum_j a[j] w^(jk) for a few k, in Python integers
    from tests.test_oracle_fr import from_mont
    w = pow(from_mont(L("const_ROOT_OF_UNITY")), 1 << (32 - log_n), Q)
    R = (1 << 256) % Q
    rinv = pow(R, -1, Q)
    ints = [sum(int(x) << (64 * i) for i, x in enumerate(row)) * rinv % Q for row in a[:4096]]
    # a sparse check: the polynomial with only the first 4096 coefficients
    sp = np.zeros_like(a)
    sp[:4096] = a[:4096]
    fs = eng.fr_ntt(sp)
    for k in (0, 1, 12345, n - 1):
        x = pow(w, k, Q)
        acc = 0
        for c in reversed(ints):
            acc = (acc * x + c) % Q
        assert from_mont(fs[k]) == acc


def test_ntt_dev_in_place(eng, orc):
    import torch
    rng = np.random.default_rng(7600)
    log_n = 14
    a = rand_fr(rng, 1 << log_n)
    t = torch.from_numpy(a.view(np.int64)).cuda()
    out = torch.empty_like(t)
    torch.cuda.synchronize()              # the engine works on its own non-blocking stream
    eng.fr_ntt_dev(t, log_n, out)
    eng.fr_ntt_dev(t, log_n, t)           # in == out
    torch.cuda.synchronize()
    want = orc.fr_ntt(a, threads=8)
    assert np.array_equal(out.cpu().numpy().view(np.uint64), want)
    assert np.array_equal(t.cpu().numpy().view(np.uint64), want)
    prod = torch.empty_like(t)
    eng.fr_op_dev("mul", t, out, 1 << log_n, prod)
    torch.cuda.synchronize()
    assert np.array_equal(prod.cpu().numpy().view(np.uint64), orc.fr_op("mul", want, want, threads=8))


def test_fr_argument_errors(eng):
    import bls12_381_b200
    with pytest.raises(ValueError):
        eng.fr_ntt(np.zeros((3, 4), np.uint64))
    a = np.zeros((4, 4), np.uint64)
    out = np.empty_like(a)
    rc = eng.lib.b200_fr_ntt(eng.h, a.ctypes.data, 29, 0, 0, out.ctypes.data)
    assert rc == -1
    rc = eng.lib.b200_fr_op(eng.h, 6, a.ctypes.data, a.ctypes.data, 4, out.ctypes.data)   # frobenius: not an Fr op
    assert rc == -1
