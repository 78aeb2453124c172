"""Scalar-field kernels and the NTT launch sequence of bls12_381_b200/csrc/fr_ntt.cuh run on the HOST (tests/emul/: same
kernel source, every (block, thread) executed in a loop, bit-exact models of the PTX carry instructions) against the
oracle.  The real-GPU counterpart is tests/test_gpu_zz_fr.py."""
import ctypes as C

import numpy as np
import pytest

from tests.emul import build as emul_build
from tests.test_oracle_fr import Q, to_mont, raw, L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def em():
    return C.CDLL(emul_build.build("default"))


def rand_fr(rng, n):
    return np.concatenate([to_mont(int.from_bytes(rng.bytes(40), "little") % Q) for _ in range(n)])


def edge_fr():
    vals = [0, 1, 2, Q - 1, Q - 2, (Q - 1) // 2, (Q + 1) // 2, 1 << 254, (1 << 32) - 1, 1 << 32, Q - (1 << 32)]
    return np.concatenate([raw(v) for v in vals])   # as limb patterns (all < q), Montgomery or not


def test_fr_ops(em, orc):
    rng = np.random.default_rng(6100)
    a = np.concatenate([rand_fr(rng, 300), edge_fr(), edge_fr()[::-1]])
    b = np.concatenate([rand_fr(rng, 300), edge_fr(), edge_fr()])
    for name, code in (("mul", 0), ("add", 1), ("sub", 2), ("square", 3), ("neg", 4), ("double", 11)):
        out = np.empty_like(a)
        bb = b if name in ("mul", "add", "sub") else None
        assert em.emul_fr_op(code, _p(a), _p(bb), _p(out), C.c_size_t(a.shape[0])) == 0
        assert np.array_equal(out, orc.fr_op(name, a, bb)), name
    sub = np.ascontiguousarray(np.concatenate([a[:40], edge_fr()]))
    out = np.empty_like(sub)
    em.emul_fr_op(5, _p(sub), None, _p(out), C.c_size_t(sub.shape[0]))
    assert np.array_equal(out, orc.fr_op("invert", sub))
    # to_bytes / from_bytes
    tb = np.empty((a.shape[0], 32), np.uint8)
    em.emul_fr_to_bytes(_p(a), _p(tb), C.c_size_t(a.shape[0]))
    assert np.array_equal(tb, orc.scalar_to_bytes(a))
    enc = np.concatenate([tb[:50], np.full((1, 32), 0xff, np.uint8),
                          np.frombuffer(Q.to_bytes(32, "little"), np.uint8).reshape(1, 32),
                          np.frombuffer((Q + 1).to_bytes(32, "little"), np.uint8).reshape(1, 32),
                          np.frombuffer((Q - 1).to_bytes(32, "little"), np.uint8).reshape(1, 32)])
    enc = np.ascontiguousarray(enc)
    back = np.empty((enc.shape[0], 4), np.uint64)
    ok = np.empty(enc.shape[0], np.uint8)
    em.emul_fr_from_bytes(_p(enc), _p(back), _p(ok), C.c_size_t(enc.shape[0]))
    want, wok = orc.fr_from_bytes(enc)
    assert np.array_equal(ok, wok) and list(ok[50:]) == [0, 0, 0, 1]
    assert np.array_equal(back[ok == 1], want[wok == 1])
    assert not back[ok == 0].any()
    assert np.array_equal(back[:50], a[:50])


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 6, 7, 10, 13])
def test_ntt_launch_sequence(em, orc, log_n):
    n = 1 << log_n
    rng = np.random.default_rng(6200 + log_n)
    a = rand_fr(rng, n) if n <= 64 else np.ascontiguousarray(
        np.frombuffer(rng.bytes(32 * n), np.uint64).reshape(n, 4) & np.uint64(0x0fffffffffffffff))
    # (the large case uses random limb patterns < 2^252 < q: valid canonical elements)
    for inverse in (0, 1):
        for coset in (0, 1):
            out = np.empty_like(a)
            rc = em.emul_fr_ntt(_p(a), log_n, inverse, coset, _p(out), 8)
            assert rc == max(1, (log_n + 2) // 3)
            assert np.array_equal(out, orc.fr_ntt(a, inverse=bool(inverse), coset=bool(coset), threads=8)), (inverse, coset)
    if n <= 64:
        out = np.empty_like(a)
        em.emul_fr_ntt(_p(a), log_n, 0, 0, _p(out), 1)
        assert np.array_equal(out, orc.fr_dft_naive(a))
