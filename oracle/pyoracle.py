"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes/numpy wrapper over oracle/liboracle.so (the CPU restatement of the reference hot path,
oracle/bls12_381_oracle.hpp).  Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package bls12_381_b200 never imports this.

Array conventions (identical to include/bls12381_b200.h): uint64 little-endian limbs, Montgomery form;
Fp (n,6)  Fp2 (n,12)  Fp6 (n,36)  Fp12 (n,72)  G1 affine (n,12)+inf(n,) uint8  G1 projective (n,18)
G2 affine (n,24)+inf  G2 projective (n,36)  scalars (n,32) uint8 canonical little-endian.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "liboracle.so")


def _cpu_tag():
    """CPU model + ISA flags of this machine: liboracle.so is compiled -march=native, so a copy built elsewhere (it
    travels with the repo snapshot) must be rebuilt here"""
    import hashlib
    model, flags = "", ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name") and not model:
                model = line.split(":", 1)[1].strip()
            elif line.startswith("flags") and not flags:
                flags = line.split(":", 1)[1].strip()
            if model and flags:
                break
    except OSError:
        pass
    return hashlib.sha256((model + "|" + flags).encode()).hexdigest()[:16]


def build(force=False):
    src = [os.path.join(_DIR, f) for f in ("oracle_capi.cpp", "bls12_381_oracle.hpp", "h2c_oracle.hpp", "h2c_constants.inc",
                                           "Makefile")]
    stamp = os.path.join(_DIR, "liboracle.stamp")
    tag = _cpu_tag()
    stale = (not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
             or not os.path.exists(stamp) or open(stamp).read().strip() != tag)
    if force or stale:
        subprocess.check_call(["make", "-C", _DIR, "-s", "-B", "liboracle.so"])
        open(stamp, "w").write(tag)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_tower_op.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u64(a, width):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a.reshape(-1, width)


def _u8(a, width=None):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a if width is None else a.reshape(-1, width)


def hardware_threads():
    return int(lib().orc_hardware_threads())


OPS = dict(mul=0, add=1, sub=2, square=3, neg=4, invert=5, frobenius=6, conjugate=7, mul_by_nonresidue=8,
           cyclotomic_square=9)
WIDTH = {1: 6, 2: 12, 6: 36, 12: 72}


def tower(level, op, a, b=None):
    w = WIDTH[level]
    a = _u64(a, w)
    if b is not None:
        b = _u64(b, w)
    out = np.empty_like(a)
    rc = lib().orc_tower_op(level, OPS[op], _p(a), _p(b), _p(out), C.c_size_t(a.shape[0]))
    if rc != 0:
        raise ValueError("oracle: unsupported op %s at level %d" % (op, level))
    return out


def fp12_mul_by_014(f, c0, c1, c4):
    f = _u64(f, 72)
    c0, c1, c4 = _u64(c0, 12), _u64(c1, 12), _u64(c4, 12)
    out = np.empty_like(f)
    lib().orc_fp12_mul_by_014(_p(f), _p(c0), _p(c1), _p(c4), _p(out), C.c_size_t(f.shape[0]))
    return out


def fp_sqrt(a):
    a = _u64(a, 6)
    out = np.empty_like(a)
    ok = lib().orc_fp_sqrt(_p(a), _p(out))
    return bool(ok), out


def fp2_sqrt(a):
    a = _u64(a, 12)
    out = np.empty_like(a)
    ok = lib().orc_fp2_sqrt(_p(a), _p(out))
    return bool(ok), out


def fp_from_bytes(b):
    b = _u8(b)
    out = np.empty((1, 6), np.uint64)
    ok = lib().orc_fp_from_bytes(_p(b), _p(out))
    return bool(ok), out


def fp_to_bytes(a):
    a = _u64(a, 6)
    out = np.empty(48, np.uint8)
    lib().orc_fp_to_bytes(_p(a), _p(out))
    return out


def fp_lex_largest(a):
    return bool(lib().orc_fp_lex_largest(_p(_u64(a, 6))))


def scalar_from_wide(wide):
    wide = _u8(wide, 64)
    out = np.empty((wide.shape[0], 32), np.uint8)
    lib().orc_scalar_from_wide(_p(wide), _p(out), C.c_size_t(wide.shape[0]))
    return out


def scalar_to_bytes(mont):
    mont = _u64(mont, 4)
    out = np.empty((mont.shape[0], 32), np.uint8)
    lib().orc_scalar_to_bytes(_p(mont), _p(out), C.c_size_t(mont.shape[0]))
    return out


FR_OPS = dict(mul=0, add=1, sub=2, square=3, neg=4, invert=5, double=11)


def fr_op(op, a, b=None, threads=1):
    """Scalar-field ops on Montgomery limbs (n,4) uint64  (src/scalar.rs:554-627, :341, :408, :249)"""
    a = _u64(a, 4)
    b = None if b is None else _u64(b, 4)
    out = np.empty_like(a)
    rc = lib().orc_fr_op(FR_OPS[op], _p(a), _p(b), _p(out), C.c_size_t(a.shape[0]), threads)
    assert rc == 0
    return out


def fr_from_bytes(b):
    b = _u8(b, 32)
    out = np.empty((b.shape[0], 4), np.uint64)
    ok = np.empty(b.shape[0], np.uint8)
    lib().orc_fr_from_bytes(_p(b), _p(out), _p(ok), C.c_size_t(b.shape[0]))
    return out, ok


def fr_pow(a, by):
    a, by = _u64(a, 4), _u64(by, 4)
    out = np.empty((1, 4), np.uint64)
    lib().orc_fr_pow(_p(a), _p(by), _p(out))
    return out


def fr_const(name):
    out = np.empty((1, 4), np.uint64)
    lib().orc_fr_const(dict(one=0, two_inv=1, root_of_unity=2, root_of_unity_inv=3, generator=4)[name], _p(out))
    return out


def fr_dft_naive(a, inverse=False, coset=False):
    a = _u64(a, 4)
    log_n = int(a.shape[0]).bit_length() - 1
    assert a.shape[0] == 1 << log_n
    out = np.empty_like(a)
    lib().orc_fr_dft_naive(_p(a), log_n, int(inverse), int(coset), _p(out))
    return out


def fr_ntt(a, inverse=False, coset=False, threads=1):
    a = _u64(a, 4)
    log_n = int(a.shape[0]).bit_length() - 1
    assert a.shape[0] == 1 << log_n
    out = np.empty_like(a)
    lib().orc_fr_ntt(_p(a), log_n, int(inverse), int(coset), _p(out), threads)
    return out


# ---------------------------------------------------------------- hash to curve (src/hash_to_curve/)
def expand_message_xmd(msg, dst, len_in_bytes):
    msg, dst = _u8(np.frombuffer(bytes(msg), np.uint8)), _u8(np.frombuffer(bytes(dst), np.uint8))
    out = np.empty(len_in_bytes, np.uint8)
    rc = lib().orc_expand_message_xmd_sha256(_p(msg), C.c_size_t(msg.size), _p(dst), C.c_size_t(dst.size),
                                             C.c_size_t(len_in_bytes), _p(out))
    if rc != 0:
        raise ValueError("expand_message_xmd: ell > 255 or len_in_bytes > 65535")
    return out


def sha256(msg):
    msg = _u8(np.frombuffer(bytes(msg), np.uint8))
    out = np.empty(32, np.uint8)
    lib().orc_sha256(_p(msg), C.c_size_t(msg.size), _p(out))
    return out.tobytes()


def pack_messages(msgs):
    """list of bytes -> (concatenated uint8 array, offsets uint64 (n+1,))"""
    off = np.zeros(len(msgs) + 1, np.uint64)
    off[1:] = np.cumsum([len(m) for m in msgs])
    cat = np.frombuffer(b"".join(bytes(m) for m in msgs) or b"\0", np.uint8).copy()
    return cat, off


def hash_to_curve(k, msgs, dst, encode=False, threads=1):
    """HashToCurve<ExpandMsgXmd<Sha256>>::hash_to_curve / encode_to_curve for G{k} -> (n, 18k) projective"""
    cat, off = pack_messages(msgs)
    dst = _u8(np.frombuffer(bytes(dst), np.uint8))
    out = np.empty((len(msgs), 18 * k), np.uint64)
    rc = getattr(lib(), "orc_g%d_hash" % k)(_p(cat), _p(off), C.c_size_t(len(msgs)), _p(dst), C.c_size_t(dst.size),
                                            int(encode), _p(out), threads)
    assert rc == 0
    return out


H2C_STAGE = dict(g1_sswu=0, g1_iso_map=1, g1_map_to_curve=2, g1_clear_cofactor=3, g2_sswu=4, g2_iso_map=5,
                 g2_map_to_curve=6, g2_clear_cofactor=7)


def h2c_stage(name, a):
    kind = H2C_STAGE[name]
    k = 1 if kind < 4 else 2
    win = 6 * k if kind % 4 in (0, 2) else 18 * k
    a = _u64(a, win)
    out = np.empty((a.shape[0], 18 * k), np.uint64)
    rc = lib().orc_h2c_stage(kind, _p(a), _p(out), C.c_size_t(a.shape[0]))
    assert rc == 0
    return out


def fp_from_okm(okm):
    okm = _u8(okm, 64)
    out = np.empty((okm.shape[0], 6), np.uint64)
    lib().orc_fp_from_okm(_p(okm), _p(out), C.c_size_t(okm.shape[0]))
    return out


def sgn0(level, a):
    a = _u64(a, 6 * level)
    out = np.empty(a.shape[0], np.uint8)
    lib().orc_sgn0(level, _p(a), _p(out), C.c_size_t(a.shape[0]))
    return out


def fr_from_okm(okm):
    okm = _u8(okm, 48)
    out = np.empty((okm.shape[0], 4), np.uint64)
    lib().orc_fr_from_okm(_p(okm), _p(out), C.c_size_t(okm.shape[0]))
    return out


def fr_hash_to_field(msgs, dst, count=1):
    cat, off = pack_messages(msgs)
    dst = _u8(np.frombuffer(bytes(dst), np.uint8))
    out = np.empty((len(msgs) * count, 4), np.uint64)
    rc = lib().orc_fr_hash_to_field(_p(cat), _p(off), C.c_size_t(len(msgs)), _p(dst), C.c_size_t(dst.size), count, _p(out))
    assert rc == 0
    return out


class _Group:
    """G1 (k=1) or G2 (k=2) entry points; coordinates are k*6 limbs wide."""

    def __init__(self, k):
        self.k = k
        self.aw = 12 * k
        self.pw = 18 * k
        self.n = "g%d" % k

    def _f(self, name):
        return getattr(lib(), "orc_%s_%s" % (self.n, name))

    def generator(self):
        out = np.empty((1, self.pw), np.uint64)
        self._f("generator")(_p(out))
        return out

    def identity(self, n=1):
        out = np.zeros((n, self.pw), np.uint64)
        out[:, 6 * self.k:6 * self.k + 6] = R_LIMBS
        return out

    def affine_identity(self, n=1):
        out = np.zeros((n, self.aw), np.uint64)
        out[:, 6 * self.k:6 * self.k + 6] = R_LIMBS
        return out, np.ones(n, np.uint8)

    def double(self, p):
        p = _u64(p, self.pw)
        out = np.empty_like(p)
        self._f("double")(_p(p), _p(out), C.c_size_t(p.shape[0]))
        return out

    def add(self, p, q):
        p, q = _u64(p, self.pw), _u64(q, self.pw)
        out = np.empty_like(p)
        self._f("add")(_p(p), _p(q), _p(out), C.c_size_t(p.shape[0]))
        return out

    def add_mixed(self, p, qxy, qinf=None):
        p, qxy = _u64(p, self.pw), _u64(qxy, self.aw)
        qinf = None if qinf is None else _u8(qinf)
        out = np.empty_like(p)
        self._f("add_mixed")(_p(p), _p(qxy), _p(qinf), _p(out), C.c_size_t(p.shape[0]))
        return out

    def mul(self, p, s, threads=1):
        p, s = _u64(p, self.pw), _u8(s, 32)
        out = np.empty_like(p)
        self._f("mul")(_p(p), _p(s), _p(out), C.c_size_t(p.shape[0]), threads)
        return out

    def to_affine(self, p):
        p = _u64(p, self.pw)
        xy = np.empty((p.shape[0], self.aw), np.uint64)
        inf = np.empty(p.shape[0], np.uint8)
        self._f("to_affine")(_p(p), _p(xy), _p(inf), C.c_size_t(p.shape[0]))
        return xy, inf

    def batch_normalize(self, p):
        p = _u64(p, self.pw)
        xy = np.empty((p.shape[0], self.aw), np.uint64)
        inf = np.empty(p.shape[0], np.uint8)
        self._f("batch_normalize")(_p(p), _p(xy), _p(inf), C.c_size_t(p.shape[0]))
        return xy, inf

    def from_affine(self, xy, inf=None):
        xy = _u64(xy, self.aw)
        n = xy.shape[0]
        out = np.zeros((n, self.pw), np.uint64)
        out[:, :self.aw] = xy
        z = np.zeros((n, 6 * self.k), np.uint64)
        z[:, :6] = R_LIMBS
        if inf is not None:
            z[np.asarray(inf) != 0] = 0
        out[:, self.aw:] = z
        return out

    def msm_naive(self, xy, inf, s, threads=1):
        xy, s = _u64(xy, self.aw), _u8(s, 32)
        inf = None if inf is None else _u8(inf)
        out = np.empty((1, self.pw), np.uint64)
        self._f("msm_naive")(_p(xy), _p(inf), _p(s), C.c_size_t(xy.shape[0]), _p(out), threads)
        return out

    def msm_pippenger(self, xy, inf, s, c=8, threads=1):
        xy, s = _u64(xy, self.aw), _u8(s, 32)
        inf = None if inf is None else _u8(inf)
        out = np.empty((1, self.pw), np.uint64)
        self._f("msm_pippenger")(_p(xy), _p(inf), _p(s), C.c_size_t(xy.shape[0]), _p(out), c, threads)
        return out

    def checks(self, xy, inf=None):
        """per point: bit 0 = is_on_curve, bit 1 = is_torsion_free"""
        xy = _u64(xy, self.aw)
        inf = None if inf is None else _u8(inf)
        out = np.empty(xy.shape[0], np.uint8)
        self._f("checks")(_p(xy), _p(inf), C.c_size_t(xy.shape[0]), _p(out))
        return out

    def to_compressed(self, xy, inf=0):
        out = np.empty(48 * self.k, np.uint8)
        self._f("to_compressed")(_p(_u64(xy, self.aw)), int(inf), _p(out))
        return out

    def to_uncompressed(self, xy, inf=0):
        out = np.empty(96 * self.k, np.uint8)
        self._f("to_uncompressed")(_p(_u64(xy, self.aw)), int(inf), _p(out))
        return out

    def from_compressed(self, b):
        xy = np.empty((1, self.aw), np.uint64)
        inf = np.zeros(1, np.uint8)
        ok = self._f("from_compressed")(_p(_u8(b)), _p(xy), _p(inf))
        return bool(ok), xy, int(inf[0])

    def from_uncompressed(self, b):
        xy = np.empty((1, self.aw), np.uint64)
        inf = np.zeros(1, np.uint8)
        ok = self._f("from_uncompressed")(_p(_u8(b)), _p(xy), _p(inf))
        return bool(ok), xy, int(inf[0])


R_LIMBS = np.array([0x760900000002fffd, 0xebf4000bc40c0002, 0x5f48985753c758ba, 0x77ce585370525745,
                    0x5c071a97a256ec6d, 0x15f65ec3fa80e493], dtype=np.uint64)
G1 = _Group(1)
G2 = _Group(2)


def _pairs(pxy, pinf, qxy, qinf):
    pxy, qxy = _u64(pxy, 12), _u64(qxy, 24)
    pinf = None if pinf is None else _u8(pinf)
    qinf = None if qinf is None else _u8(qinf)
    return pxy, pinf, qxy, qinf


def miller_loop(pxy, pinf, qxy, qinf, threads=1):
    pxy, pinf, qxy, qinf = _pairs(pxy, pinf, qxy, qinf)
    out = np.empty((pxy.shape[0], 72), np.uint64)
    lib().orc_miller_loop(_p(pxy), _p(pinf), _p(qxy), _p(qinf), C.c_size_t(pxy.shape[0]), _p(out), threads)
    return out


def final_exponentiation(f, threads=1):
    f = _u64(f, 72)
    out = np.empty_like(f)
    lib().orc_final_exp(_p(f), C.c_size_t(f.shape[0]), _p(out), threads)
    return out


def pairing(pxy, pinf, qxy, qinf, threads=1):
    pxy, pinf, qxy, qinf = _pairs(pxy, pinf, qxy, qinf)
    out = np.empty((pxy.shape[0], 72), np.uint64)
    lib().orc_pairing(_p(pxy), _p(pinf), _p(qxy), _p(qinf), C.c_size_t(pxy.shape[0]), _p(out), threads)
    return out


def multi_miller_loop(pxy, pinf, qxy, qinf):
    pxy, pinf, qxy, qinf = _pairs(pxy, pinf, qxy, qinf)
    out = np.empty((1, 72), np.uint64)
    lib().orc_multi_miller_loop(_p(pxy), _p(pinf), _p(qxy), _p(qinf), C.c_size_t(pxy.shape[0]), _p(out))
    return out


def gt_mul(g, s, threads=1):
    g, s = _u64(g, 72), _u8(s, 32)
    out = np.empty_like(g)
    lib().orc_gt_mul(_p(g), _p(s), _p(out), C.c_size_t(g.shape[0]), threads)
    return out


def g2_prepare(qxy, qinf=0):
    out = np.empty((68, 36), np.uint64)
    n = lib().orc_g2_prepare(_p(_u64(qxy, 24)), int(qinf), _p(out))
    assert n == 68
    return out
