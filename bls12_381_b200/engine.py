"""Host-side harness over the C ABI (include/bls12381_b200.h).

The reference is a Rust crate; its drop-in host shim is the Rust `-sys` binding shown in INTEGRATION.md
and the C++ mirror in bls12_381_b200/host/bls12_381.hpp.  This module is the Python face of the same
ABI used by tests/ and bench.py.  Method names follow the reference functions they batch:
`g1_mul_batch` = G1Projective * Scalar (src/g1.rs:556), `g1_msm` = sum_i p_i*s_i (:573, :161),
`batch_normalize` (:806), `pairing_batch` = pairing() (src/pairings.rs:607), `multi_miller_loop` (:554),
`final_exponentiation_batch` (:48).

Arrays: numpy uint64 limbs / uint8 flags+scalars for the host-pointer entry points; torch CUDA tensors
(any dtype, contiguous, byte-exact layout) for the `_dev` entry points, which run on the engine's own
stream and return after synchronising it.
"""
import ctypes as C

import numpy as np

from . import _lib

OPS = dict(mul=0, add=1, sub=2, square=3, neg=4, invert=5, frobenius=6, conjugate=7, mul_by_nonresidue=8,
           cyclotomic_square=9, invert_fast=10)


class B200Error(RuntimeError):
    pass


def _np(a, dtype, width=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a if width is None else a.reshape(-1, width)


def _hp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dp(t):
    """device pointer of a torch CUDA tensor (or None)"""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("expected a contiguous CUDA tensor")
    return C.c_void_p(t.data_ptr())


class Engine:
    """One b200_ctx = one GPU + one stream + scratch memory."""

    AFF = {1: 12, 2: 24}
    PROJ = {1: 18, 2: 36}

    def __init__(self, device=-1, stream=None):
        """stream: optional cudaStream_t (int) owned by the caller, e.g. torch.cuda.Stream().cuda_stream — the ctx then
        enqueues on it and never destroys it (b200_ctx_create_on_stream)"""
        self.lib = _lib.load()
        h = C.c_void_p()
        if stream:
            rc = self.lib.b200_ctx_create_on_stream(int(device), C.c_void_p(int(stream)), C.byref(h))
        else:
            rc = self.lib.b200_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise B200Error("b200_ctx_create: %s" % self.lib.b200_strerror(rc).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200_ctx_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc, what):
        if rc != 0:
            raise B200Error("%s: %s (%s)" % (what, self.lib.b200_strerror(rc).decode(),
                                             self.lib.b200_last_error(self.h).decode()))

    @property
    def device(self):
        return self.lib.b200_ctx_device(self.h)

    @property
    def stream(self):
        return self.lib.b200_ctx_stream(self.h)

    @property
    def launches(self):
        return int(self.lib.b200_ctx_launch_count(self.h))

    def set_msm_window(self, c):
        return self.lib.b200_ctx_set_msm_window(self.h, int(c))

    def set_tuning(self, key, value):
        self._ck(self.lib.b200_ctx_set_tuning(self.h, key.encode(), int(value)), "set_tuning(%s)" % key)

    def set_timing(self, on=True):
        self._ck(self.lib.b200_ctx_set_timing(self.h, int(bool(on))), "set_timing")

    def get_timing(self, max_records=8192):
        """[(kernel_name, ms), ...] for every launch since set_timing(True)"""
        names = C.create_string_buffer(64 * max_records)
        ms = (C.c_float * max_records)()
        n = self.lib.b200_ctx_get_timing(self.h, names, len(names), ms, max_records)
        if n < 0:
            self._ck(n, "get_timing")
        nm = names.value.decode().split("\n")
        return [(nm[i], float(ms[i])) for i in range(min(n, max_records)) if i < len(nm) and nm[i]]

    def imad_peak(self, iters=2000, mode=0):
        """mode 0: mad.wide.u32 streams (the roofline denominator); 1: the same product as a mad.lo.cc / madc.hi pair"""
        v, ms = C.c_double(), C.c_double()
        self._ck(self.lib.b200_imad_peak_mode(self.h, iters, mode, C.byref(v), C.byref(ms)), "imad_peak")
        return v.value, ms.value

    # ---------------------------------------------------------------- field tower (parity surface)
    def tower(self, level, op, a, b=None):
        w = 6 * level
        a = _np(a, np.uint64, w)
        b = None if b is None else _np(b, np.uint64, w)
        out = np.empty_like(a)
        self._ck(self.lib.b200_tower_op(self.h, level, OPS[op], _hp(a), _hp(b), _hp(out), a.shape[0]), "tower_op")
        return out

    # ---------------------------------------------------------------- groups, host pointers
    def _g(self, k):
        return "b200_g%d_" % k

    def mul_batch(self, k, p, s):
        p, s = _np(p, np.uint64, self.PROJ[k]), _np(s, np.uint8, 32)
        out = np.empty_like(p)
        self._ck(getattr(self.lib, self._g(k) + "mul_batch")(self.h, _hp(p), _hp(s), p.shape[0], _hp(out)), "mul_batch")
        return out

    def double_batch(self, k, p):
        p = _np(p, np.uint64, self.PROJ[k])
        out = np.empty_like(p)
        self._ck(getattr(self.lib, self._g(k) + "double_batch")(self.h, _hp(p), p.shape[0], _hp(out)), "double_batch")
        return out

    def add_batch(self, k, p, q):
        p, q = _np(p, np.uint64, self.PROJ[k]), _np(q, np.uint64, self.PROJ[k])
        out = np.empty_like(p)
        self._ck(getattr(self.lib, self._g(k) + "add_batch")(self.h, _hp(p), _hp(q), p.shape[0], _hp(out)), "add_batch")
        return out

    def add_mixed_batch(self, k, p, qxy, qinf=None):
        p, qxy = _np(p, np.uint64, self.PROJ[k]), _np(qxy, np.uint64, self.AFF[k])
        qinf = None if qinf is None else _np(qinf, np.uint8)
        out = np.empty_like(p)
        self._ck(getattr(self.lib, self._g(k) + "add_mixed_batch")(self.h, _hp(p), _hp(qxy), _hp(qinf), p.shape[0],
                                                                   _hp(out)), "add_mixed_batch")
        return out

    def batch_normalize(self, k, p):
        p = _np(p, np.uint64, self.PROJ[k])
        xy = np.empty((p.shape[0], self.AFF[k]), np.uint64)
        inf = np.empty(p.shape[0], np.uint8)
        self._ck(getattr(self.lib, self._g(k) + "batch_normalize")(self.h, _hp(p), p.shape[0], _hp(xy), _hp(inf)),
                 "batch_normalize")
        return xy, inf

    def msm(self, k, xy, inf, s):
        xy, s = _np(xy, np.uint64, self.AFF[k]), _np(s, np.uint8, 32)
        if xy.shape[0] != s.shape[0]:
            raise ValueError("points/scalars length mismatch")     # the reference's zip() would truncate
        inf = None if inf is None else _np(inf, np.uint8)
        out = np.empty((1, self.PROJ[k]), np.uint64)
        self._ck(getattr(self.lib, self._g(k) + "msm")(self.h, _hp(xy), _hp(inf), _hp(s), xy.shape[0], _hp(out)), "msm")
        return out

    def glv_decompose(self, s):
        """(k1, k2) signed Python ints per scalar with k = k1 + k2*lambda (mod q) — test surface of csrc/glv.cuh"""
        s = _np(s, np.uint8, 32)
        n = s.shape[0]
        kk = np.empty((n, 32), np.uint8)
        sg = np.empty(n, np.uint8)
        self._ck(self.lib.b200_glv_decompose(self.h, _hp(s), n, _hp(kk), _hp(sg)), "glv_decompose")
        out = []
        for i in range(n):
            k1 = int.from_bytes(kk[i, :16].tobytes(), "little")
            k2 = int.from_bytes(kk[i, 16:].tobytes(), "little")
            out.append((-k1 if sg[i] & 1 else k1, -k2 if sg[i] & 2 else k2))
        return out

    def gt_mul_batch(self, g, s):
        """out[i] = &Gt * &Scalar (src/pairings.rs:296-323): (n,72) Fp12 limbs, (n,32) canonical LE scalars"""
        g, s = _np(g, np.uint64, 72), _np(s, np.uint8, 32)
        if g.shape[0] != s.shape[0]:
            raise ValueError("gt_mul_batch: length mismatch")
        out = np.empty_like(g)
        self._ck(self.lib.b200_gt_mul_batch(self.h, _hp(g), _hp(s), g.shape[0], _hp(out)), "gt_mul_batch")
        return out

    # ---------------------------------------------------------------- scalar field Fr + NTT
    FR_OPS = dict(mul=0, add=1, sub=2, square=3, neg=4, invert=5, double=11)

    def fr_op(self, op, a, b=None):
        """element-wise Scalar arithmetic on Montgomery limbs (n,4) uint64 (src/scalar.rs:554-627, :341, :408, :249)"""
        a = _np(a, np.uint64, 4)
        b = None if b is None else _np(b, np.uint64, 4)
        out = np.empty_like(a)
        self._ck(self.lib.b200_fr_op(self.h, self.FR_OPS[op], _hp(a), _hp(b), a.shape[0], _hp(out)), "fr_op")
        return out

    def fr_to_bytes(self, a):
        """Scalar::to_bytes for a batch -> (n,32) uint8 canonical little-endian (the scalars msm()/mul_batch() take)"""
        a = _np(a, np.uint64, 4)
        out = np.empty((a.shape[0], 32), np.uint8)
        self._ck(self.lib.b200_fr_to_bytes(self.h, _hp(a), a.shape[0], _hp(out)), "fr_to_bytes")
        return out

    def fr_from_bytes(self, b):
        """Scalar::from_bytes for a batch -> (limbs (n,4), ok (n,)); non-canonical encodings give ok = 0, limbs 0"""
        b = _np(b, np.uint8, 32)
        out = np.empty((b.shape[0], 4), np.uint64)
        ok = np.empty(b.shape[0], np.uint8)
        self._ck(self.lib.b200_fr_from_bytes(self.h, _hp(b), b.shape[0], _hp(out), _hp(ok)), "fr_from_bytes")
        return out, ok

    def fr_ntt(self, a, inverse=False, coset=False):
        """NTT over Fr on n = 2^k Montgomery elements, natural order in/out (see include/bls12381_b200.h)"""
        a = _np(a, np.uint64, 4)
        n = a.shape[0]
        log_n = n.bit_length() - 1
        if n < 1 or n != 1 << log_n:
            raise ValueError("fr_ntt: length must be a power of two")
        out = np.empty_like(a)
        self._ck(self.lib.b200_fr_ntt(self.h, _hp(a), log_n, int(inverse), int(coset), _hp(out)), "fr_ntt")
        return out

    def fr_op_dev(self, op, a, b, n, out):
        self._ck(self.lib.b200_fr_op_dev(self.h, self.FR_OPS[op], _dp(a), _dp(b), n, _dp(out)), "fr_op_dev")

    def fr_ntt_dev(self, a, log_n, out, inverse=False, coset=False):
        self._ck(self.lib.b200_fr_ntt_dev(self.h, _dp(a), log_n, int(inverse), int(coset), _dp(out)), "fr_ntt_dev")

    # ---------------------------------------------------------------- hash to curve
    @staticmethod
    def _pack(msgs):
        off = np.zeros(len(msgs) + 1, np.uint64)
        if len(msgs):
            off[1:] = np.cumsum([len(m) for m in msgs])
        cat = np.frombuffer(b"".join(bytes(m) for m in msgs) or b"\0", np.uint8).copy()
        return cat, off

    def expand_message_xmd(self, msgs, dst, len_in_bytes):
        """ExpandMsgXmd<Sha256> for a list of byte strings -> (n, len_in_bytes) uint8 (src/hash_to_curve/expand_msg.rs:230)"""
        cat, off = self._pack(msgs)
        d = np.frombuffer(bytes(dst) or b"\0", np.uint8).copy()
        out = np.empty((len(msgs), len_in_bytes), np.uint8)
        self._ck(self.lib.b200_expand_message_xmd_sha256(self.h, _hp(cat), _hp(off), len(msgs), _hp(d), len(dst), len_in_bytes,
                                                         _hp(out)), "expand_message_xmd")
        return out

    def hash_to_curve(self, k, msgs, dst, encode=False):
        """HashToCurve<ExpandMsgXmd<Sha256>>::hash_to_curve (encode=False) / encode_to_curve for G{k}: list of byte
        strings -> (n, 18k) projective limbs (src/hash_to_curve/mod.rs:86-108)"""
        cat, off = self._pack(msgs)
        d = np.frombuffer(bytes(dst) or b"\0", np.uint8).copy()
        out = np.empty((len(msgs), self.PROJ[k]), np.uint64)
        self._ck(getattr(self.lib, self._g(k) + "hash_to_curve")(self.h, _hp(cat), _hp(off), len(msgs), _hp(d), len(dst),
                                                                  int(encode), _hp(out)), "hash_to_curve")
        return out

    def fr_from_okm(self, okm):
        """Scalar::from_okm on (n, 48) uniform bytes -> (n, 4) Montgomery limbs (src/hash_to_curve/map_scalar.rs:17)"""
        okm = _np(okm, np.uint8, 48)
        out = np.empty((okm.shape[0], 4), np.uint64)
        self._ck(self.lib.b200_fr_from_okm(self.h, _hp(okm), okm.shape[0], _hp(out)), "fr_from_okm")
        return out

    def fr_hash_to_field(self, msgs, dst, count=1):
        """Scalar::hash_to_field::<ExpandMsgXmd<Sha256>>: `count` scalars per message -> (n * count, 4)"""
        cat, off = self._pack(msgs)
        d = np.frombuffer(bytes(dst) or b"\0", np.uint8).copy()
        out = np.empty((len(msgs) * count, 4), np.uint64)
        self._ck(self.lib.b200_fr_hash_to_field(self.h, _hp(cat), _hp(off), len(msgs), _hp(d), len(dst), count, _hp(out)),
                 "fr_hash_to_field")
        return out

    H2C_KIND = dict(sswu=0, iso_map=1, map_to_curve=2, clear_cofactor=3)

    def h2c_stage(self, k, kind, a):
        kind = self.H2C_KIND[kind]
        a = _np(a, np.uint64, 6 * k if kind in (0, 2) else self.PROJ[k])
        out = np.empty((a.shape[0], self.PROJ[k]), np.uint64)
        self._ck(self.lib.b200_h2c_stage(self.h, k, kind, _hp(a), a.shape[0], _hp(out)), "h2c_stage")
        return out

    # ---------------------------------------------------------------- (de)serialization
    def serialize(self, k, xy, inf=None, compressed=True):
        """G{k}Affine::to_compressed / to_uncompressed for a batch -> (n, 48k | 96k) uint8"""
        xy = _np(xy, np.uint64, self.AFF[k])
        inf = None if inf is None else _np(inf, np.uint8)
        out = np.empty((xy.shape[0], (48 if compressed else 96) * k), np.uint8)
        self._ck(getattr(self.lib, self._g(k) + "serialize")(self.h, _hp(xy), _hp(inf), xy.shape[0], int(compressed),
                                                            _hp(out)), "serialize")
        return out

    def check(self, k, xy, inf=None):
        """per point: bit 0 = is_on_curve, bit 1 = is_torsion_free (src/g1.rs:401-418, src/g2.rs:475-491)"""
        xy = _np(xy, np.uint64, self.AFF[k])
        inf = None if inf is None else _np(inf, np.uint8)
        st = np.empty(xy.shape[0], np.uint8)
        self._ck(getattr(self.lib, self._g(k) + "check")(self.h, _hp(xy), _hp(inf), xy.shape[0], _hp(st)), "check")
        return st

    def deserialize(self, k, data, compressed=True):
        """from_{un,}compressed_unchecked + is_on_curve -> (xy, inf, status); status bit0 = Some, bit1 = on curve"""
        data = _np(data, np.uint8, (48 if compressed else 96) * k)
        n = data.shape[0]
        xy = np.empty((n, self.AFF[k]), np.uint64)
        inf = np.empty(n, np.uint8)
        st = np.empty(n, np.uint8)
        self._ck(getattr(self.lib, self._g(k) + "deserialize")(self.h, _hp(data), n, int(compressed), _hp(xy), _hp(inf),
                                                              _hp(st)), "deserialize")
        return xy, inf, st

    # ---------------------------------------------------------------- pairings, host pointers
    def _pairs(self, pxy, pinf, qxy, qinf):
        pxy, qxy = _np(pxy, np.uint64, 12), _np(qxy, np.uint64, 24)
        if pxy.shape[0] != qxy.shape[0]:
            raise ValueError("p/q length mismatch")
        pinf = None if pinf is None else _np(pinf, np.uint8)
        qinf = None if qinf is None else _np(qinf, np.uint8)
        return pxy, pinf, qxy, qinf

    def miller_loop_batch(self, pxy, pinf, qxy, qinf):
        pxy, pinf, qxy, qinf = self._pairs(pxy, pinf, qxy, qinf)
        out = np.empty((pxy.shape[0], 72), np.uint64)
        self._ck(self.lib.b200_miller_loop_batch(self.h, _hp(pxy), _hp(pinf), _hp(qxy), _hp(qinf), pxy.shape[0],
                                                 _hp(out)), "miller_loop_batch")
        return out

    def final_exponentiation_batch(self, f):
        f = _np(f, np.uint64, 72)
        out = np.empty_like(f)
        self._ck(self.lib.b200_final_exponentiation_batch(self.h, _hp(f), f.shape[0], _hp(out)), "final_exp")
        return out

    def pairing_batch(self, pxy, pinf, qxy, qinf, out=None):
        """out: optional caller-owned (n, 72) uint64 result buffer (a pinned one makes the 576-byte-per-pair read-back a
        direct DMA instead of a staged copy into freshly faulted pageable memory)"""
        pxy, pinf, qxy, qinf = self._pairs(pxy, pinf, qxy, qinf)
        if out is None:
            out = np.empty((pxy.shape[0], 72), np.uint64)
        elif out.shape != (pxy.shape[0], 72) or out.dtype != np.uint64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous (n, 72) uint64 array")
        self._ck(self.lib.b200_pairing_batch(self.h, _hp(pxy), _hp(pinf), _hp(qxy), _hp(qinf), pxy.shape[0], _hp(out)),
                 "pairing_batch")
        return out

    def multi_miller_loop(self, pxy, pinf, qxy, qinf):
        pxy, pinf, qxy, qinf = self._pairs(pxy, pinf, qxy, qinf)
        out = np.empty((1, 72), np.uint64)
        self._ck(self.lib.b200_multi_miller_loop(self.h, _hp(pxy), _hp(pinf), _hp(qxy), _hp(qinf), pxy.shape[0],
                                                 _hp(out)), "multi_miller_loop")
        return out

    def pairing_product_batch(self, pxy, pinf, qxy, qinf, terms, final_exp=True):
        """n_products x `terms` pairs -> (n_products, 72): multi_miller_loop of every product with one shared squaring per
        bit (src/pairings.rs:554-603), followed by final_exponentiation when final_exp"""
        pxy, pinf, qxy, qinf = self._pairs(pxy, pinf, qxy, qinf)
        if terms < 1 or pxy.shape[0] % terms:
            raise ValueError("pairing_product_batch: the number of pairs must be a multiple of terms")
        npr = pxy.shape[0] // terms
        out = np.empty((npr, 72), np.uint64)
        self._ck(self.lib.b200_pairing_product_batch(self.h, _hp(pxy), _hp(pinf), _hp(qxy), _hp(qinf), terms, npr,
                                                     int(bool(final_exp)), _hp(out)), "pairing_product_batch")
        return out

    def pairing_product_batch_dev(self, p, pinf, q, qinf, terms, n_products, out, final_exp=True):
        self._ck(self.lib.b200_pairing_product_batch_dev(self.h, _dp(p), _dp(pinf), _dp(q), _dp(qinf), terms, n_products,
                                                         int(bool(final_exp)), _dp(out)), "pairing_product_batch_dev")

    def multi_miller_loop_dev(self, p, pinf, q, qinf, n, out):
        self._ck(self.lib.b200_multi_miller_loop_dev(self.h, _dp(p), _dp(pinf), _dp(q), _dp(qinf), n, _dp(out)),
                 "multi_miller_loop_dev")

    def g2_prepare(self, qxy, qinf=None):
        """G2Prepared::from for a batch -> (n, 68, 36) uint64 line coefficients (src/pairings.rs:504-546)"""
        qxy = _np(qxy, np.uint64, 24)
        qinf = None if qinf is None else _np(qinf, np.uint8)
        out = np.empty((qxy.shape[0], 68, 36), np.uint64)
        self._ck(self.lib.b200_g2_prepare(self.h, _hp(qxy), _hp(qinf), qxy.shape[0], _hp(out)), "g2_prepare")
        return out

    def multi_miller_loop_prepared(self, pxy, pinf, coeffs, qinf=None):
        pxy = _np(pxy, np.uint64, 12)
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 68, 36)
        if coeffs.shape[0] != pxy.shape[0]:
            raise ValueError("p/prepared length mismatch")
        pinf = None if pinf is None else _np(pinf, np.uint8)
        qinf = None if qinf is None else _np(qinf, np.uint8)
        out = np.empty((1, 72), np.uint64)
        self._ck(self.lib.b200_multi_miller_loop_prepared(self.h, _hp(pxy), _hp(pinf), _hp(coeffs), _hp(qinf), pxy.shape[0],
                                                          _hp(out)), "multi_miller_loop_prepared")
        return out

    def g2_prepare_dev(self, q, qinf, n, coeffs):
        self._ck(self.lib.b200_g2_prepare_dev(self.h, _dp(q), _dp(qinf), n, _dp(coeffs)), "g2_prepare_dev")

    def miller_loop_prepared_batch_dev(self, p, pinf, coeffs, qinf, n, out):
        self._ck(self.lib.b200_miller_loop_prepared_batch_dev(self.h, _dp(p), _dp(pinf), _dp(coeffs), _dp(qinf), n, _dp(out)),
                 "miller_loop_prepared_batch_dev")

    # ---------------------------------------------------------------- device-pointer entry points (torch tensors)
    def mul_batch_dev(self, k, p, s, out, n):
        self._ck(getattr(self.lib, self._g(k) + "mul_batch_dev")(self.h, _dp(p), _dp(s), n, _dp(out)), "mul_batch_dev")

    def batch_normalize_dev(self, k, p, n, out_xy, out_inf):
        self._ck(getattr(self.lib, self._g(k) + "batch_normalize_dev")(self.h, _dp(p), n, _dp(out_xy), _dp(out_inf)),
                 "batch_normalize_dev")

    def msm_dev(self, k, xy, inf, s, n, out, shard=0, n_shards=1):
        self._ck(getattr(self.lib, self._g(k) + "msm_shard_dev")(self.h, _dp(xy), _dp(inf), _dp(s), n, shard, n_shards,
                                                                 _dp(out)), "msm_dev")

    # ---------------------------------------------------------------- multi-GPU: one process per GPU (capi_multi.cu)
    SHARD = dict(window=0, points=1)

    def comm_unique_id(self):
        """128-byte NCCL id made on rank 0; ship it to the other ranks, then comm_init everywhere"""
        buf = (C.c_uint8 * 128)()
        self._ck(self.lib.b200_comm_unique_id(buf), "comm_unique_id")
        return bytes(buf)

    def comm_init(self, uid, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(uid))
        self._ck(self.lib.b200_ctx_comm_init(self.h, buf, int(rank), int(world)), "comm_init")

    def comm_destroy(self):
        self._ck(self.lib.b200_ctx_comm_destroy(self.h), "comm_destroy")

    @property
    def comm_world(self):
        return int(self.lib.b200_ctx_comm_world(self.h))

    def msm_sharded_dev(self, k, xy, inf, s, n, out, mode="window"):
        """collective: shard + ncclAllGather + combine inside the library, one host sync at the end"""
        self._ck(getattr(self.lib, self._g(k) + "msm_sharded_dev")(self.h, _dp(xy), _dp(inf), _dp(s), n, self.SHARD[mode],
                                                                   _dp(out)), "msm_sharded_dev")

    def sum_dev(self, k, parts, n, out):
        self._ck(getattr(self.lib, self._g(k) + "sum_dev")(self.h, _dp(parts), n, _dp(out)), "sum_dev")

    def miller_loop_batch_dev(self, p, pinf, q, qinf, n, out):
        self._ck(self.lib.b200_miller_loop_batch_dev(self.h, _dp(p), _dp(pinf), _dp(q), _dp(qinf), n, _dp(out)),
                 "miller_loop_batch_dev")

    def final_exponentiation_batch_dev(self, f, n, out):
        self._ck(self.lib.b200_final_exponentiation_batch_dev(self.h, _dp(f), n, _dp(out)), "final_exp_dev")

    def pairing_batch_dev(self, p, pinf, q, qinf, n, out):
        self._ck(self.lib.b200_pairing_batch_dev(self.h, _dp(p), _dp(pinf), _dp(q), _dp(qinf), n, _dp(out)),
                 "pairing_batch_dev")

    def fp12_product_dev(self, f, n, out):
        self._ck(self.lib.b200_fp12_product_dev(self.h, _dp(f), n, _dp(out)), "fp12_product_dev")


class MultiEngine:
    """One process, several GPUs (b200_multi_*): owns a ctx + NCCL communicator + host thread per device."""

    def __init__(self, n_gpus=0, mode="points"):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.b200_multi_create(int(n_gpus), C.byref(h))
        if rc != 0:
            raise B200Error("b200_multi_create: %s" % self.lib.b200_strerror(rc).decode())
        self.h = h
        self.set_sharding(mode)

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200_multi_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def gpus(self):
        return int(self.lib.b200_multi_gpus(self.h))

    def set_sharding(self, mode):
        rc = self.lib.b200_multi_set_sharding(self.h, Engine.SHARD[mode])
        if rc != 0:
            raise B200Error("b200_multi_set_sharding: %s" % self.lib.b200_strerror(rc).decode())

    def msm(self, k, xy, inf, s):
        xy, s = _np(xy, np.uint64, Engine.AFF[k]), _np(s, np.uint8, 32)
        if xy.shape[0] != s.shape[0]:
            raise ValueError("points/scalars length mismatch")
        inf = None if inf is None else _np(inf, np.uint8)
        out = np.empty((1, Engine.PROJ[k]), np.uint64)
        rc = getattr(self.lib, "b200_multi_g%d_msm" % k)(self.h, _hp(xy), _hp(inf), _hp(s), xy.shape[0], _hp(out))
        if rc != 0:
            raise B200Error("b200_multi_g%d_msm: %s" % (k, self.lib.b200_strerror(rc).decode()))
        return out
