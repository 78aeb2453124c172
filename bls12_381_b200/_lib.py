"""ctypes loader for libbls12381_b200.so (the C ABI declared in include/bls12381_b200.h).

There is NO CPU fallback: if the shared library is missing it is built with nvcc; if that fails, or no
CUDA device is usable when a context is created, the error propagates."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libbls12381_b200.so")
HEADER = os.path.join(_HERE, "..", "include", "bls12381_b200.h")

_vp, _sz, _i = C.c_void_p, C.c_size_t, C.c_int
# name -> argtypes (restype is int unless listed in _RESTYPE)
SIGNATURES = {
    "b200_ctx_create": [_i, C.POINTER(_vp)],
    "b200_ctx_create_on_stream": [_i, _vp, C.POINTER(_vp)],
    "b200_ctx_destroy": [_vp],
    "b200_strerror": [_i],
    "b200_last_error": [_vp],
    "b200_ctx_device": [_vp],
    "b200_ctx_stream": [_vp],
    "b200_ctx_launch_count": [_vp],
    "b200_ctx_set_msm_window": [_vp, _i],
    "b200_ctx_set_tuning": [_vp, C.c_char_p, _i],
    "b200_ctx_set_timing": [_vp, _i],
    "b200_ctx_get_timing": [_vp, C.c_char_p, _sz, C.POINTER(C.c_float), _i],
    "b200_tower_op": [_vp, _i, _i, _vp, _vp, _vp, _sz],
    "b200_glv_decompose": [_vp, _vp, _sz, _vp, _vp],
    "b200_imad_peak": [_vp, _i, C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "b200_imad_peak_mode": [_vp, _i, _i, C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "b200_gt_mul_batch": [_vp, _vp, _vp, _sz, _vp],
    "b200_gt_mul_batch_dev": [_vp, _vp, _vp, _sz, _vp],
    "b200_fr_op": [_vp, _i, _vp, _vp, _sz, _vp],
    "b200_fr_to_bytes": [_vp, _vp, _sz, _vp],
    "b200_fr_from_bytes": [_vp, _vp, _sz, _vp, _vp],
    "b200_fr_ntt": [_vp, _vp, _i, _i, _i, _vp],
    "b200_fr_op_dev": [_vp, _i, _vp, _vp, _sz, _vp],
    "b200_fr_ntt_dev": [_vp, _vp, _i, _i, _i, _vp],
    "b200_expand_message_xmd_sha256": [_vp, _vp, _vp, _sz, _vp, _sz, _sz, _vp],
    "b200_g1_hash_to_curve": [_vp, _vp, _vp, _sz, _vp, _sz, _i, _vp],
    "b200_g2_hash_to_curve": [_vp, _vp, _vp, _sz, _vp, _sz, _i, _vp],
    "b200_h2c_stage": [_vp, _i, _i, _vp, _sz, _vp],
    "b200_fr_from_okm": [_vp, _vp, _sz, _vp],
    "b200_fr_hash_to_field": [_vp, _vp, _vp, _sz, _vp, _sz, _i, _vp],
    "b200_miller_loop_batch": [_vp, _vp, _vp, _vp, _vp, _sz, _vp],
    "b200_final_exponentiation_batch": [_vp, _vp, _sz, _vp],
    "b200_pairing_batch": [_vp, _vp, _vp, _vp, _vp, _sz, _vp],
    "b200_multi_miller_loop": [_vp, _vp, _vp, _vp, _vp, _sz, _vp],
    "b200_multi_miller_loop_dev": [_vp, _vp, _vp, _vp, _vp, _sz, _vp],
    "b200_pairing_product_batch": [_vp, _vp, _vp, _vp, _vp, _sz, _sz, _i, _vp],
    "b200_pairing_product_batch_dev": [_vp, _vp, _vp, _vp, _vp, _sz, _sz, _i, _vp],
    "b200_miller_loop_batch_dev": [_vp, _vp, _vp, _vp, _vp, _sz, _vp],
    "b200_final_exponentiation_batch_dev": [_vp, _vp, _sz, _vp],
    "b200_pairing_batch_dev": [_vp, _vp, _vp, _vp, _vp, _sz, _vp],
    "b200_fp12_product_dev": [_vp, _vp, _sz, _vp],
    "b200_g2_prepare": [_vp, _vp, _vp, _sz, _vp],
    "b200_multi_miller_loop_prepared": [_vp, _vp, _vp, _vp, _vp, _sz, _vp],
    "b200_g2_prepare_dev": [_vp, _vp, _vp, _sz, _vp],
    "b200_miller_loop_prepared_batch_dev": [_vp, _vp, _vp, _vp, _vp, _sz, _vp],
    "b200_comm_unique_id": [_vp],
    "b200_ctx_comm_init": [_vp, _vp, _i, _i],
    "b200_ctx_comm_destroy": [_vp],
    "b200_ctx_comm_rank": [_vp],
    "b200_ctx_comm_world": [_vp],
    "b200_multi_create": [_i, C.POINTER(_vp)],
    "b200_multi_destroy": [_vp],
    "b200_multi_gpus": [_vp],
    "b200_multi_ctx": [_vp, _i],
    "b200_multi_set_sharding": [_vp, _i],
    "b200_multi_g1_msm": [_vp, _vp, _vp, _vp, _sz, _vp],
    "b200_multi_g2_msm": [_vp, _vp, _vp, _vp, _sz, _vp],
}
for _g in ("g1", "g2"):
    SIGNATURES.update({
        "b200_%s_mul_batch" % _g: [_vp, _vp, _vp, _sz, _vp],
        "b200_%s_double_batch" % _g: [_vp, _vp, _sz, _vp],
        "b200_%s_add_batch" % _g: [_vp, _vp, _vp, _sz, _vp],
        "b200_%s_add_mixed_batch" % _g: [_vp, _vp, _vp, _vp, _sz, _vp],
        "b200_%s_batch_normalize" % _g: [_vp, _vp, _sz, _vp, _vp],
        "b200_%s_msm" % _g: [_vp, _vp, _vp, _vp, _sz, _vp],
        "b200_%s_mul_batch_dev" % _g: [_vp, _vp, _vp, _sz, _vp],
        "b200_%s_batch_normalize_dev" % _g: [_vp, _vp, _sz, _vp, _vp],
        "b200_%s_msm_dev" % _g: [_vp, _vp, _vp, _vp, _sz, _vp],
        "b200_%s_msm_shard_dev" % _g: [_vp, _vp, _vp, _vp, _sz, _i, _i, _vp],
        "b200_%s_sum_dev" % _g: [_vp, _vp, _sz, _vp],
        "b200_%s_msm_sharded_dev" % _g: [_vp, _vp, _vp, _vp, _sz, _i, _vp],
        "b200_%s_check" % _g: [_vp, _vp, _vp, _sz, _vp],
        "b200_%s_serialize" % _g: [_vp, _vp, _vp, _sz, _i, _vp],
        "b200_%s_deserialize" % _g: [_vp, _vp, _sz, _i, _vp, _vp, _vp],
    })
_RESTYPE = {"b200_ctx_destroy": None, "b200_multi_destroy": None, "b200_multi_ctx": _vp, "b200_strerror": C.c_char_p, "b200_last_error": C.c_char_p,
            "b200_ctx_stream": _vp, "b200_ctx_launch_count": C.c_uint64}

_lib = None


def load():
    """Loads (building first if needed) the CUDA shared library. Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    _build.build()          # no-op when the digest stamp of csrc/ + include/ matches the built library (never run a stale .so)
    lib = C.CDLL(SO_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, C.c_int)
    _lib = lib
    return lib
