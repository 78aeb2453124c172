"""TEST INFRASTRUCTURE: compiles the device headers of bls12_381_b200/csrc for the host (see cuda_host_shim.h)."""
import hashlib
import os
import subprocess

_DIR = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_DIR))
_CSRC = os.path.join(_ROOT, "bls12_381_b200", "csrc")


def _digest(extra):
    h = hashlib.sha256(extra.encode())
    for d in (_DIR, _CSRC):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cuh", ".h", ".cpp", ".inc", ".cu", "build.py")):
                h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()


def build(variant="default"):
    """variant: 'default' (lazy-reduction Fp2: capi_basic/capi_msm/capi_serial), 'kcall' (pairing units) or 'kdual'
    (the experimental dual-stream Fp2 of pairing_v5.cu)"""
    so = os.path.join(_DIR, "libemul_%s.so" % variant)
    flags = {"kcall": ["-DB200_FP2_KCALL"], "lazy3": ["-DB200_FP2_LAZY3"]}.get(variant, [])
    if os.path.exists(os.path.join(_CSRC, "fr.cuh")):
        flags.append("-DEMUL_WITH_FR")
    stamp = so + ".stamp"
    dg = _digest(" ".join(flags))
    if os.path.exists(so) and os.path.exists(stamp) and open(stamp).read() == dg:
        return so
    cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas", "-I", _DIR, "-I", _CSRC,
           "-I", os.path.join(_ROOT, "include")] + flags + [os.path.join(_DIR, "emul_capi.cpp"), "-o", so]
    subprocess.check_call(cmd)
    open(stamp, "w").write(dg)
    return so


def build_cabi():
    """The HOST side of the C ABI against the mock CUDA runtime (tests/emul/mock): every unit — capi_basic.cu (the real ctx, tuning, timing, tower / group / config-1
    kernels), capi_pairing.cu + pairing_v4/v5/v6.cu, capi_serial.cu, capi_fr.cu, capi_h2c.cu, capi_gt.cu, capi_msm.cu (cp.async guarded in the
    source, match.any / shuffles / atomics on the fiber scheduler).  Every launch
    runs on the fiber scheduler (fiber_warp.h), so warp- and block-cooperative kernels work.  -> libemul_cabi.so"""
    so = os.path.join(_DIR, "libemul_cabi.so")
    stamp = so + ".stamp"
    dg = _digest("cabi_full")
    if os.path.exists(so) and os.path.exists(stamp) and open(stamp).read() == dg:
        return so
    common = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-pthread", "-Wno-unknown-pragmas", "-Wno-unused-variable",
              "-DEMUL_LAUNCH_COOPERATIVE", "-I", os.path.join(_DIR, "mock"), "-I", _DIR,
              "-I", _CSRC, "-I", os.path.join(_ROOT, "include"), "-include", os.path.join(_DIR, "cuda_host_shim.h")]
    units = ["capi_basic.cu", "capi_msm.cu", "capi_pairing.cu", "pairing_v4.cu", "pairing_coop.cu", "capi_multi.cu", "capi_serial.cu",
             "capi_fr.cu", "capi_h2c.cu", "capi_gt.cu"]
    from concurrent.futures import ThreadPoolExecutor

    def one(u):
        obj = os.path.join(_DIR, u + ".emul.o")
        subprocess.check_call(common + ["-x", "c++", "-c", os.path.join(_CSRC, u), "-o", obj])
        return obj

    with ThreadPoolExecutor(len(units)) as ex:
        objs = list(ex.map(one, units))
    subprocess.check_call(["g++", "-shared", "-pthread", "-o", so] + objs + ["-ldl"])
    open(stamp, "w").write(dg)
    return so
