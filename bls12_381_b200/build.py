"""Builds bls12_381_b200/libbls12381_b200.so (the C-ABI shared library) with nvcc for sm_100a, in-tree.

    python -m bls12_381_b200.build [--force]

One nvcc per translation unit, run in parallel; objects land in bls12_381_b200/build/ (git-ignored),
the .so next to this file (git-ignored too, but it travels to the GPU box with the gpurun snapshot).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libbls12381_b200.so")
OBJ = os.path.join(HERE, "build")
# (source, extra nvcc flags).  The pairing kernels have their own unit with their own Fp2-multiply variant (fp2.cuh).
UNITS = [("capi_basic.cu", []), ("capi_pairing.cu", []), ("capi_msm.cu", []), ("capi_serial.cu", []),
         ("pairing_v4.cu", []), ("pairing_coop.cu", []), ("capi_multi.cu", []), ("capi_fr.cu", []), ("capi_h2c.cu", []), ("capi_gt.cu", [])]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h", ".inc")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    h.update((" ".join(NVCC_FLAGS) + repr(UNITS)).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    nvcc = _nvcc()

    def one(spec):
        unit, extra = spec
        obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, unit), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        open(obj + ".log", "w").write(r.stdout + r.stderr)      # ptxas -v: registers / spills per kernel
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (unit, r.stderr[-4000:]))
        return obj

    with ThreadPoolExecutor(len(UNITS)) as ex:
        objs = list(ex.map(one, UNITS))
    cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    open(stamp, "w").write(dig)
    if verbose:
        print("built", OUT)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
