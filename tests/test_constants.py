"""Constants derived from first principles (tools/gen_constants.py, bls12_381_b200/constants_host.py)
equal the limbs the reference hard-codes (via the pinned oracle and the extracted KATs)."""
import re
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _words(name):
    s = open(os.path.join(ROOT, "bls12_381_b200", "csrc", "constants.cuh")).read()
    m = re.search(name + r"\[12\] = \{([^}]*)\}", s)
    w = [int(x.strip().rstrip("u"), 16) for x in m.group(1).split(",")]
    return np.array([w[2 * i] | (w[2 * i + 1] << 32) for i in range(6)], dtype=np.uint64)


def test_generated_device_constants_match_reference(orc):
    one = np.zeros(36, np.uint64)
    # frobenius(v) = c1 * v  and frobenius(v^2) = c2 * v^2 recover the Fp6 coefficients (src/fp6.rs:161-185)
    v = np.zeros(36, np.uint64)
    v[12:18] = orc.R_LIMBS
    fv = orc.tower(6, "frobenius", v)[0]
    assert np.array_equal(fv[18:24], _words("K_FROB6_C1_U")) and not fv[12:18].any()
    v2 = np.zeros(36, np.uint64)
    v2[24:30] = orc.R_LIMBS
    fv2 = orc.tower(6, "frobenius", v2)[0]
    assert np.array_equal(fv2[24:30], _words("K_FROB6_C2_R")) and not fv2[30:36].any()
    # frobenius(w) = c * w recovers the Fp12 coefficient (src/fp12.rs:151-168)
    w = np.zeros(72, np.uint64)
    w[36:42] = orc.R_LIMBS
    fw = orc.tower(12, "frobenius", w)[0]
    assert np.array_equal(fw[36:42], _words("K_FROB12_C1_R")) and np.array_equal(fw[42:48], _words("K_FROB12_C1_U"))
    g1 = orc.G1.generator()[0]
    assert np.array_equal(g1[0:6], _words("K_G1_GEN_X")) and np.array_equal(g1[6:12], _words("K_G1_GEN_Y"))
    g2 = orc.G2.generator()[0]
    for i, nme in enumerate(["K_G2_GEN_X0", "K_G2_GEN_X1", "K_G2_GEN_Y0", "K_G2_GEN_Y1"]):
        assert np.array_equal(g2[6 * i:6 * i + 6], _words(nme))
    del one


def test_host_constants(orc):
    from bls12_381_b200 import constants_host as ch
    assert np.array_equal(ch.generator_projective(1).view(np.uint64), orc.G1.generator())
    assert np.array_equal(ch.generator_projective(2).view(np.uint64), orc.G2.generator())
