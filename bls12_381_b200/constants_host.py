"""Host-side constants (Montgomery limbs) computed from the published BLS12-381 parameters — the same
values the reference hard-codes at src/g1.rs:199-214 and src/g2.rs:212-247 (checked in tests/test_constants.py)."""
import numpy as np

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
Q = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
R = (1 << 384) % P
G1_GEN = (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
          0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)
G2_GEN = ((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
           0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
          (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
           0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be))


def fp_mont(v):
    m = v * R % P
    return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(6)]


def generator_projective(k):
    """(1, 18k) int64 view of G{k}Projective::generator() limbs (x, y, z = one)"""
    if k == 1:
        l = fp_mont(G1_GEN[0]) + fp_mont(G1_GEN[1]) + fp_mont(1)
    else:
        l = (fp_mont(G2_GEN[0][0]) + fp_mont(G2_GEN[0][1]) + fp_mont(G2_GEN[1][0]) + fp_mont(G2_GEN[1][1]) +
             fp_mont(1) + fp_mont(0))
    return np.array(l, dtype=np.uint64).view(np.int64).reshape(1, -1)
