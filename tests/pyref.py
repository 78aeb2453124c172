"""Second, independent oracle: textbook BLS12-381 arithmetic on Python big integers (affine
coordinates, no Montgomery form, no shared formulas with oracle/ or the CUDA code).  Used to
cross-check the C++ oracle at the group level (SURVEY.md §8c "second, independent oracle")."""
import numpy as np

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
Q = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
R = (1 << 384) % P
RINV = pow(R, -1, P)


def limbs_to_int(l):
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


def int_to_limbs(v, n=6):
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)], dtype=np.uint64)


def from_mont(l):
    return limbs_to_int(l) * RINV % P


def to_mont(v):
    return int_to_limbs(v * R % P)


# ---- Fp2 as (a, b) = a + b u, u^2 = -1
def f2_add(x, y): return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
def f2_sub(x, y): return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)
def f2_mul(x, y): return ((x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
def f2_inv(x):
    d = pow(x[0] * x[0] + x[1] * x[1], -1, P)
    return (x[0] * d % P, -x[1] * d % P)


class Curve:
    """y^2 = x^3 + b over Fp (k=1) or Fp2 (k=2); points are None (infinity) or (x, y)."""

    def __init__(self, k):
        self.k = k
        if k == 1:
            self.b = 4
            self.add_, self.sub_, self.mul_ = (lambda a, b: (a + b) % P), (lambda a, b: (a - b) % P), (lambda a, b: a * b % P)
            self.inv_ = lambda a: pow(a, -1, P)
            self.zero = 0
        else:
            self.b = (4, 4)
            self.add_, self.sub_, self.mul_, self.inv_ = f2_add, f2_sub, f2_mul, f2_inv
            self.zero = (0, 0)

    def is_on_curve(self, p):
        if p is None:
            return True
        x, y = p
        return self.mul_(y, y) == self.add_(self.mul_(self.mul_(x, x), x), self.b)

    def neg(self, p):
        return None if p is None else (p[0], self.sub_(self.zero, p[1]))

    def add(self, p, q):
        if p is None: return q
        if q is None: return p
        (x1, y1), (x2, y2) = p, q
        if x1 == x2:
            if y1 != y2 or y1 == self.zero:
                return None
            xx = self.mul_(x1, x1)
            lam = self.mul_(self.add_(self.add_(xx, xx), xx), self.inv_(self.add_(y1, y1)))
        else:
            lam = self.mul_(self.sub_(y2, y1), self.inv_(self.sub_(x2, x1)))
        x3 = self.sub_(self.sub_(self.mul_(lam, lam), x1), x2)
        y3 = self.sub_(self.mul_(lam, self.sub_(x1, x3)), y1)
        return (x3, y3)

    def mul(self, p, k):
        acc = None
        while k:
            if k & 1:
                acc = self.add(acc, p)
            p = self.add(p, p)
            k >>= 1
        return acc

    # ---- conversion from/to the Montgomery-limb arrays used everywhere else
    def from_affine_limbs(self, xy, inf=0):
        if inf:
            return None
        xy = np.asarray(xy, dtype=np.uint64).reshape(-1)
        if self.k == 1:
            return (from_mont(xy[0:6]), from_mont(xy[6:12]))
        return ((from_mont(xy[0:6]), from_mont(xy[6:12])), (from_mont(xy[12:18]), from_mont(xy[18:24])))

    def to_affine_limbs(self, p):
        if p is None:
            out = np.zeros(12 * self.k, np.uint64)
            out[6 * self.k:6 * self.k + 6] = int_to_limbs(R)
            return out, 1
        if self.k == 1:
            return np.concatenate([to_mont(p[0]), to_mont(p[1])]), 0
        return np.concatenate([to_mont(p[0][0]), to_mont(p[0][1]), to_mont(p[1][0]), to_mont(p[1][1])]), 0


E1 = Curve(1)
E2 = Curve(2)
G1_GEN = (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
          0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)
G2_GEN = ((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
           0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
          (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
           0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be))
