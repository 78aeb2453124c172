// GLV decomposition of a G1 scalar (not a reference function: the reference walks all 255 bits of the scalar,
// src/g1.rs:754-774).  BLS12-381 has q = lambda^2 + lambda + 1 with lambda = z^2 - 1 a 128-bit number and the
// endomorphism phi(x, y) = (beta x, y) = [lambda](x, y) on G1 (the reference's `endomorphism`, src/g1.rs:421-432,
// uses the other cube root, i.e. lambda^2).  Every scalar k < q is written  k = k1 + k2 lambda (mod q)  with
// |k1|, |k2| < 2^127 (Babai rounding on the basis (-lambda, 1), (1, lambda + 1)):
//     r  = round(k / lambda)            (via the 161-bit reciprocal floor(2^288 / lambda))
//     c2 = [k >= (q + 1) / 2]
//     k1 = k - r lambda - c2            k2 = r - c2 (lambda + 1)
// so that  k P = k1 P + k2 phi(P): an N-point MSM over 255-bit scalars becomes a 2N-point MSM over 127-bit
// scalars — the same number of bucket additions, HALF the windows, bucket reductions and Horner doublings.
// Checked against Python big integers on random and extreme scalars (tests/test_gpu_parity.py::test_glv_decompose).
#pragma once
#include "constants.cuh"

namespace b200 {

struct glv_parts {
  uint32_t k1[4], k2[4];  // magnitudes, < 2^127
  bool neg1, neg2;
};

B200_DEV glv_parts glv_decompose(const uint32_t s[8]) {
  glv_parts g;
  // prod = s * mu  (8 x 6 words -> 14 words), + 2^287, >> 288  ->  r (4 words)
  uint32_t prod[15];
#pragma unroll
  for (int i = 0; i < 15; i++) prod[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t carry = 0;
#pragma unroll
    for (int j = 0; j < 6; j++) {
      uint64_t t = (uint64_t)s[i] * K_GLV_MU[j] + prod[i + j] + carry;
      prod[i + j] = (uint32_t)t;
      carry = t >> 32;
    }
    prod[i + 6] = (uint32_t)carry;
  }
  {  // + 2^287 = bit 31 of word 8, ripple the carry upwards
    uint64_t t = (uint64_t)prod[8] + 0x80000000u;
    prod[8] = (uint32_t)t;
    uint64_t c = t >> 32;
#pragma unroll
    for (int i = 9; i < 14; i++) {
      t = (uint64_t)prod[i] + c;
      prod[i] = (uint32_t)t;
      c = t >> 32;
    }
  }
  uint32_t r[4] = {prod[9], prod[10], prod[11], prod[12]};
  // c2 = s >= (q+1)/2
  bool c2;
  {
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint64_t t = (uint64_t)s[i] - K_GLV_HALFQ[i] - borrow;
      borrow = (t >> 32) & 1;
    }
    c2 = borrow == 0;
  }
  // rl = r * lambda (8 words)
  uint32_t rl[8];
#pragma unroll
  for (int i = 0; i < 8; i++) rl[i] = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint64_t carry = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint64_t t = (uint64_t)r[i] * K_GLV_LAMBDA[j] + rl[i + j] + carry;
      rl[i + j] = (uint32_t)t;
      carry = t >> 32;
    }
    rl[i + 4] = (uint32_t)carry;
  }
  // k1 = s - rl - c2   (two's complement, 8 words), then sign/magnitude
  uint32_t d[8];
  {
    uint64_t borrow = c2 ? 1 : 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint64_t t = (uint64_t)s[i] - rl[i] - borrow;
      d[i] = (uint32_t)t;
      borrow = (t >> 32) & 1;
    }
    g.neg1 = borrow != 0;
  }
  if (g.neg1) {
    uint64_t c = 1;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint64_t t = (uint64_t)(~d[i]) + c;
      d[i] = (uint32_t)t;
      c = t >> 32;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) g.k1[i] = d[i];
  // k2 = r - c2 (lambda + 1)
  if (c2) {
    uint64_t borrow = 0;
    uint32_t m[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
      uint64_t t = (uint64_t)K_GLV_LAMBDA1[i] - (i < 4 ? r[i] : 0u) - borrow;
      m[i] = (uint32_t)t;
      borrow = (t >> 32) & 1;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) g.k2[i] = m[i];
    g.neg2 = (m[0] | m[1] | m[2] | m[3]) != 0;
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) g.k2[i] = r[i];
    g.neg2 = false;
  }
  if ((g.k1[0] | g.k1[1] | g.k1[2] | g.k1[3]) == 0) g.neg1 = false;
  return g;
}

}  // namespace b200
