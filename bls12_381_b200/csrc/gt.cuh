// Gt group operations that are more than one tower op (`Gt: Group<Scalar = Scalar>`).
// Add = Fp12 mul, Neg = conjugate, double = square are b200_tower_op / b200_fp12_product_dev; this file adds
// `&Gt * &Scalar` (src/pairings.rs:296-323): double-and-add over the 32-byte little-endian scalar, most significant bit
// first, the leading bit (always unset for a canonical scalar) skipped.  The reference computes acc + self at every bit
// and selects; the value is the same when the multiplication is simply skipped on a 0 bit (variable time, like the
// rest of the GPU path).  The generic Fp12 square is used (not the cyclotomic one), so the result is the reference's
// for ANY Fp12 input, not only for elements of the cyclotomic subgroup.
#pragma once
#include "tower.cuh"

namespace b200 {

B200_DEV void gt_mul_scalar(fp12 *acc, const fp12 *g, const uint32_t by[8]) {
  fp12_set_one(acc);
#pragma unroll 1
  for (int bit = 254; bit >= 0; bit--) {
    fp12_sqr(acc, acc);
    if ((by[bit >> 5] >> (bit & 31)) & 1) fp12_mul(acc, acc, g);
  }
}

static __global__ void __launch_bounds__(64, 4) k_gt_mul_batch(const char *g, const uint32_t *s, size_t n, char *out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t by[8];
#pragma unroll
  for (int k = 0; k < 8; k++) by[k] = __ldg(s + 8 * i + k);
  fp12 x, acc;
  fp12_load(&x, g + 576 * i);
  gt_mul_scalar(&acc, &x, by);
  fp12_store(out + 576 * i, &acc);
}

}  // namespace b200
