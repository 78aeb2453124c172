// Fp2 = Fp[u]/(u^2+1).  Replaces src/fp2.rs (mul :205, square :182, add/sub/neg :224-243,
// mul_by_nonresidue :156, conjugate/frobenius :141-153, invert :300).  Values are canonical, hence
// bit-identical to the reference; the multiplication is 3-mul Karatsuba over the register-resident
// fp_mul instead of the reference's two interleaved sum_of_products (same field element).
#pragma once
#include "fp.cuh"

namespace b200 {

struct fp2 {
  fp c0, c1;
};

B200_DEV fp2 fp2_zero() { return fp2{fp_zero(), fp_zero()}; }
B200_DEV fp2 fp2_one() { return fp2{fp_one(), fp_zero()}; }
B200_DEV fp2 fp2_add(const fp2 &a, const fp2 &b) { return fp2{fp_add(a.c0, b.c0), fp_add(a.c1, b.c1)}; }
B200_DEV fp2 fp2_sub(const fp2 &a, const fp2 &b) { return fp2{fp_sub(a.c0, b.c0), fp_sub(a.c1, b.c1)}; }
B200_DEV fp2 fp2_neg(const fp2 &a) { return fp2{fp_neg(a.c0), fp_neg(a.c1)}; }
B200_DEV fp2 fp2_dbl(const fp2 &a) { return fp2{fp_dbl(a.c0), fp_dbl(a.c1)}; }
B200_DEV fp2 fp2_conj(const fp2 &a) { return fp2{a.c0, fp_neg(a.c1)}; }
B200_DEV bool fp2_is_zero(const fp2 &a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
B200_DEV bool fp2_eq(const fp2 &a, const fp2 &b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }
// (a + bu)(1 + u) = (a - b) + (a + b)u      (src/fp2.rs:156-166)
B200_DEV fp2 fp2_mul_by_nonresidue(const fp2 &a) { return fp2{fp_sub(a.c0, a.c1), fp_add(a.c0, a.c1)}; }
B200_DEV fp2 fp2_mul(const fp2 &a, const fp2 &b) {
  fp t0 = fp_mul(a.c0, b.c0);
  fp t1 = fp_mul(a.c1, b.c1);
  fp s = fp_mul(fp_add(a.c0, a.c1), fp_add(b.c0, b.c1));
  return fp2{fp_sub(t0, t1), fp_sub(fp_sub(s, t0), t1)};
}
// complex squaring (src/fp2.rs:182-203)
B200_DEV fp2 fp2_sqr(const fp2 &a) {
  fp s = fp_add(a.c0, a.c1), d = fp_sub(a.c0, a.c1), t = fp_dbl(a.c0);
  return fp2{fp_mul(s, d), fp_mul(t, a.c1)};
}
B200_DEV fp2 fp2_mul_fp(const fp2 &a, const fp &k) { return fp2{fp_mul(a.c0, k), fp_mul(a.c1, k)}; }
// src/fp2.rs:300-320 ; returns 0 for 0 (== CtOption::unwrap_or(zero) at the call sites)
B200_DEV fp2 fp2_inv(const fp2 &a) {
  fp t = fp_inv(fp_add(fp_sqr(a.c0), fp_sqr(a.c1)));
  return fp2{fp_mul(a.c0, t), fp_mul(a.c1, fp_neg(t))};
}
B200_DEV fp2 fp2_select(const fp2 &a, const fp2 &b, bool choose_b) {
  return fp2{fp_select(a.c0, b.c0, choose_b), fp_select(a.c1, b.c1, choose_b)};
}
B200_DEV fp2 fp2_load(const void *p) {
  const char *q = reinterpret_cast<const char *>(p);
  return fp2{fp_load(q), fp_load(q + 48)};
}
B200_DEV fp2 fp2_load_ro(const void *p) {
  const char *q = reinterpret_cast<const char *>(p);
  return fp2{fp_load_ro(q), fp_load_ro(q + 48)};
}
B200_DEV void fp2_store(void *p, const fp2 &a) {
  char *q = reinterpret_cast<char *>(p);
  fp_store(q, a.c0);
  fp_store(q + 48, a.c1);
}

// Called (non-inlined) Fp2 mul/sqr: ONE copy of the body per kernel instead of one per call site.
// Operands and result travel BY VALUE in registers (48 words in, 24 out — checked in SASS: no STL/LDL,
// 0-byte frame), and the three/two Fp products inside are calls into the single fp_mul_c body.
// Everything at Fp2 level and above (G2, Fp6, Fp12, pairing) goes through these.
#define B200_NOINL static __device__ __noinline__
// Implementations of the called Fp2 multiply, chosen per translation unit (all return the same
// canonical element, bit-identical to src/fp2.rs:205-222; measured on B200, rounds 1-2):
//   default            Karatsuba with LAZY reduction: three unreduced 768-bit products, two Montgomery
//                      reductions (753 IMAD instead of 915).  Best where the IMAD pipe is saturated
//                      (G2 MSM bucket kernel: -6 % time).  c0 = a0 b0 - a1 b1 (+ p 2^384 if negative),
//                      c1 = (a0+a1)(b0+b1) - a0 b0 - a1 b1 (never negative; operand sums < 2p not reduced).
//   B200_FP2_KCALL     Karatsuba over three calls of fp_mul_c: shortest dependent chains; best for the
//                      latency-bound pairing kernels (2 warps per SMSP), where the lazy variant is 9 % slower.
//   B200_FP2_KINLINE   Karatsuba with the three/two Fp products inlined side by side (more ILP for ptxas).
//   B200_FP2_LAZY3     the lazy variant with its three wide products and two reductions row-alternated: the MSM unit's
//                      choice (G2 MSM 2^20: 30.1 vs 31.3 ms, round 2).  Row-alternated Karatsuba products inside the
//                      pairing kernels (dual / triple streams) measured SLOWER (63.5 / 67.4 vs 59.3 ms) and were removed.
#if defined(B200_FP2_KCALL)
B200_NOINL fp2 fp2_mul_c(fp2 a, fp2 b) {
  fp t0 = fp_mul_c(a.c0, b.c0);
  fp t1 = fp_mul_c(a.c1, b.c1);
  fp s = fp_mul_c(fp_add(a.c0, a.c1), fp_add(b.c0, b.c1));
  return fp2{fp_sub(t0, t1), fp_sub(fp_sub(s, t0), t1)};
}
B200_NOINL fp2 fp2_sqr_c(fp2 a) {
  fp s = fp_add(a.c0, a.c1), d = fp_sub(a.c0, a.c1), t = fp_dbl(a.c0);
  return fp2{fp_mul_c(s, d), fp_mul_c(t, a.c1)};
}
#elif defined(B200_FP2_LAZY3)
// the lazy-reduction multiply with its three wide products and its two reductions row-alternated
// (fp_mul_wide_triple, fp_redc_wide_dual): used by the MSM unit (G2 bucket kernel)
B200_NOINL fp2 fp2_mul_c(fp2 a, fp2 b) {
  fpw3 w = fp_mul_wide_triple(a.c0, b.c0, a.c1, b.c1, fp_add_nr(a.c0, a.c1), fp_add_nr(b.c0, b.c1));
  fp_pair r = fp_redc_wide_dual(fpw_sub(w.w2, fpw_add(w.w0, w.w1)), fpw_sub_mod(w.w0, w.w1));
  return fp2{r.r1, r.r0};
}
B200_NOINL fp2 fp2_sqr_c(fp2 a) {
  fp s = fp_add(a.c0, a.c1), d = fp_sub(a.c0, a.c1), t = fp_dbl(a.c0);
  fp_pair p = fp_mul_dual(s, d, t, a.c1);
  return fp2{p.r0, p.r1};
}
#elif defined(B200_FP2_KINLINE)
B200_NOINL fp2 fp2_mul_c(fp2 a, fp2 b) { return fp2_mul(a, b); }
B200_NOINL fp2 fp2_sqr_c(fp2 a) { return fp2_sqr(a); }
#else
B200_NOINL fp2 fp2_mul_c(fp2 a, fp2 b) {
  fpw w0 = fp_mul_wide(a.c0, b.c0);
  fpw w1 = fp_mul_wide(a.c1, b.c1);
  fpw w2 = fp_mul_wide(fp_add_nr(a.c0, a.c1), fp_add_nr(b.c0, b.c1));
  fp c1 = fp_redc_wide(fpw_sub(w2, fpw_add(w0, w1)));
  fp c0 = fp_redc_wide(fpw_sub_mod(w0, w1));
  return fp2{c0, c1};
}
B200_NOINL fp2 fp2_sqr_c(fp2 a) {
  fp s = fp_add(a.c0, a.c1), d = fp_sub(a.c0, a.c1), t = fp_dbl(a.c0);
  return fp2{fp_mul_c(s, d), fp_mul_c(t, a.c1)};
}
#endif
B200_NOINL fp fp_inv_c(fp a) { return fp_inv(a); }
B200_DEV fp2 M2(const fp2 &a, const fp2 &b) { return fp2_mul_c(a, b); }
B200_DEV fp2 S2(const fp2 &a) { return fp2_sqr_c(a); }
B200_DEV void fp_mul_ni(fp *r, const fp *a, const fp *b) { *r = fp_mul_c(*a, *b); }
B200_DEV fp2 fp2_inv_ni(const fp2 &a) {
  fp t = fp_inv_c(fp_add(fp_mul_c(a.c0, a.c0), fp_mul_c(a.c1, a.c1)));
  return fp2{fp_mul_c(a.c0, t), fp_mul_c(a.c1, fp_neg(t))};
}

// ---- uniform names so the curve templates (curve.cuh) work over Fp (G1) and Fp2 (G2)
B200_DEV fp f_add(const fp &a, const fp &b) { return fp_add(a, b); }
B200_DEV fp f_sub(const fp &a, const fp &b) { return fp_sub(a, b); }
B200_DEV fp f_mul(const fp &a, const fp &b) { return fp_mul_c(a, b); }
B200_DEV fp f_sqr(const fp &a) { return fp_sqr_c(a); }
B200_DEV fp f_neg(const fp &a) { return fp_neg(a); }
B200_DEV fp f_dbl(const fp &a) { return fp_dbl(a); }
B200_DEV fp f_inv(const fp &a) { return fp_inv(a); }
B200_DEV bool f_is_zero(const fp &a) { return fp_is_zero(a); }
B200_DEV bool f_eq(const fp &a, const fp &b) { return fp_eq(a, b); }
B200_DEV fp f_select(const fp &a, const fp &b, bool c) { return fp_select(a, b, c); }
B200_DEV void f_store(void *p, const fp &a) { fp_store(p, a); }
B200_DEV fp2 f_add(const fp2 &a, const fp2 &b) { return fp2_add(a, b); }
B200_DEV fp2 f_sub(const fp2 &a, const fp2 &b) { return fp2_sub(a, b); }
B200_DEV fp2 f_mul(const fp2 &a, const fp2 &b) { return M2(a, b); }
B200_DEV fp2 f_sqr(const fp2 &a) { return S2(a); }
B200_DEV fp2 f_neg(const fp2 &a) { return fp2_neg(a); }
B200_DEV fp2 f_dbl(const fp2 &a) { return fp2_dbl(a); }
B200_DEV fp2 f_inv(const fp2 &a) { return fp2_inv_ni(a); }
B200_DEV bool f_is_zero(const fp2 &a) { return fp2_is_zero(a); }
B200_DEV bool f_eq(const fp2 &a, const fp2 &b) { return fp2_eq(a, b); }
B200_DEV fp2 f_select(const fp2 &a, const fp2 &b, bool c) { return fp2_select(a, b, c); }
B200_DEV void f_store(void *p, const fp2 &a) { fp2_store(p, a); }

template <class F> struct field_traits;
template <> struct field_traits<fp> {
  static constexpr int bytes = 48;
  static B200_DEV fp zero() { return fp_zero(); }
  static B200_DEV fp one() { return fp_one(); }
  static B200_DEV fp load(const void *p) { return fp_load(p); }
  static B200_DEV fp load_ro(const void *p) { return fp_load_ro(p); }
  // 3b = 12 for E: y^2 = x^3 + 4   (src/g1.rs:597-601 computes the same by repeated adds)
  static B200_DEV fp mul_by_3b(const fp &a) {
    fp t = fp_dbl(fp_dbl(a));       // 4a
    return fp_add(fp_dbl(t), t);    // 12a
  }
};
template <> struct field_traits<fp2> {
  static constexpr int bytes = 96;
  static B200_DEV fp2 zero() { return fp2_zero(); }
  static B200_DEV fp2 one() { return fp2_one(); }
  static B200_DEV fp2 load(const void *p) { return fp2_load(p); }
  static B200_DEV fp2 load_ro(const void *p) { return fp2_load_ro(p); }
  // 3b' = 12(1+u) for E': y^2 = x^3 + 4(1+u).  The reference multiplies by the constant B3 with a
  // full Fp2 mul (src/g2.rs:650-652); the same canonical element is 12 * mul_by_nonresidue(a).
  static B200_DEV fp2 mul_by_3b(const fp2 &a) {
    fp2 n = fp2_mul_by_nonresidue(a);
    fp2 t = fp2_dbl(fp2_dbl(n));
    return fp2_add(fp2_dbl(t), t);
  }
};

}  // namespace b200
