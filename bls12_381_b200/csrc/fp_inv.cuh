// Fast modular inversion in Fp for the GPU: Kaliski's "almost Montgomery inverse" (binary extended GCD with only
// shifts, adds and subtracts on 384-bit integers, no multiplications), followed by ONE Montgomery multiplication
// by a tabulated power of two.  Not a reference algorithm: the reference inverts by Fermat, a^(p-2) = 384
// squarings + 229 multiplications (src/fp.rs:346-358).  Same value (the inverse is unique, results canonical).
//
// Why: Fermat costs 613 FpM = 187 000 IMAD.WIDE on the pipe that bounds every kernel here.  The binary GCD is
// ~600 iterations x ~150 IADD3/SHF/SEL on the ALU pipe, which the field-multiplication kernels leave mostly idle
// (ncu: ALU 18 % while the IMAD pipe is at 88 %), so inversions by some warps overlap multiplications by others.
//
// For x = a*R (Montgomery form) phase 1 returns t = x^-1 * 2^k mod p with 381 <= k <= 762; we need
// a^-1 * R = x^-1 * 2^768, i.e. t * 2^(768-k) = fp_mul(t, 2^(1152-k) mod p).  POW2[k] = 2^(1152-k) mod p is a
// 769-entry device table filled once per context (table[768] = R, table[k-1] = 2 * table[k]).
#pragma once
#include "fp.cuh"

namespace b200 {

constexpr int FP_INV_TABLE_WORDS = 769 * 12;  // 2^(1152-k) mod p, k = 0..768; owned by the b200_ctx (device memory)

static __global__ void k_fp_inv_table_init(uint32_t *pow2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  fp v = fp_one();  // 2^384 mod p = table[768]
  for (int k = 768; k >= 0; k--) {
    for (int i = 0; i < 12; i++) pow2[12 * k + i] = v.v[i];
    v = fp_dbl(v);
  }
}

// x^-1 (Montgomery in, Montgomery out); returns 0 for 0, like the callers' unwrap_or(zero)
B200_DEV fp fp_inv_fast(const fp &x, const uint32_t *pow2) {
  uint32_t u[12], v[12], r[12], s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    u[i] = fp_modw(i);
    v[i] = x.v[i];
    r[i] = 0;
    s[i] = 0;
  }
  s[0] = 1;
  int k = 0;
  bool vnz = !fp_is_zero(x);
  if (!vnz) return fp_zero();
#pragma unroll 1
  while (vnz) {
    // d = u - v  (borrow => u < v)
    uint32_t d[12], borrow;
    ptx_sub_cc(d[0], u[0], v[0]);
#pragma unroll
    for (int i = 1; i < 12; i++) ptx_subc_cc(d[i], u[i], v[i]);
    ptx_subc(borrow, 0u, 0u);               // 0xffffffff when u < v
    bool ue = (u[0] & 1u) == 0, ve = (v[0] & 1u) == 0;
    bool caseA = ue, caseB = !ue && ve, caseC = !ue && !ve && borrow == 0, caseD = !ue && !ve && borrow != 0;
    // note: u == v (both odd) gives borrow == 0, d == 0: case C sets u = 0 ... but gcd(p, x) = 1 means u == v only at 1,
    // where the classical algorithm takes the "else" branch (v = 0).  Route d == 0 to case D.
    uint32_t dz = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) dz |= d[i];
    if (dz == 0 && !ue && !ve) {
      caseC = false;
      caseD = true;
    }
    // |u - v| for case D = -(d)
    uint32_t nd[12];
    ptx_sub_cc(nd[0], 0u, d[0]);
#pragma unroll
    for (int i = 1; i < 11; i++) ptx_subc_cc(nd[i], 0u, d[i]);
    ptx_subc(nd[11], 0u, d[11]);
    // t = r + s
    uint32_t t[12];
    ptx_add_cc(t[0], r[0], s[0]);
#pragma unroll
    for (int i = 1; i < 11; i++) ptx_addc_cc(t[i], r[i], s[i]);
    ptx_addc(t[11], r[11], s[11]);
    // the value to be halved: A: u, B: v, C: u - v, D: v - u ; it replaces u (A, C) or v (B, D)
    bool into_u = caseA || caseC;
    uint32_t h[12];
#pragma unroll
    for (int i = 0; i < 12; i++) h[i] = caseA ? u[i] : caseB ? v[i] : caseC ? d[i] : nd[i];
#pragma unroll
    for (int i = 0; i < 11; i++) h[i] = __funnelshift_r(h[i], h[i + 1], 1);
    h[11] >>= 1;
    // C: r = r + s ; D: s = s + r
#pragma unroll
    for (int i = 0; i < 12; i++) {
      r[i] = caseC ? t[i] : r[i];
      s[i] = caseD ? t[i] : s[i];
    }
    // A, C: s <<= 1 ; B, D: r <<= 1
    uint32_t w[12];
#pragma unroll
    for (int i = 0; i < 12; i++) w[i] = into_u ? s[i] : r[i];
#pragma unroll
    for (int i = 11; i > 0; i--) w[i] = __funnelshift_l(w[i - 1], w[i], 1);
    w[0] <<= 1;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      s[i] = into_u ? w[i] : s[i];
      r[i] = into_u ? r[i] : w[i];
      u[i] = into_u ? h[i] : u[i];
      v[i] = into_u ? v[i] : h[i];
    }
    k++;
    uint32_t vz = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) vz |= v[i];
    vnz = vz != 0;
  }
  // r < 2p: bring into [0, p), then result = p - r  (= x^-1 * 2^k mod p)
  fp rr;
  {
    uint32_t tt[12], borrow;
    ptx_sub_cc(tt[0], r[0], fp_modw(0));
#pragma unroll
    for (int i = 1; i < 12; i++) ptx_subc_cc(tt[i], r[i], fp_modw(i));
    ptx_subc(borrow, 0u, 0u);
#pragma unroll
    for (int i = 0; i < 12; i++) rr.v[i] = borrow ? r[i] : tt[i];
  }
  fp t1 = fp_neg(rr);
  fp c;
#pragma unroll
  for (int i = 0; i < 12; i++) c.v[i] = __ldg(pow2 + 12 * k + i);
  return fp_mul_c(t1, c);
}

}  // namespace b200
