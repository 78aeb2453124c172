// Fr = GF(q), q the 255-bit order of G1/G2/Gt (the "scalar field").  Device-side arithmetic for sm_100a.
//
// Replaces src/scalar.rs of the reference (Scalar type :24, add :600, sub :582, neg :613, mul :554, square :341,
// montgomery_reduce :506, invert :408, to_bytes :284, from_bytes :256).  Same VALUE semantics: 4x64-bit
// little-endian limbs in memory, Montgomery form with R = 2^256, every stored result canonical (< q) — results are
// bit-identical to the reference because the canonical representative is unique.
//
// Same construction as fp.cuh, two thirds the size: an element lives in 8 x 32-bit registers and the product is the
// word-serial interleaved Montgomery multiplication with an even-aligned and an odd-aligned accumulator, every row
// ONE carry chain of mad.lo.cc/madc.hi.cc pairs: 8 x (8 + 8) = 128 32x32+64 multiply-adds.  In SASS (nvcc 12.9) the
// a*b_i rows are 64 IMAD.WIDE.U32[.X]; the m*q rows with immediate modulus words come out as 48 IMAD.X (low half) +
// 64 IMAD.HI.U32[.X] pairs instead of wide ones, and m*1 / the low-word bookkeeping as IADD3 on the ALU pipe.
// q = 1 (mod 2^32), so -q^-1 mod 2^32 = 0xffffffff and the Montgomery factor of a row is just -acc[0].
// Bounds: for a, b < q every running value is < a + q < 2q < 2^256, so 8 words hold it and the top chain of a row
// never carries out (the row total t + a*b_i + m*q is < 2q * 2^32 < 2^288 = 9 words from word 0).
#pragma once
#include "fp.cuh"

namespace b200 {

struct fr {
  uint32_t v[8];
};

// q, little-endian 32-bit words (src/scalar.rs:76-81)
#define FR_Q0 0x00000001u
#define FR_Q1 0xffffffffu
#define FR_Q2 0xfffe5bfeu
#define FR_Q3 0x53bda402u
#define FR_Q4 0x09a1d805u
#define FR_Q5 0x3339d808u
#define FR_Q6 0x299d7d48u
#define FR_Q7 0x73eda753u

B200_DEV uint32_t fr_modw(int i) {
  switch (i) {
    case 0: return FR_Q0;
    case 1: return FR_Q1;
    case 2: return FR_Q2;
    case 3: return FR_Q3;
    case 4: return FR_Q4;
    case 5: return FR_Q5;
    case 6: return FR_Q6;
    default: return FR_Q7;
  }
}
// R = 2^256 mod q == Scalar::one() (src/scalar.rs:159-164)
B200_DEV fr fr_one() {
  fr r = {{0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u}};
  return r;
}
// R^2 = 2^512 mod q (src/scalar.rs:167-172)
B200_DEV fr fr_r2() {
  fr r = {{0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u}};
  return r;
}
// R^3 = 2^768 mod q (src/scalar.rs:175-180)
B200_DEV fr fr_r3() {
  fr r = {{0x439b73afu, 0xc62c1807u, 0x8cf06990u, 0x1b3e0d18u, 0xc7b5f418u, 0x73d13c71u, 0xc8db33e9u, 0x6e2a5bb9u}};
  return r;
}
// 2^-1 (src/scalar.rs:183-188)
B200_DEV fr fr_two_inv() {
  fr r = {{0xffffffffu, 0x00000000u, 0x0001a401u, 0xac425bfdu, 0xf65e27fau, 0xccc627f7u, 0xd66282b7u, 0x0c1258acu}};
  return r;
}
// 2^32-th root of unity and its inverse (src/scalar.rs:200-213); GENERATOR = 7 (:100-105)
B200_DEV fr fr_root_of_unity() {
  fr r = {{0x5f0e466au, 0xb9b58d8cu, 0x1819d7ecu, 0x5b1b4c80u, 0x52a31e64u, 0x0af53ae3u, 0x19e9b27bu, 0x5bf3addau}};
  return r;
}
B200_DEV fr fr_root_of_unity_inv() {
  fr r = {{0xdcf3219au, 0x4256481au, 0x96b6cad3u, 0x45f37b7fu, 0x5f7a3b27u, 0xf9c3f1d7u, 0x658afd43u, 0x2d2fc049u}};
  return r;
}
B200_DEV fr fr_generator() {
  fr r = {{0xfffffff1u, 0x0000000eu, 0x00189c0fu, 0x17e363d3u, 0x6f8457b0u, 0xff9c5787u, 0x8fc5a8c4u, 0x35133220u}};
  return r;
}
B200_DEV fr fr_zero() {
  fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
  return r;
}

// acc[0..8) += x[0],x[2],x[4],x[6] * s : one carry chain, carry out of acc[7] left in CC.CF
B200_DEV void fr_cmad_row(uint32_t *acc, const uint32_t *x, uint32_t s) {
  ptx_mad_lo_cc(acc[0], x[0], s, acc[0]);
  ptx_madc_hi_cc(acc[1], x[0], s, acc[1]);
#pragma unroll
  for (int j = 2; j < 8; j += 2) {
    ptx_madc_lo_cc(acc[j], x[j], s, acc[j]);
    ptx_madc_hi_cc(acc[j + 1], x[j], s, acc[j + 1]);
  }
}
template <int which>
B200_DEV void fr_cmad_row_mod(uint32_t *acc, uint32_t s) {
  ptx_mad_lo_cc(acc[0], fr_modw(which), s, acc[0]);
  ptx_madc_hi_cc(acc[1], fr_modw(which), s, acc[1]);
#pragma unroll
  for (int j = 2; j < 8; j += 2) {
    ptx_madc_lo_cc(acc[j], fr_modw(which + j), s, acc[j]);
    ptx_madc_hi_cc(acc[j + 1], fr_modw(which + j), s, acc[j + 1]);
  }
}
// acc[k] = x*s + acc[k+2] (shift right by two words while accumulating), carry-in from CC.CF
B200_DEV void fr_madc_rshift_row(uint32_t *acc, const uint32_t *x, uint32_t s) {
#pragma unroll
  for (int j = 0; j < 6; j += 2) {
    ptx_madc_lo_cc(acc[j], x[j], s, acc[j + 2]);
    ptx_madc_hi_cc(acc[j + 1], x[j], s, acc[j + 3]);
  }
  ptx_madc_lo_cc(acc[6], x[6], s, 0u);
  ptx_madc_hi(acc[7], x[6], s, 0u);
}
// one Montgomery reduction step on (A at word 0, B at word 1): makes A[0] == 0
B200_DEV void fr_redc_step(uint32_t *A, uint32_t *B) {
  uint32_t m = 0u - A[0];  // A[0] * (-q^-1 mod 2^32) with -q^-1 = 0xffffffff (low word of src/scalar.rs:156)
  fr_cmad_row_mod<1>(B, m);  // no carry out (value bound)
  fr_cmad_row_mod<0>(A, m);
  ptx_addc(B[7], B[7], 0u);
}

// r = a * b * R^-1 mod q, canonical (value-identical to src/scalar.rs:554-575 + :506-550)
B200_DEV fr fr_mul(const fr &a, const fr &b) {
  uint32_t ev[8], od[8];
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    ptx_mul_lo(ev[j], a.v[j], b.v[0]);
    ptx_mul_hi(ev[j + 1], a.v[j], b.v[0]);
    ptx_mul_lo(od[j], a.v[j + 1], b.v[0]);
    ptx_mul_hi(od[j + 1], a.v[j + 1], b.v[0]);
  }
  fr_redc_step(ev, od);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    ptx_add_cc(od[0], od[0], ev[1]);
    fr_madc_rshift_row(ev, a.v + 1, b.v[i]);
    fr_cmad_row(od, a.v, b.v[i]);
    ptx_addc(ev[7], ev[7], 0u);
    fr_redc_step(od, ev);
    if (i + 1 < 8) {
      ptx_add_cc(ev[0], ev[0], od[1]);
      fr_madc_rshift_row(od, a.v + 1, b.v[i + 1]);
      fr_cmad_row(ev, a.v, b.v[i + 1]);
      ptx_addc(od[7], od[7], 0u);
      fr_redc_step(ev, od);
    }
  }
  // after 8 rows: "even role" = od (od[0] == 0), "odd role" = ev ; result = (od >> 32) + ev  (< 2q < 2^256)
  fr r;
  ptx_add_cc(r.v[0], ev[0], od[1]);
#pragma unroll
  for (int k = 1; k < 7; k++) ptx_addc_cc(r.v[k], ev[k], od[k + 1]);
  ptx_addc(r.v[7], ev[7], 0u);
  uint32_t t[8], borrow;
  ptx_sub_cc(t[0], r.v[0], fr_modw(0));
#pragma unroll
  for (int k = 1; k < 8; k++) ptx_subc_cc(t[k], r.v[k], fr_modw(k));
  ptx_subc(borrow, 0u, 0u);
#pragma unroll
  for (int k = 0; k < 8; k++) r.v[k] = borrow ? r.v[k] : t[k];
  return r;
}
// src/scalar.rs:341-370 (same canonical value)
B200_DEV fr fr_sqr(const fr &a) { return fr_mul(a, a); }
// one copy of the body per kernel; operands and result in registers (16 words in, 8 out), like fp_mul_c
static __device__ __noinline__ fr fr_mul_c(fr a, fr b) { return fr_mul(a, b); }

// src/scalar.rs:600-610  (a + b < 2q < 2^256: no carry out of the top word)
B200_DEV fr fr_add(const fr &a, const fr &b) {
  fr r;
  ptx_add_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
  for (int k = 1; k < 7; k++) ptx_addc_cc(r.v[k], a.v[k], b.v[k]);
  ptx_addc(r.v[7], a.v[7], b.v[7]);
  uint32_t t[8], borrow;
  ptx_sub_cc(t[0], r.v[0], fr_modw(0));
#pragma unroll
  for (int k = 1; k < 8; k++) ptx_subc_cc(t[k], r.v[k], fr_modw(k));
  ptx_subc(borrow, 0u, 0u);
#pragma unroll
  for (int k = 0; k < 8; k++) r.v[k] = borrow ? r.v[k] : t[k];
  return r;
}
// src/scalar.rs:582-597
B200_DEV fr fr_sub(const fr &a, const fr &b) {
  fr r;
  uint32_t borrow;
  ptx_sub_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
  for (int k = 1; k < 8; k++) ptx_subc_cc(r.v[k], a.v[k], b.v[k]);
  ptx_subc(borrow, 0u, 0u);  // 0xffffffff when a < b
  ptx_add_cc(r.v[0], r.v[0], fr_modw(0) & borrow);
#pragma unroll
  for (int k = 1; k < 7; k++) ptx_addc_cc(r.v[k], r.v[k], fr_modw(k) & borrow);
  ptx_addc(r.v[7], r.v[7], fr_modw(7) & borrow);
  return r;
}
B200_DEV bool fr_is_zero(const fr &a) {
  uint32_t o = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) o |= a.v[k];
  return o == 0;
}
// src/scalar.rs:613-627
B200_DEV fr fr_neg(const fr &a) {
  fr r;
  uint32_t mask = fr_is_zero(a) ? 0u : 0xffffffffu;
  ptx_sub_cc(r.v[0], fr_modw(0), a.v[0]);
#pragma unroll
  for (int k = 1; k < 7; k++) ptx_subc_cc(r.v[k], fr_modw(k), a.v[k]);
  ptx_subc(r.v[7], fr_modw(7), a.v[7]);
#pragma unroll
  for (int k = 0; k < 8; k++) r.v[k] &= mask;
  return r;
}
B200_DEV fr fr_dbl(const fr &a) { return fr_add(a, a); }  // src/scalar.rs:249-252

// a^e for a 64-bit exponent, MSB first (src/scalar.rs:392-404 on one limb)
B200_DEV fr fr_pow_u64(const fr &a, unsigned long long e) {
  fr res = fr_one();
#pragma unroll 1
  for (int i = 63; i >= 0; i--) {
    res = fr_mul_c(res, res);
    if ((e >> i) & 1) res = fr_mul_c(res, a);
  }
  return res;
}
// a^(q-2); 0 -> 0.  The reference uses a fixed addition chain (src/scalar.rs:408-503) for the same power — its
// test_invert_is_pow (:1184) pins invert() == pow_vartime(q - 2); the inverse is unique, so the limbs are identical.
B200_DEV fr fr_inv(const fr &a) {
  fr res = fr_one();
#pragma unroll 1
  for (int w = 7; w >= 0; w--) {
    // q - 2: the low word 0x00000001 - 2 borrows from word 1 (0xffffffff); the other words are those of q
    uint32_t e = w == 0 ? 0xffffffffu : (w == 1 ? 0xfffffffeu : fr_modw(w));
#pragma unroll 1
    for (int i = 31; i >= 0; i--) {
      res = fr_mul_c(res, res);
      if ((e >> i) & 1) res = fr_mul_c(res, a);
    }
  }
  return res;
}

// x < q ?  (canonical check of src/scalar.rs:266-275)
B200_DEV bool fr_is_canonical(const fr &x) {
  uint32_t t, borrow;
  ptx_sub_cc(t, x.v[0], fr_modw(0));
#pragma unroll
  for (int k = 1; k < 8; k++) ptx_subc_cc(t, x.v[k], fr_modw(k));
  ptx_subc(borrow, 0u, 0u);
  return borrow != 0;
}
// Montgomery -> canonical integer (src/scalar.rs:284-296: montgomery_reduce of the limbs) == a * 1 * R^-1
B200_DEV fr fr_from_mont(const fr &a) {
  fr one_raw = fr_zero();
  one_raw.v[0] = 1;
  return fr_mul(a, one_raw);
}
// canonical integer (< q) -> Montgomery (src/scalar.rs:278: tmp *= R2)
B200_DEV fr fr_to_mont(const fr &a) { return fr_mul(a, fr_r2()); }

// x mod q for any 256-bit x (2^256 < 3q: at most two subtractions)
B200_DEV fr fr_reduce_256(const fr &x) {
  fr r = x;
#pragma unroll 1
  for (int it = 0; it < 2; it++) {
    uint32_t t[8], borrow;
    ptx_sub_cc(t[0], r.v[0], fr_modw(0));
#pragma unroll
    for (int k = 1; k < 8; k++) ptx_subc_cc(t[k], r.v[k], fr_modw(k));
    ptx_subc(borrow, 0u, 0u);
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = borrow ? r.v[k] : t[k];
  }
  return r;
}
// Scalar::from_bytes_wide / from_u512 (src/scalar.rs:300-331): lo + hi * 2^256 as a field element, Montgomery form:
// lo * R2 + hi * R3.  The reference feeds the raw 256-bit halves to its (wide-product) multiplication; the interleaved
// multiplication here wants operands < q, so the halves are reduced first — the field element is the same.
B200_DEV fr fr_from_wide(const fr &lo, const fr &hi) {
  return fr_add(fr_mul(fr_reduce_256(lo), fr_r2()), fr_mul(fr_reduce_256(hi), fr_r3()));
}

// ---- global memory <-> registers: 8 consecutive little-endian 32-bit words (== Scalar([u64; 4])), 32 B = one sector
B200_DEV fr fr_load(const void *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = q[0], b = q[1];
  fr r = {{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
  return r;
}
B200_DEV fr fr_load_ro(const void *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  fr r = {{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
  return r;
}
B200_DEV void fr_store(void *p, const fr &r) {
  uint4 *q = reinterpret_cast<uint4 *>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

}  // namespace b200
