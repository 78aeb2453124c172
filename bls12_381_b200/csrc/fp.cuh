// Fp = GF(p), p the 381-bit BLS12-381 base-field modulus.  Device-side arithmetic for sm_100a.
//
// Replaces src/fp.rs of the reference (Fp type :15, add :382, neg :396, sub :421, mul :565, square :613,
// montgomery_reduce :487, invert :346).  Same VALUE semantics: 6x64-bit little-endian limbs in memory,
// Montgomery form with R = 2^384, every stored result canonical (< p), so results are bit-identical
// to the reference (the canonical representative is unique).
//
// NOT the same algorithm: the reference does a 6x6 64-bit schoolbook product followed by a separate
// HAC-14.32 reduction.  A B200 SM has no 64-bit multiplier: the integer datapath is the 32x32+64->64
// IMAD.WIDE.U32 on the FMA pipe.  So an element lives in 12 x 32-bit registers and mul is a word-serial
// interleaved (CIOS-style) Montgomery product in which the partial products are split into an
// "even-aligned" and an "odd-aligned" accumulator so every row is ONE carry chain of
// mad.lo.cc/madc.hi.cc pairs that ptxas fuses into IMAD.WIDE.U32(.X) — 12 wide-IMADs per row for
// a*b_i, 12 for m*p, 1 for m: 300 IMADs per multiplication, no carry ripple between lo/hi halves.
#pragma once
#include <cstdint>

namespace b200 {

#define B200_DEV __device__ __forceinline__

struct fp {
  uint32_t v[12];
};

// p, little-endian 32-bit words  (src/fp.rs:70-77)
#define FP_P0 0xffffaaabu
#define FP_P1 0xb9feffffu
#define FP_P2 0xb153ffffu
#define FP_P3 0x1eabfffeu
#define FP_P4 0xf6b0f624u
#define FP_P5 0x6730d2a0u
#define FP_P6 0xf38512bfu
#define FP_P7 0x64774b84u
#define FP_P8 0x434bacd7u
#define FP_P9 0x4b1ba7b6u
#define FP_P10 0x397fe69au
#define FP_P11 0x1a0111eau
// -p^{-1} mod 2^32 (low word of src/fp.rs:80)
#define FP_INV32 0xfffcfffdu

__device__ __constant__ const uint32_t FP_MOD[12] = {FP_P0, FP_P1, FP_P2, FP_P3, FP_P4,  FP_P5,
                                                    FP_P6, FP_P7, FP_P8, FP_P9, FP_P10, FP_P11};

B200_DEV uint32_t fp_modw(int i) {
  switch (i) {
    case 0: return FP_P0;
    case 1: return FP_P1;
    case 2: return FP_P2;
    case 3: return FP_P3;
    case 4: return FP_P4;
    case 5: return FP_P5;
    case 6: return FP_P6;
    case 7: return FP_P7;
    case 8: return FP_P8;
    case 9: return FP_P9;
    case 10: return FP_P10;
    default: return FP_P11;
  }
}

// R = 2^384 mod p  == Fp::one()  (src/fp.rs:83-90)
B200_DEV fp fp_one() {
  fp r = {{0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
           0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u}};
  return r;
}
B200_DEV fp fp_zero() {
  fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = 0;
  return r;
}

// ---- PTX carry-chain primitives (CC.CF lives between adjacent volatile asm statements)
#ifdef B200_HOST_EMUL  // CPU test harness (tests/emul/): bit-exact C models of the same instructions
#include "emul_ptx.h"
#else
B200_DEV void ptx_add_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
B200_DEV void ptx_addc_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
B200_DEV void ptx_addc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
B200_DEV void ptx_sub_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
B200_DEV void ptx_subc_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
B200_DEV void ptx_subc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
B200_DEV void ptx_mul_lo(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
B200_DEV void ptx_mul_hi(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
B200_DEV void ptx_mad_lo_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
B200_DEV void ptx_madc_lo_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
B200_DEV void ptx_madc_hi_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
B200_DEV void ptx_madc_hi(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
#endif

// acc[0..12) += x[0],x[2],..,x[10] (stride-2 words of a 12-word operand starting at x) * s ; one
// carry chain; the carry out of acc[11] is left in CC.CF for the caller.
B200_DEV void fp_cmad_row(uint32_t *acc, const uint32_t *x, uint32_t s) {
  ptx_mad_lo_cc(acc[0], x[0], s, acc[0]);
  ptx_madc_hi_cc(acc[1], x[0], s, acc[1]);
#pragma unroll
  for (int j = 2; j < 12; j += 2) {
    ptx_madc_lo_cc(acc[j], x[j], s, acc[j]);
    ptx_madc_hi_cc(acc[j + 1], x[j], s, acc[j + 1]);
  }
}
// same with the modulus words as immediates: which = 0 (even words p0,p2,..) or 1 (odd words p1,p3,..)
template <int which>
B200_DEV void fp_cmad_row_mod(uint32_t *acc, uint32_t s) {
  ptx_mad_lo_cc(acc[0], fp_modw(which), s, acc[0]);
  ptx_madc_hi_cc(acc[1], fp_modw(which), s, acc[1]);
#pragma unroll
  for (int j = 2; j < 12; j += 2) {
    ptx_madc_lo_cc(acc[j], fp_modw(which + j), s, acc[j]);
    ptx_madc_hi_cc(acc[j + 1], fp_modw(which + j), s, acc[j + 1]);
  }
}
// acc[k] = x*s + acc[k+2] (shift right by two words while accumulating), carry-in from CC.CF
B200_DEV void fp_madc_rshift_row(uint32_t *acc, const uint32_t *x, uint32_t s) {
#pragma unroll
  for (int j = 0; j < 10; j += 2) {
    ptx_madc_lo_cc(acc[j], x[j], s, acc[j + 2]);
    ptx_madc_hi_cc(acc[j + 1], x[j], s, acc[j + 3]);
  }
  ptx_madc_lo_cc(acc[10], x[10], s, 0u);
  ptx_madc_hi(acc[11], x[10], s, 0u);
}
// one Montgomery reduction step on (A at word 0, B at word 1): makes A[0] == 0
B200_DEV void fp_redc_step(uint32_t *A, uint32_t *B) {
  uint32_t m = A[0] * FP_INV32;
  fp_cmad_row_mod<1>(B, m);  // no carry out (value bound)
  fp_cmad_row_mod<0>(A, m);
  ptx_addc(B[11], B[11], 0u);
}

// r = a * b * R^-1 mod p, canonical.   (value-identical to src/fp.rs:565-609)
B200_DEV fp fp_mul(const fp &a, const fp &b) {
  uint32_t ev[12], od[12];
  // row 0: plain products
#pragma unroll
  for (int j = 0; j < 12; j += 2) {
    ptx_mul_lo(ev[j], a.v[j], b.v[0]);
    ptx_mul_hi(ev[j + 1], a.v[j], b.v[0]);
    ptx_mul_lo(od[j], a.v[j + 1], b.v[0]);
    ptx_mul_hi(od[j + 1], a.v[j + 1], b.v[0]);
  }
  fp_redc_step(ev, od);
#pragma unroll
  for (int i = 1; i < 12; i += 2) {
    // odd row: A = od (word aligned after the shift), B = ev (shifted in)
    ptx_add_cc(od[0], od[0], ev[1]);
    fp_madc_rshift_row(ev, a.v + 1, b.v[i]);
    fp_cmad_row(od, a.v, b.v[i]);
    ptx_addc(ev[11], ev[11], 0u);
    fp_redc_step(od, ev);
    if (i + 1 < 12) {
      // even row: roles swapped back
      ptx_add_cc(ev[0], ev[0], od[1]);
      fp_madc_rshift_row(od, a.v + 1, b.v[i + 1]);
      fp_cmad_row(ev, a.v, b.v[i + 1]);
      ptx_addc(od[11], od[11], 0u);
      fp_redc_step(ev, od);
    }
  }
  // after 12 rows: "even role" = od (od[0]==0), "odd role" = ev ; result = (od >> 32) + ev  (< 2p)
  fp r;
  ptx_add_cc(r.v[0], ev[0], od[1]);
#pragma unroll
  for (int k = 1; k < 11; k++) ptx_addc_cc(r.v[k], ev[k], od[k + 1]);
  ptx_addc(r.v[11], ev[11], 0u);
  // conditional subtract p
  uint32_t t[12], borrow;
  ptx_sub_cc(t[0], r.v[0], fp_modw(0));
#pragma unroll
  for (int k = 1; k < 12; k++) ptx_subc_cc(t[k], r.v[k], fp_modw(k));
  ptx_subc(borrow, 0u, 0u);
#pragma unroll
  for (int k = 0; k < 12; k++) r.v[k] = borrow ? r.v[k] : t[k];
  return r;
}
B200_DEV fp fp_sqr(const fp &a) { return fp_mul(a, a); }

// Called (non-inlined) multiply.  nvcc passes the two 12-word operands and the result in registers
// (no stack traffic — checked in SASS: CALL.REL.NOINC, 0-byte frame), so a group operation becomes
// ~a dozen calls into ONE 6 KB body that stays resident in the instruction cache, instead of a
// 70+ KB fully inlined loop body (ncu round 1: `no_instruction` was the top stall of the MSM kernel).
static __device__ __noinline__ fp fp_mul_c(fp a, fp b) { return fp_mul(a, b); }

// ---------------------------------------------------------------------------------------------------
// TWO independent products with their rows alternated: (a*b, c*d).  Each product is the dependent carry-chain stream
// of fp_mul; one warp alone cannot keep the multiplier busy with it (ncu, pairing kernels: the `wait` stall — fixed-
// latency dependency — is 3.3-3.6 cycles per issue at 2 warps per scheduler).  Emitting the chains of two products
// next to each other (a whole chain at a time: PTX has a single CC.CF, so chains cannot be interleaved instruction by
// instruction in the source — ptxas renames the carries into predicate registers and overlaps adjacent independent
// chains) gives the scheduler two independent streams per warp.  Same values as two fp_mul calls.
// Used by the squaring of the B200_FP2_LAZY3 variant (fp2.cuh).
struct fp_pair {
  fp r0, r1;
};
B200_DEV fp fp_mul_tail(const uint32_t *ev, const uint32_t *od) {
  fp r;
  ptx_add_cc(r.v[0], ev[0], od[1]);
#pragma unroll
  for (int k = 1; k < 11; k++) ptx_addc_cc(r.v[k], ev[k], od[k + 1]);
  ptx_addc(r.v[11], ev[11], 0u);
  uint32_t t[12], borrow;
  ptx_sub_cc(t[0], r.v[0], fp_modw(0));
#pragma unroll
  for (int k = 1; k < 12; k++) ptx_subc_cc(t[k], r.v[k], fp_modw(k));
  ptx_subc(borrow, 0u, 0u);
#pragma unroll
  for (int k = 0; k < 12; k++) r.v[k] = borrow ? r.v[k] : t[k];
  return r;
}
// one interleaved-Montgomery row of fp_mul on (A, B) = (the accumulator that is word-aligned after the shift, the other)
B200_DEV void fp_mul_row_acc(uint32_t *A, uint32_t *B, const fp &x, uint32_t y) {
  ptx_add_cc(A[0], A[0], B[1]);
  fp_madc_rshift_row(B, x.v + 1, y);
  fp_cmad_row(A, x.v, y);
  ptx_addc(B[11], B[11], 0u);
}
B200_DEV fp_pair fp_mul_dual(const fp &a, const fp &b, const fp &c, const fp &d) {
  uint32_t ev0[12], od0[12], ev1[12], od1[12];
#pragma unroll
  for (int j = 0; j < 12; j += 2) {
    ptx_mul_lo(ev0[j], a.v[j], b.v[0]);
    ptx_mul_hi(ev0[j + 1], a.v[j], b.v[0]);
    ptx_mul_lo(od0[j], a.v[j + 1], b.v[0]);
    ptx_mul_hi(od0[j + 1], a.v[j + 1], b.v[0]);
    ptx_mul_lo(ev1[j], c.v[j], d.v[0]);
    ptx_mul_hi(ev1[j + 1], c.v[j], d.v[0]);
    ptx_mul_lo(od1[j], c.v[j + 1], d.v[0]);
    ptx_mul_hi(od1[j + 1], c.v[j + 1], d.v[0]);
  }
  fp_redc_step(ev0, od0);
  fp_redc_step(ev1, od1);
#pragma unroll
  for (int i = 1; i < 12; i += 2) {
    fp_mul_row_acc(od0, ev0, a, b.v[i]);
    fp_mul_row_acc(od1, ev1, c, d.v[i]);
    fp_redc_step(od0, ev0);
    fp_redc_step(od1, ev1);
    if (i + 1 < 12) {
      fp_mul_row_acc(ev0, od0, a, b.v[i + 1]);
      fp_mul_row_acc(ev1, od1, c, d.v[i + 1]);
      fp_redc_step(ev0, od0);
      fp_redc_step(ev1, od1);
    }
  }
  return fp_pair{fp_mul_tail(ev0, od0), fp_mul_tail(ev1, od1)};
}
// ---------------------------------------------------------------------------------------------------
// Lazy reduction support (used by the Fp2 multiplication): an unreduced 768-bit product, its Montgomery
// reduction, and plain (non-modular) 384/768-bit add/sub.  Karatsuba Fp2 mul = 3 wide products (3 x 144
// IMAD) + 2 reductions (2 x 156) = 744 IMAD instead of 3 x 300 = 900.  8p < 2^384 leaves room for the
// unreduced operand sums; every result that leaves these helpers is canonical again.
struct fpw {
  uint32_t v[24];
};

// chain over acc[0..12) += x[0],x[2],..,x[10] * s, then the carry into acc[12] (when it exists)
template <bool TOP>
B200_DEV void fp_cmad_row_c(uint32_t *acc, const uint32_t *x, uint32_t s) {
  fp_cmad_row(acc, x, s);
  if (TOP) ptx_addc(acc[12], acc[12], 0u);
}

// t = a * b as a 24-word integer (no reduction).  Even/odd aligned accumulators as in fp_mul.
B200_DEV fpw fp_mul_wide(const fp &a, const fp &b) {
  uint32_t E[24], O[24];  // O[k] sits at word k+1
#pragma unroll
  for (int k = 0; k < 24; k++) E[k] = O[k] = 0;
#pragma unroll
  for (int i = 0; i < 12; i += 2) {
    // even row i: a_even*b_i at even offset i (E), a_odd*b_i at odd offset i+1 (O index i)
    fp_cmad_row_c<true>(E + i, a.v, b.v[i]);
    fp_cmad_row_c<true>(O + i, a.v + 1, b.v[i]);
    // odd row i+1: a_even*b_{i+1} at odd offset i+1 (O index i), a_odd*b_{i+1} at even offset i+2 (E)
    fp_cmad_row_c<true>(O + i, a.v, b.v[i + 1]);
    if (i + 14 < 24)
      fp_cmad_row_c<true>(E + i + 2, a.v + 1, b.v[i + 1]);
    else
      fp_cmad_row_c<false>(E + i + 2, a.v + 1, b.v[i + 1]);  // top chain: no carry out (a*b < 2^768)
  }
  fpw t;
  t.v[0] = E[0];
  ptx_add_cc(t.v[1], E[1], O[0]);
#pragma unroll
  for (int k = 2; k < 23; k++) ptx_addc_cc(t.v[k], E[k], O[k - 1]);
  ptx_addc(t.v[23], E[23], O[22]);
  return t;
}
B200_DEV fpw fpw_add(const fpw &a, const fpw &b) {  // caller guarantees no overflow of 768 bits
  fpw r;
  ptx_add_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
  for (int k = 1; k < 23; k++) ptx_addc_cc(r.v[k], a.v[k], b.v[k]);
  ptx_addc(r.v[23], a.v[23], b.v[23]);
  return r;
}
// a - b ; when the difference is negative, p * 2^384 is added (result in [0, p*2^384))
B200_DEV fpw fpw_sub_mod(const fpw &a, const fpw &b) {
  fpw r;
  uint32_t borrow;
  ptx_sub_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
  for (int k = 1; k < 24; k++) ptx_subc_cc(r.v[k], a.v[k], b.v[k]);
  ptx_subc(borrow, 0u, 0u);
  ptx_add_cc(r.v[12], r.v[12], fp_modw(0) & borrow);
#pragma unroll
  for (int k = 1; k < 11; k++) ptx_addc_cc(r.v[12 + k], r.v[12 + k], fp_modw(k) & borrow);
  ptx_addc(r.v[23], r.v[23], fp_modw(11) & borrow);
  return r;
}
B200_DEV fpw fpw_sub(const fpw &a, const fpw &b) {  // caller guarantees a >= b
  fpw r;
  ptx_sub_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
  for (int k = 1; k < 23; k++) ptx_subc_cc(r.v[k], a.v[k], b.v[k]);
  ptx_subc(r.v[23], a.v[23], b.v[23]);
  return r;
}
// plain 384-bit a + b (no reduction; caller guarantees < 2^384, e.g. both < p)
B200_DEV fp fp_add_nr(const fp &a, const fp &b) {
  fp r;
  ptx_add_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
  for (int k = 1; k < 11; k++) ptx_addc_cc(r.v[k], a.v[k], b.v[k]);
  ptx_addc(r.v[11], a.v[11], b.v[11]);
  return r;
}
// Montgomery reduction of t < p * 2^384:  (t + m p) / 2^384 mod p, canonical.
//   REDC(t) = REDC(t_low) + t_high, and REDC(t_low) is the interleaved multiplier run with b = 1:
//   row 0 contributes a*1, rows 1..11 only reduce and shift (the shift is an add-with-carry chain on the
//   otherwise idle ALU pipe; the IMAD pipe sees 12 x 12 + 12 = 156 multiplies).
B200_DEV void fp_redc_wide_step(uint32_t *A, uint32_t *B);
B200_DEV fp fp_redc_wide_tail(const uint32_t *ev, const uint32_t *od, const fpw &t);
B200_DEV fp fp_redc_wide(const fpw &t) {
  uint32_t ev[12], od[12];
#pragma unroll
  for (int j = 0; j < 12; j += 2) {
    ev[j] = t.v[j];
    ev[j + 1] = 0;
    od[j] = t.v[j + 1];
    od[j + 1] = 0;
  }
  fp_redc_step(ev, od);
#pragma unroll
  for (int i = 1; i < 12; i += 2) {
    fp_redc_wide_step(od, ev);  // roles as in fp_mul: A = od (aligned at word 0 after the shift), B = ev (shifted in by two words)
    if (i + 1 < 12) fp_redc_wide_step(ev, od);
  }
  // q_low = (od >> 32) + ev ; result = q_low + t_high, then one conditional subtraction
  return fp_redc_wide_tail(ev, od, t);
}
// acc[0..2*len) += x[0], x[2], .., x[2*(len-1)] * s as one carry chain; the carry goes into acc[2*len] when TOP
template <int LEN, bool TOP>
B200_DEV void fp_cmad_n(uint32_t *acc, const uint32_t *x, uint32_t s) {
  if (LEN <= 0) return;
  ptx_mad_lo_cc(acc[0], x[0], s, acc[0]);
  ptx_madc_hi_cc(acc[1], x[0], s, acc[1]);
#pragma unroll
  for (int j = 1; j < LEN; j++) {
    ptx_madc_lo_cc(acc[2 * j], x[2 * j], s, acc[2 * j]);
    ptx_madc_hi_cc(acc[2 * j + 1], x[2 * j], s, acc[2 * j + 1]);
  }
  if (TOP) ptx_addc(acc[2 * LEN], acc[2 * LEN], 0u);
}
template <int I>
B200_DEV void fp_sqr_row(uint32_t *E, uint32_t *O, const fp &a) {
  // row I: a_j * a_I for j > I.  j of the parity of I land on even word offsets (E), the others on odd ones (O).
  constexpr int LE = (11 - I) / 2;      // j = I+2, I+4, ...  at word offsets 2I+2, 2I+4, ...
  constexpr int LO = (12 - I) / 2;      // j = I+1, I+3, ...  at word offsets 2I+1, ...  -> O index 2I, 2I+2, ...
  fp_cmad_n<LE, (2 * I + 2 + 2 * LE < 24)>(E + 2 * I + 2, a.v + I + 2, a.v[I]);
  fp_cmad_n<LO, (2 * I + 2 * LO < 23)>(O + 2 * I, a.v + I + 1, a.v[I]);
}
// a^2 as a 24-word integer: 66 off-diagonal products, doubled, plus the 12 squares (78 IMAD.WIDE instead of 144)
B200_DEV fpw fp_sqr_wide(const fp &a) {
  uint32_t E[24], O[24];
#pragma unroll
  for (int k = 0; k < 24; k++) E[k] = O[k] = 0;
  fp_sqr_row<0>(E, O, a);
  fp_sqr_row<1>(E, O, a);
  fp_sqr_row<2>(E, O, a);
  fp_sqr_row<3>(E, O, a);
  fp_sqr_row<4>(E, O, a);
  fp_sqr_row<5>(E, O, a);
  fp_sqr_row<6>(E, O, a);
  fp_sqr_row<7>(E, O, a);
  fp_sqr_row<8>(E, O, a);
  fp_sqr_row<9>(E, O, a);
  fp_sqr_row<10>(E, O, a);
  // T = E + (O << 32)
  uint32_t T[24];
  T[0] = E[0];
  ptx_add_cc(T[1], E[1], O[0]);
#pragma unroll
  for (int k = 2; k < 23; k++) ptx_addc_cc(T[k], E[k], O[k - 1]);
  ptx_addc(T[23], E[23], O[22]);
  // T = 2 T
#pragma unroll
  for (int k = 23; k > 0; k--) T[k] = __funnelshift_l(T[k - 1], T[k], 1);
  T[0] <<= 1;
  // T += sum a_i^2 * 2^(64 i): one chain over all 24 words
  fpw t;
  ptx_mad_lo_cc(t.v[0], a.v[0], a.v[0], T[0]);
  ptx_madc_hi_cc(t.v[1], a.v[0], a.v[0], T[1]);
#pragma unroll
  for (int i = 1; i < 12; i++) {
    ptx_madc_lo_cc(t.v[2 * i], a.v[i], a.v[i], T[2 * i]);
    if (i < 11)
      ptx_madc_hi_cc(t.v[2 * i + 1], a.v[i], a.v[i], T[2 * i + 1]);
    else
      ptx_madc_hi(t.v[23], a.v[11], a.v[11], T[23]);
  }
  return t;
}
// Row-alternated versions of the lazy-reduction building blocks (B200_FP2_LAZY3, the MSM unit's Fp2 multiply): three unreduced
// products with their rows alternated, and two Montgomery reductions with their steps alternated.
struct fpw3 {
  fpw w0, w1, w2;
};
B200_DEV fpw fpw_join(const uint32_t *E, const uint32_t *O) {
  fpw t;
  t.v[0] = E[0];
  ptx_add_cc(t.v[1], E[1], O[0]);
#pragma unroll
  for (int k = 2; k < 23; k++) ptx_addc_cc(t.v[k], E[k], O[k - 1]);
  ptx_addc(t.v[23], E[23], O[22]);
  return t;
}
template <int I>
B200_DEV void fp_mul_wide_rows(uint32_t *E, uint32_t *O, const fp &a, const fp &b) {  // rows I and I+1 of fp_mul_wide
  fp_cmad_row_c<true>(E + I, a.v, b.v[I]);
  fp_cmad_row_c<true>(O + I, a.v + 1, b.v[I]);
  fp_cmad_row_c<true>(O + I, a.v, b.v[I + 1]);
  if (I + 14 < 24)
    fp_cmad_row_c<true>(E + I + 2, a.v + 1, b.v[I + 1]);
  else
    fp_cmad_row_c<false>(E + I + 2, a.v + 1, b.v[I + 1]);
}
template <int I>
B200_DEV void fp_mul_wide_rows3(uint32_t *E0, uint32_t *O0, uint32_t *E1, uint32_t *O1, uint32_t *E2, uint32_t *O2, const fp &a,
                                const fp &b, const fp &c, const fp &d, const fp &e, const fp &f) {
  fp_mul_wide_rows<I>(E0, O0, a, b);
  fp_mul_wide_rows<I>(E1, O1, c, d);
  fp_mul_wide_rows<I>(E2, O2, e, f);
}
B200_DEV fpw3 fp_mul_wide_triple(const fp &a, const fp &b, const fp &c, const fp &d, const fp &e, const fp &f) {
  uint32_t E0[24], O0[24], E1[24], O1[24], E2[24], O2[24];
#pragma unroll
  for (int k = 0; k < 24; k++) E0[k] = O0[k] = E1[k] = O1[k] = E2[k] = O2[k] = 0;
  fp_mul_wide_rows3<0>(E0, O0, E1, O1, E2, O2, a, b, c, d, e, f);
  fp_mul_wide_rows3<2>(E0, O0, E1, O1, E2, O2, a, b, c, d, e, f);
  fp_mul_wide_rows3<4>(E0, O0, E1, O1, E2, O2, a, b, c, d, e, f);
  fp_mul_wide_rows3<6>(E0, O0, E1, O1, E2, O2, a, b, c, d, e, f);
  fp_mul_wide_rows3<8>(E0, O0, E1, O1, E2, O2, a, b, c, d, e, f);
  fp_mul_wide_rows3<10>(E0, O0, E1, O1, E2, O2, a, b, c, d, e, f);
  return fpw3{fpw_join(E0, O0), fpw_join(E1, O1), fpw_join(E2, O2)};
}
// one shift-and-reduce step of fp_redc_wide on (A aligned at word 0 after the shift, B shifted in by two words).
// The two-word shift of B rides on the m * p_odd products (three-operand multiply-add: B[k] = p*m + B[k+2], carry-in = the
// carry of A[0] += B[1]), exactly like the a*b_i rows of fp_mul — no add-with-carry chain on the ALU pipe per row (round 1
// spent 12 ALU instructions per row there: 132 of the ~320 non-multiply instructions of a reduction).
B200_DEV void fp_redc_wide_step(uint32_t *A, uint32_t *B) {
  uint32_t m;
  ptx_add_cc(A[0], A[0], B[1]);
  ptx_mul_lo(m, A[0], FP_INV32);  // mul.lo leaves CC.CF alone
#pragma unroll
  for (int j = 0; j < 10; j += 2) {
    ptx_madc_lo_cc(B[j], fp_modw(1 + j), m, B[j + 2]);
    ptx_madc_hi_cc(B[j + 1], fp_modw(1 + j), m, B[j + 3]);
  }
  ptx_madc_lo_cc(B[10], fp_modw(11), m, 0u);
  ptx_madc_hi(B[11], fp_modw(11), m, 0u);  // no carry out (value bound)
  fp_cmad_row_mod<0>(A, m);
  ptx_addc(B[11], B[11], 0u);
}
B200_DEV fp fp_redc_wide_tail(const uint32_t *ev, const uint32_t *od, const fpw &t) {
  fp r;
  ptx_add_cc(r.v[0], ev[0], od[1]);
#pragma unroll
  for (int k = 1; k < 11; k++) ptx_addc_cc(r.v[k], ev[k], od[k + 1]);
  ptx_addc(r.v[11], ev[11], 0u);
  ptx_add_cc(r.v[0], r.v[0], t.v[12]);
#pragma unroll
  for (int k = 1; k < 11; k++) ptx_addc_cc(r.v[k], r.v[k], t.v[12 + k]);
  ptx_addc(r.v[11], r.v[11], t.v[23]);
  uint32_t d[12], borrow;
  ptx_sub_cc(d[0], r.v[0], fp_modw(0));
#pragma unroll
  for (int k = 1; k < 12; k++) ptx_subc_cc(d[k], r.v[k], fp_modw(k));
  ptx_subc(borrow, 0u, 0u);
#pragma unroll
  for (int k = 0; k < 12; k++) r.v[k] = borrow ? r.v[k] : d[k];
  return r;
}
B200_DEV fp_pair fp_redc_wide_dual(const fpw &t0, const fpw &t1) {
  uint32_t ev0[12], od0[12], ev1[12], od1[12];
#pragma unroll
  for (int j = 0; j < 12; j += 2) {
    ev0[j] = t0.v[j];
    ev0[j + 1] = 0;
    od0[j] = t0.v[j + 1];
    od0[j + 1] = 0;
    ev1[j] = t1.v[j];
    ev1[j + 1] = 0;
    od1[j] = t1.v[j + 1];
    od1[j + 1] = 0;
  }
  fp_redc_step(ev0, od0);
  fp_redc_step(ev1, od1);
#pragma unroll
  for (int i = 1; i < 12; i += 2) {
    fp_redc_wide_step(od0, ev0);
    fp_redc_wide_step(od1, ev1);
    if (i + 1 < 12) {
      fp_redc_wide_step(ev0, od0);
      fp_redc_wide_step(ev1, od1);
    }
  }
  return fp_pair{fp_redc_wide_tail(ev0, od0, t0), fp_redc_wide_tail(ev1, od1, t1)};
}

// a^2 * R^-1 mod p, canonical (same value as src/fp.rs:613-660): 78 + 156 = 234 IMAD instead of 305
B200_DEV fp fp_sqr_fast(const fp &a) { return fp_redc_wide(fp_sqr_wide(a)); }
static __device__ __noinline__ fp fp_sqr_c(fp a) { return fp_sqr_fast(a); }
static __device__ __noinline__ fpw fp_mul_wide_c(fp a, fp b) { return fp_mul_wide(a, b); }
static __device__ __noinline__ fp fp_redc_wide_c(fpw t) { return fp_redc_wide(t); }

// src/fp.rs:382-393
B200_DEV fp fp_add(const fp &a, const fp &b) {
  fp r;
  ptx_add_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
  for (int k = 1; k < 11; k++) ptx_addc_cc(r.v[k], a.v[k], b.v[k]);
  ptx_addc(r.v[11], a.v[11], b.v[11]);
  uint32_t t[12], borrow;
  ptx_sub_cc(t[0], r.v[0], fp_modw(0));
#pragma unroll
  for (int k = 1; k < 12; k++) ptx_subc_cc(t[k], r.v[k], fp_modw(k));
  ptx_subc(borrow, 0u, 0u);
#pragma unroll
  for (int k = 0; k < 12; k++) r.v[k] = borrow ? r.v[k] : t[k];
  return r;
}
// a - b mod p  (same value as src/fp.rs:421-423, which computes neg(b) + a)
B200_DEV fp fp_sub(const fp &a, const fp &b) {
  fp r;
  uint32_t borrow;
  ptx_sub_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
  for (int k = 1; k < 12; k++) ptx_subc_cc(r.v[k], a.v[k], b.v[k]);
  ptx_subc(borrow, 0u, 0u);  // 0xffffffff when a < b
  ptx_add_cc(r.v[0], r.v[0], fp_modw(0) & borrow);
#pragma unroll
  for (int k = 1; k < 11; k++) ptx_addc_cc(r.v[k], r.v[k], fp_modw(k) & borrow);
  ptx_addc(r.v[11], r.v[11], fp_modw(11) & borrow);
  return r;
}
B200_DEV bool fp_is_zero(const fp &a) {
  uint32_t o = 0;
#pragma unroll
  for (int k = 0; k < 12; k++) o |= a.v[k];
  return o == 0;
}
B200_DEV bool fp_eq(const fp &a, const fp &b) {
  uint32_t o = 0;
#pragma unroll
  for (int k = 0; k < 12; k++) o |= a.v[k] ^ b.v[k];
  return o == 0;
}
// src/fp.rs:396-418
B200_DEV fp fp_neg(const fp &a) {
  fp r;
  uint32_t mask = fp_is_zero(a) ? 0u : 0xffffffffu;
  ptx_sub_cc(r.v[0], fp_modw(0), a.v[0]);
#pragma unroll
  for (int k = 1; k < 11; k++) ptx_subc_cc(r.v[k], fp_modw(k), a.v[k]);
  ptx_subc(r.v[11], fp_modw(11), a.v[11]);
#pragma unroll
  for (int k = 0; k < 12; k++) r.v[k] &= mask;
  return r;
}
B200_DEV fp fp_dbl(const fp &a) { return fp_add(a, a); }
B200_DEV fp fp_select(const fp &a, const fp &b, bool choose_b) {
  fp r;
#pragma unroll
  for (int k = 0; k < 12; k++) r.v[k] = choose_b ? b.v[k] : a.v[k];
  return r;
}

// a^(p-2)  (src/fp.rs:346-358 via pow_vartime :309-321: 384 squarings, multiply on set bits, MSB first)
B200_DEV fp fp_inv(const fp &a) {
  fp res = fp_one();
#pragma unroll 1
  for (int w = 11; w >= 0; w--) {
    uint32_t e = FP_MOD[w] - (w == 0 ? 2u : 0u);  // p - 2 : low word 0xffffaaab - 2, no borrow
#pragma unroll 1
    for (int i = 31; i >= 0; i--) {
      res = fp_sqr(res);
      if ((e >> i) & 1) res = fp_mul(res, a);
    }
  }
  return res;
}

// ---- global memory <-> registers.  Memory layout = the reference's Fp([u64;6]) little-endian limbs,
// i.e. 12 consecutive little-endian 32-bit words, 16-byte aligned groups -> three 128-bit accesses.
B200_DEV fp fp_load(const void *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = q[0], b = q[1], c = q[2];
  fp r = {{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w}};
  return r;
}
B200_DEV void fp_store(void *p, const fp &r) {
  uint4 *q = reinterpret_cast<uint4 *>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
  q[2] = make_uint4(r.v[8], r.v[9], r.v[10], r.v[11]);
}
B200_DEV fp fp_load_ro(const void *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  fp r = {{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w}};
  return r;
}

}  // namespace b200
