// Lane-cooperative Fp12 arithmetic for the pairing kernels (pairing_coop.cu).
//
// One pairing is worked on by a GROUP of 6 consecutive lanes of a warp (5 groups per warp; lanes 30/31 exit at once).  An Fp12 element is held as a degree-5 polynomial in w over Fp2,
//     X = x_0 + x_1 w + ... + x_5 w^5,   w^6 = xi = 1 + u,
// lane k of the group keeping the ONE coefficient x_k (24 registers).  In terms of the reference's structs
// (src/fp12.rs:11-14, src/fp6.rs:9-13; Fp12 = Fp6[w]/(w^2 - v), Fp6 = Fp2[v]/(v^3 - xi)):  x_{2j} = c0.c_j,
// x_{2j+1} = c1.c_j.  Nothing of an Fp12 ever lives in local memory: operands are exchanged through a small
// shared-memory BOARD per group (24 slots of one Fp2), every lane computes one output coefficient as a dot product of
// Fp2 values with LAZY reduction (all partial products accumulated as unreduced 768-bit integers, two Montgomery
// reductions per coefficient), and the lanes of a group stay in step with __syncwarp(mask of the group).
//
// Every function returns canonical field elements, so results are bit-identical to src/fp12.rs Mul :197 / square
// :174 / mul_by_014 :116 / frobenius_map :145 / conjugate :136 / invert :187 and src/pairings.rs cyclotomic_square :66.
// Cost per lane in 32x32+64 multiply-adds (one Fp multiplication = 300):
//   mul 18x144 + 2x156 = 2904 (reference model 54 FpM / 6 lanes = 2700)     sqr 9x144 + 312 = 1608 (model 1800)
//   mul by a sparse line (ell) 300 + 1608 (model 43 FpM / 6 = 2150)         cyclotomic square 900 (model 900)
#pragma once
#include "constants.cuh"
#include "fp_inv.cuh"
#include "tower.cuh"

namespace b200 {

constexpr int CO_LANES = 6;                     // lanes per pairing
constexpr int CO_GROUPS = 5;                    // pairings per warp
constexpr int CO_SLOT = 28;                     // 32-bit words between board slots: an Fp2 (c0 then c1, 24 words) + 4 words of padding, so that the
                                                // six slots a group reads at once start in six different 16-byte bank groups (7 i mod 8)
constexpr int CO_NSLOT = 24;                    // slots per group
constexpr int CO_BOARD = CO_SLOT * CO_NSLOT;    // words per group
constexpr int CO_WARP_SMEM = CO_GROUPS * CO_BOARD * 4;  // bytes of shared memory per warp

struct cgrp {
  uint32_t *bd;    // this group's board (shared memory)
  int k;           // coefficient index of this lane, 0..5
  unsigned mask;   // the six lanes of this group: groups synchronise independently (lanes 30/31 of a warp exit at once)
};

B200_DEV void co_sync(const cgrp &g) { __syncwarp(g.mask); }
B200_DEV uint32_t *co_slot(const cgrp &g, int s) { return g.bd + CO_SLOT * s; }
B200_DEV fp co_ld(const uint32_t *p) { return fp_load(p); }
B200_DEV fp2 co_ld2(const uint32_t *p) { return fp2{fp_load(p), fp_load(p + 12)}; }
B200_DEV void co_put(const cgrp &g, int s, const fp2 &a) {
  fp_store(co_slot(g, s), a.c0);
  fp_store(co_slot(g, s) + 12, a.c1);
}
B200_DEV void co_put_half(const cgrp &g, int s, int half, const fp &a) { fp_store(co_slot(g, s) + 12 * half, a); }
B200_DEV int co_mod6(int i) { return i < 0 ? i + 6 : (i >= 6 ? i - 6 : i); }
B200_DEV int co_mod3(int i) { return i < 0 ? i + 3 : (i >= 3 ? i - 3 : i); }

// t - p 2^384 when t >= p 2^384 (t < 2 p 2^384): brings a lazily accumulated sum back under the bound of fp_redc_wide
B200_DEV fpw fpw_csub_pR(const fpw &t) {
  uint32_t d[12], borrow;
  ptx_sub_cc(d[0], t.v[12], fp_modw(0));
#pragma unroll
  for (int i = 1; i < 12; i++) ptx_subc_cc(d[i], t.v[12 + i], fp_modw(i));
  ptx_subc(borrow, 0u, 0u);
  fpw r = t;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[12 + i] = borrow ? t.v[12 + i] : d[i];
  return r;
}
B200_DEV fpw fpw_zero() {
  fpw r;
#pragma unroll
  for (int i = 0; i < 24; i++) r.v[i] = 0;
  return r;
}

// sum_{t < N} A_t * B_t over Fp2, A_t / B_t = the 24-word operands pa(t) / pb(t) point at (canonical, < p).
// Karatsuba per product, everything accumulated unreduced:
//   w0 = sum a0 b0, w1 = sum a1 b1, w2 = sum (a0+a1)(b0+b1)          (each < 4 N p^2 < 2^768 for N <= 6)
//   re = REDC(w0 - w1 [+ p 2^384 if negative]),  im = REDC(w2 - w0 - w1 [- p 2^384 if >= p 2^384])
// |w0 - w1| < N p^2 <= 6 p^2 < p 2^384 (2^384 = 9.84 p);  w2 - w0 - w1 = sum (a0 b1 + a1 b0) < 2 N p^2 <= 12 p^2 < 2 p 2^384.
// 3 N wide products (144 multiply-adds each) + 2 reductions (156 each).
template <int N, class PA, class PB>
B200_DEV fp2 co_dot(PA pa, PB pb) {
  static_assert(N >= 1 && N <= 6, "lazy-reduction bound");
  fpw w0 = fpw_zero(), w1 = fpw_zero();
#pragma unroll 1
  for (int t = 0; t < N; t++) w0 = fpw_add(w0, fp_mul_wide_c(co_ld(pa(t)), co_ld(pb(t))));
#pragma unroll 1
  for (int t = 0; t < N; t++) w1 = fpw_add(w1, fp_mul_wide_c(co_ld(pa(t) + 12), co_ld(pb(t) + 12)));
  fp re = fp_redc_wide_c(fpw_sub_mod(w0, w1));
  w0 = fpw_add(w0, w1);
  w1 = fpw_zero();
#pragma unroll 1
  for (int t = 0; t < N; t++) {
    const uint32_t *a = pa(t), *b = pb(t);
    w1 = fpw_add(w1, fp_mul_wide_c(fp_add_nr(co_ld(a), co_ld(a + 12)), fp_add_nr(co_ld(b), co_ld(b + 12))));
  }
  w1 = fpw_sub(w1, w0);
  if (N > 4) w1 = fpw_csub_pR(w1);
  return fp2{re, fp_redc_wide_c(w1)};
}

// The operations below are CALLED (not inlined): one copy of each in the kernel.  Fully inlined, the pairing kernel was
// 29 000 instructions (468 KB) and ncu showed the instruction fetch as its top stall (no_instruction 5.1 cycles per issue).
// board slot map (slots 18.. are owned by the Miller loop: R = (X, Y, Z), Q = (x, y), line coefficients)
constexpr int CS_F = 0, CS_G = 6, CS_GX = 12;

// Z = X * Y   (src/fp12.rs:197-214).  z_k = sum_{i+j=k} x_i y_j + xi sum_{i+j=k+6} x_i y_j: the lanes publish x, y and xi*y,
// lane k takes y_j for j <= k and xi*y_j for j > k.
B200_NOINL fp2 co_mul(cgrp g, fp2 x, fp2 y) {
  co_put(g, CS_F + g.k, x);
  co_put(g, CS_G + g.k, y);
  co_put(g, CS_GX + g.k, fp2_mul_by_nonresidue(y));
  co_sync(g);
  const int k = g.k;
  const uint32_t *bd = g.bd;
  fp2 z = co_dot<6>([=](int t) { return bd + CO_SLOT * (CS_F + co_mod6(k - t)); },
                    [=](int t) { return bd + CO_SLOT * ((t > k ? CS_GX : CS_G) + t); });
  co_sync(g);
  return z;
}

// X^2 by the complex method over Fp6 (src/fp12.rs:174-185): with A = (x0, x2, x4), B = (x1, x3, x5) in Fp6,
//   M1 = A B,  M2 = (A + B)(A + v B),  c0 = M2 - M1 - v M1,  c1 = 2 M1.
// Odd lanes compute the three coefficients of M1, even lanes those of M2 (three Fp2 products each, lazily reduced).
B200_NOINL fp2 co_sqr(cgrp g, fp2 x) {
  constexpr int S_RX1 = 6, S_L2 = 9, S_R2 = 12, S_RX2 = 15;
  const int k = g.k, c = k >> 1, par = k & 1;
  const uint32_t *bd = g.bd;
  co_put(g, CS_F + k, x);
  co_sync(g);
  if (par) {
    co_put(g, S_RX1 + c, fp2_mul_by_nonresidue(x));  // xi * B_c
  } else {
    fp2 b = co_ld2(bd + CO_SLOT * (CS_F + k + 1));                                    // B_c
    fp2 vb = c == 0 ? fp2_mul_by_nonresidue(co_ld2(bd + CO_SLOT * (CS_F + 5)))        // (v B)_0 = xi B_2
                    : co_ld2(bd + CO_SLOT * (CS_F + k - 1));                          // (v B)_c = B_{c-1}
    fp2 q = fp2_add(x, vb);
    co_put(g, S_L2 + c, fp2_add(x, b));
    co_put(g, S_R2 + c, q);
    co_put(g, S_RX2 + c, fp2_mul_by_nonresidue(q));
  }
  co_sync(g);
  fp2 m = co_dot<3>(
      [=](int t) {
        int li = co_mod3(c - t);
        return bd + CO_SLOT * (par ? CS_F + 2 * li : S_L2 + li);
      },
      [=](int t) {
        bool wrap = t > c;
        return bd + CO_SLOT * (par ? (wrap ? S_RX1 + t : CS_F + 2 * t + 1) : (wrap ? S_RX2 + t : S_R2 + t));
      });
  co_sync(g);
  co_put(g, CS_F + k, m);  // odd slots: M1_c
  co_sync(g);
  fp2 r;
  if (par) {
    r = fp2_dbl(m);
  } else {
    fp2 m1 = co_ld2(bd + CO_SLOT * (CS_F + k + 1));
    fp2 vm1 = c == 0 ? fp2_mul_by_nonresidue(co_ld2(bd + CO_SLOT * (CS_F + 5))) : co_ld2(bd + CO_SLOT * (CS_F + k - 1));
    r = fp2_sub(fp2_sub(m, m1), vm1);
  }
  co_sync(g);
  return r;
}

// X * (c0 + c1 w^2 + c4 w^3)  ==  Fp12::mul_by_014(c0, c1, c4)  (src/fp12.rs:116-128; the line of the Miller loop):
//   z_k = x_k c0 + [xi] x_{k-2} c1 + [xi] x_{k-3} c4      (xi on the terms that wrap around w^6)
// pc0 / pc1 / pc4 point at the three Fp2 (shared or global memory); x is published together with xi * x.
B200_NOINL fp2 co_mul_sparse(cgrp g, fp2 x, const uint32_t *pc0, const uint32_t *pc1, const uint32_t *pc4) {
  constexpr int S_FX = 6;
  const int k = g.k;
  const uint32_t *bd = g.bd;
  co_put(g, CS_F + k, x);
  co_put(g, S_FX + k, fp2_mul_by_nonresidue(x));
  co_sync(g);
  fp2 z = co_dot<3>(
      [=](int t) {
        int s = t == 0 ? CS_F + k : (t == 1 ? (k < 2 ? S_FX : CS_F) + co_mod6(k - 2) : (k < 3 ? S_FX : CS_F) + co_mod6(k - 3));
        return bd + CO_SLOT * s;
      },
      [=](int t) { return t == 0 ? pc0 : (t == 1 ? pc1 : pc4); });
  co_sync(g);
  return z;
}

// p - a as a plain 384-bit subtraction: the negative of a canonical a, except that 0 maps to p (not canonical).  Safe as the
// SECOND operand of fp_add with a canonical first operand: a' + p - p = a'.  Never add two of these to each other.
B200_DEV fp fp_neg_lazy(const fp &a) {
  fp r;
  ptx_sub_cc(r.v[0], fp_modw(0), a.v[0]);
#pragma unroll
  for (int i = 1; i < 11; i++) ptx_subc_cc(r.v[i], fp_modw(i), a.v[i]);
  ptx_subc(r.v[11], fp_modw(11), a.v[11]);
  return r;
}

// cyclotomic_square (src/pairings.rs:66-113).  The three fp4_square calls work on the coefficient pairs
// (x0, x3), (x1, x4), (x2, x5); lanes q and q + 3 share pair q: each squares its own coefficient (two Fp products) and
// takes one half of (a + b)^2 — three Fp multiplications per lane, the 18 of the reference spread evenly.
// Every lane publishes its square t AND p - t, so that the two kinds of output coefficient
//   c0 = a^2 + xi b^2 = (t0.re + t1.re - t1.im,  t0.im + t1.re + t1.im)          (even lanes)
//   c1 = (a+b)^2 - a^2 - b^2 = (H.re - t0.re - t1.re,  H.im - t0.im - t1.im)       (odd lanes)
// are the SAME instruction sequence — three-term sums over lane-specific pointers — instead of two divergent branches that a
// warp executes one after the other (round 2, first version: 10 modular additions per lane and squaring, now 4).
B200_DEV fp2 co_cyclotomic_sqr_body(const cgrp &g, const fp2 &x) {
  constexpr int S_T = 6, S_H = 12, S_NT = 15;
  const int k = g.k;
  const uint32_t *bd = g.bd;
  co_put(g, CS_F + k, x);
  co_sync(g);
  const bool alane = k < 3;
  fp2 y = co_ld2(bd + CO_SLOT * (CS_F + co_mod6(k + 3)));
  fp2 s = fp2_add(x, y);
  // sums that only feed a multiplication stay unreduced (< 2p; the product of the operand bounds stays far below 2^384 / p = 9.8,
  // so the one conditional subtraction of fp_mul still lands in [0, p))
  fp2 own{fp_mul_c(fp_add_nr(x.c0, x.c1), fp_sub(x.c0, x.c1)), fp_mul_c(fp_add_nr(x.c0, x.c0), x.c1)};
  fp hu = fp_add_nr(s.c0, fp_select(s.c0, s.c1, alane));   // a-lane: s0 + s1, b-lane: 2 s0
  fp hv = fp_select(s.c1, fp_sub(s.c0, s.c1), alane);      // a-lane: s0 - s1, b-lane: s1
  fp h = fp_mul_c(hu, hv);  // a-lane: re (a+b)^2, b-lane: im (a+b)^2
  co_put(g, S_T + k, own);
  co_put(g, S_NT + k, fp2{fp_neg_lazy(own.c0), fp_neg_lazy(own.c1)});
  co_put_half(g, S_H + (alane ? k : k - 3), alane ? 0 : 1, h);
  co_sync(g);
  // which pair feeds lane k, and how (z-names of the reference: x0=z0, x3=z1, x1=z2, x4=z3, x2=z4, x5=z5):
  //   x0 <- 3 c0(0) - 2 x0, x3 <- 3 c1(0) + 2 x3, x2 <- 3 c0(1) - 2 x2, x5 <- 3 c1(1) + 2 x5, x1 <- 3 xi c1(2) + 2 x1, x4 <- 3 c0(2) - 2 x4
  const int q = k == 0 || k == 3 ? 0 : (k == 2 || k == 5 ? 1 : 2);
  const bool use_c0 = k == 0 || k == 2 || k == 4;
  const uint32_t *T0 = bd + CO_SLOT * (S_T + q), *T1 = bd + CO_SLOT * (S_T + q + 3);
  const uint32_t *N0 = bd + CO_SLOT * (S_NT + q), *N1 = bd + CO_SLOT * (S_NT + q + 3), *H = bd + CO_SLOT * (S_H + q);
  // canonical operands first, at most one `p - t` per addition (fp_neg_lazy)
  const uint32_t *r0 = use_c0 ? T0 : H, *r1 = use_c0 ? T1 : N0, *r2 = use_c0 ? N1 + 12 : N1;
  const uint32_t *i0 = use_c0 ? T0 + 12 : H + 12, *i1 = use_c0 ? T1 : N0 + 12, *i2 = use_c0 ? T1 + 12 : N1 + 12;
  fp2 c{fp_add(fp_add(co_ld(r0), co_ld(r1)), co_ld(r2)), fp_add(fp_add(co_ld(i0), co_ld(i1)), co_ld(i2))};
  if (k == 1) c = fp2_mul_by_nonresidue(c);
  // d = c - x (even lanes) / c + x (odd lanes), r = 2 d + c
  fp2 xt{fp_select(x.c0, fp_neg_lazy(x.c0), use_c0), fp_select(x.c1, fp_neg_lazy(x.c1), use_c0)};
  fp2 d = fp2_add(c, xt);
  fp2 r = fp2_add(fp2_dbl(d), c);
  co_sync(g);
  return r;
}
// n >= 1 squarings in a row: the loop lives INSIDE the called function (one copy of the body, the accumulator stays in its
// registers) — called once per squaring, the argument / result moves and the caller's spills around the call were 280 of the
// 2 600 instructions of a squaring
B200_NOINL fp2 co_cyclotomic_sqr_n(cgrp g, fp2 x, int n) {
#pragma unroll 1
  for (; n > 0; n--) x = co_cyclotomic_sqr_body(g, x);
  return x;
}
B200_DEV fp2 co_cyclotomic_sqr(const cgrp &g, const fp2 &x) { return co_cyclotomic_sqr_n(g, x, 1); }

// frobenius_map^n, n = 1..3 (src/fp12.rs:145-171 applied n times): coefficient-wise conj^n(x_k) * xi^(k (p^n - 1)/6)
B200_NOINL fp2 co_frobenius(cgrp g, fp2 x, int n) {
  fp2 v = (n & 1) ? fp2_conj(x) : x;
  co_put(g, CS_F + g.k, v);   // own slot, read back by this lane only
  const uint32_t *pa = g.bd + CO_SLOT * (CS_F + g.k);
  const uint32_t *pb = &K_FROBW[6 * (n - 1) + g.k][0];
  co_sync(g);
  fp2 z = co_dot<1>([=](int) { return pa; }, [=](int) { return pb; });
  co_sync(g);
  return z;
}
B200_DEV fp2 co_conj(const cgrp &g, const fp2 &x) { return (g.k & 1) ? fp2_neg(x) : x; }
B200_DEV fp2 co_one(const cgrp &g) { return g.k == 0 ? fp2_one() : fp2_zero(); }

// X^-1 (src/fp12.rs:187-195: t = (c0^2 - v c1^2)^-1, (c0 t, -c1 t); Fp6 inverse src/fp6.rs:294-312; Fp2 inverse
// src/fp2.rs:300-320).  The Fp inversion at the bottom is the binary-GCD fp_inv_fast (same value as the reference's
// Fermat exponentiation).  Even lanes carry c0 = A, odd lanes c1 = B; both triples compute t redundantly.
B200_NOINL fp2 co_inv(cgrp g, fp2 x, const uint32_t *pow2) {
  // slots: F 0-5, xi*F 6-11 (first dot only; then c' in 6-8, T^-1 in 9-11), T 12-14 (later xi*T^-1), xi*T 15-17
  constexpr int S_FX = 6, S_T = 12, S_TX = 15, S_C = 6, S_TI = 9;
  const int k = g.k, c = k >> 1, par = k & 1;
  const uint32_t *bd = g.bd;
  // squares of the two Fp6 halves
  co_put(g, CS_F + k, x);
  co_put(g, S_FX + k, fp2_mul_by_nonresidue(x));
  co_sync(g);
  fp2 sq = co_dot<3>([=](int t) { return bd + CO_SLOT * (CS_F + 2 * co_mod3(c - t) + par); },
                     [=](int t) { return bd + CO_SLOT * ((t > c ? S_FX : CS_F) + 2 * t + par); });
  co_sync(g);
  co_put(g, CS_F + k, sq);
  co_sync(g);
  // T = A^2 - v B^2
  fp2 vb = c == 0 ? fp2_mul_by_nonresidue(co_ld2(bd + CO_SLOT * (CS_F + 5))) : co_ld2(bd + CO_SLOT * (CS_F + 2 * c - 1));
  fp2 T = fp2_sub(co_ld2(bd + CO_SLOT * (CS_F + 2 * c)), vb);
  if (!par) {
    co_put(g, S_T + c, T);
    co_put(g, S_TX + c, fp2_mul_by_nonresidue(T));
  }
  co_sync(g);
  // Fp6 inverse of T = (T0, T1, T2): c'_0 = T0 T0 - (xi T1) T2, c'_1 = (xi T2) T2 - T0 T1, c'_2 = T1 T1 - T0 T2
  const int s1a = c == 0 ? S_T + 0 : (c == 1 ? S_TX + 2 : S_T + 1), s1b = c == 0 ? S_T + 0 : (c == 1 ? S_T + 2 : S_T + 1);
  const int s2a = c == 0 ? S_TX + 1 : S_T + 0, s2b = c == 1 ? S_T + 1 : S_T + 2;
  fp2 p1a = co_ld2(bd + CO_SLOT * s1a), p1b = co_ld2(bd + CO_SLOT * s1b);
  fp2 p2a = co_ld2(bd + CO_SLOT * s2a), p2b = co_ld2(bd + CO_SLOT * s2b);
  fp2 cp = fp2_sub(M2(p1a, p1b), M2(p2a, p2b));
  if (!par) co_put(g, S_C + c, cp);
  co_sync(g);
  // N = T0 c'_0 + xi (T1 c'_2 + T2 c'_1), computed by every lane
  fp2 N = fp2_add(M2(co_ld2(bd + CO_SLOT * S_T), co_ld2(bd + CO_SLOT * S_C)),
                  fp2_mul_by_nonresidue(fp2_add(M2(co_ld2(bd + CO_SLOT * (S_T + 1)), co_ld2(bd + CO_SLOT * (S_C + 2))),
                                                M2(co_ld2(bd + CO_SLOT * (S_T + 2)), co_ld2(bd + CO_SLOT * (S_C + 1))))));
  fp ni = fp_inv_fast(fp_add(fp_sqr_c(N.c0), fp_sqr_c(N.c1)), pow2);
  fp2 Ninv{fp_mul_c(N.c0, ni), fp_mul_c(N.c1, fp_neg(ni))};
  fp2 ti = M2(cp, Ninv);  // (T^-1)_c
  co_sync(g);
  co_put(g, CS_F + k, x);
  if (!par) {
    co_put(g, S_TI + c, ti);
    co_put(g, S_T + c, fp2_mul_by_nonresidue(ti));
  }
  co_sync(g);
  fp2 r = co_dot<3>([=](int t) { return bd + CO_SLOT * (CS_F + 2 * co_mod3(c - t) + par); },
                    [=](int t) { return bd + CO_SLOT * ((t > c ? S_T : S_TI) + t); });
  co_sync(g);
  return par ? fp2_neg(r) : r;
}

// ---- memory <-> lanes.  An Fp12 in memory is 12 Fp in struct order c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2 (576 B):
// coefficient k sits at byte offset 96 * (3 * (k & 1) + (k >> 1)).
B200_DEV int co_mem_offset(int k) { return 96 * (3 * (k & 1) + (k >> 1)); }
B200_DEV fp2 co_load12(const cgrp &g, const char *p) { return fp2_load(p + co_mem_offset(g.k)); }
B200_DEV void co_store12(const cgrp &g, char *p, const fp2 &x) { fp2_store(p + co_mem_offset(g.k), x); }

}  // namespace b200
