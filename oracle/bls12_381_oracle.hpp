// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (C++17, unsigned __int128) of the zkcrypto/bls12_381 v0.8.0 algorithms on the
// hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may link or call this.  Each function cites the reference file:line it
// follows (paths relative to /root/reference).  Parity is PINNED: tests/test_oracle_golden.py checks
// this code against the reference's own known-answer tests and the four src/tests/*.dat golden
// vector files (committed as tests/golden/*.npz by tests/golden/make_golden.py).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <vector>

namespace bls_oracle {

typedef unsigned __int128 u128;
typedef uint64_t u64;

// ---------------------------------------------------------------- src/util.rs:3-19
static inline u64 adc(u64 a, u64 b, u64 &carry) {
  u128 r = (u128)a + b + carry;
  carry = (u64)(r >> 64);
  return (u64)r;
}
static inline u64 sbb(u64 a, u64 b, u64 &borrow) {
  // borrow is 0 or 0xffff..ffff, as in the reference (src/util.rs:10-14)
  u128 r = (u128)a - ((u128)b + (borrow >> 63));
  borrow = (u64)(r >> 64);
  return (u64)r;
}
static inline u64 mac(u64 a, u64 b, u64 c, u64 &carry) {
  u128 r = (u128)a + (u128)b * c + carry;
  carry = (u64)(r >> 64);
  return (u64)r;
}

// ================================================================= Fp  (src/fp.rs)
struct Fp {
  u64 l[6];
};

// src/fp.rs:70-77
static const u64 FP_MODULUS[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                  0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
// src/fp.rs:80
static const u64 FP_INV = 0x89f3fffcfffcfffdULL;
// src/fp.rs:83-90
static const Fp FP_R = {{0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                         0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL}};
// src/fp.rs:93-100
static const Fp FP_R2 = {{0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL,
                          0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL}};
// src/fp.rs:103-110
static const Fp FP_R3 = {{0xed48ac6bd94ca1e0ULL, 0x315f831e03a7adf8ULL, 0x9a53352a615e29ddULL,
                          0x34c04e5e921e1761ULL, 0x2512d43565724728ULL, 0x0aa6346091755d4dULL}};

static inline Fp fp_zero() { return Fp{{0, 0, 0, 0, 0, 0}}; }
static inline Fp fp_one() { return FP_R; }
static inline bool fp_eq(const Fp &a, const Fp &b) { return std::memcmp(a.l, b.l, 48) == 0; }
static inline bool fp_is_zero(const Fp &a) {
  return (a.l[0] | a.l[1] | a.l[2] | a.l[3] | a.l[4] | a.l[5]) == 0;
}

// src/fp.rs:361-379
static inline Fp fp_subtract_p(const Fp &a) {
  u64 borrow = 0;
  Fp r;
  for (int i = 0; i < 6; i++) r.l[i] = sbb(a.l[i], FP_MODULUS[i], borrow);
  // borrow == all-ones  => a < p => keep a
  Fp out;
  for (int i = 0; i < 6; i++) out.l[i] = (a.l[i] & borrow) | (r.l[i] & ~borrow);
  return out;
}
// src/fp.rs:382-393
static inline Fp fp_add(const Fp &a, const Fp &b) {
  u64 carry = 0;
  Fp r;
  for (int i = 0; i < 6; i++) r.l[i] = adc(a.l[i], b.l[i], carry);
  return fp_subtract_p(r);
}
// src/fp.rs:396-418
static inline Fp fp_neg(const Fp &a) {
  u64 borrow = 0;
  Fp r;
  for (int i = 0; i < 6; i++) r.l[i] = sbb(FP_MODULUS[i], a.l[i], borrow);
  u64 mask = fp_is_zero(a) ? 0 : ~(u64)0;
  for (int i = 0; i < 6; i++) r.l[i] &= mask;
  return r;
}
// src/fp.rs:421-423  (sub = neg(rhs) + self)
static inline Fp fp_sub(const Fp &a, const Fp &b) { return fp_add(fp_neg(b), a); }

// src/fp.rs:487-562  (HAC 14.32)
static inline Fp fp_montgomery_reduce(const u64 tin[12]) {
  u64 t[13];
  for (int i = 0; i < 12; i++) t[i] = tin[i];
  t[12] = 0;
  u64 carry2 = 0;  // the running top carry (r7..r12 chain in the reference)
  for (int i = 0; i < 6; i++) {
    u64 k = t[i] * FP_INV;
    u64 carry = 0;
    (void)mac(t[i], k, FP_MODULUS[0], carry);
    for (int j = 1; j < 6; j++) t[i + j] = mac(t[i + j], k, FP_MODULUS[j], carry);
    t[i + 6] = adc(t[i + 6], carry2, carry);
    carry2 = carry;
  }
  Fp r = {{t[6], t[7], t[8], t[9], t[10], t[11]}};
  return fp_subtract_p(r);
}
// src/fp.rs:565-609
static inline Fp fp_mul(const Fp &a, const Fp &b) {
  u64 t[12] = {0};
  for (int i = 0; i < 6; i++) {
    u64 carry = 0;
    for (int j = 0; j < 6; j++) t[i + j] = mac(t[i + j], a.l[i], b.l[j], carry);
    t[i + 6] = carry;
  }
  return fp_montgomery_reduce(t);
}
// src/fp.rs:613-660: dedicated squaring — the 15 off-diagonal products a_i a_j (i < j), doubled by a one-bit shift of
// the 12-limb intermediate, plus the six squares a_i^2 on the even limbs; same Montgomery reduction (21 mac instead of 36)
static inline Fp fp_square(const Fp &a) {
  u64 t[12] = {0};
  for (int i = 0; i < 5; i++) {
    u64 carry = 0;
    for (int j = i + 1; j < 6; j++) t[i + j] = mac(t[i + j], a.l[i], a.l[j], carry);
    t[i + 6] = carry;
  }
  for (int k = 11; k > 0; k--) t[k] = (t[k] << 1) | (t[k - 1] >> 63);
  t[0] = 0;
  u64 carry = 0;
  for (int i = 0; i < 6; i++) {
    t[2 * i] = mac(t[2 * i], a.l[i], a.l[i], carry);
    t[2 * i + 1] = adc(t[2 * i + 1], 0, carry);
  }
  return fp_montgomery_reduce(t);
}

// src/fp.rs:430-484  (Longa eprint 2022/367 Alg. 2: interleaved sum of products)
template <int T>
static inline Fp fp_sum_of_products(const Fp a[T], const Fp b[T]) {
  u64 u[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < 6; j++) {
    u64 t[7];
    for (int i = 0; i < 7; i++) t[i] = u[i];
    for (int i = 0; i < T; i++) {
      u64 carry = 0;
      for (int k = 0; k < 6; k++) t[k] = mac(t[k], a[i].l[j], b[i].l[k], carry);
      t[6] = adc(t[6], 0, carry);
    }
    u64 k = t[0] * FP_INV;
    u64 carry = 0;
    (void)mac(t[0], k, FP_MODULUS[0], carry);
    for (int i = 1; i < 6; i++) u[i - 1] = mac(t[i], k, FP_MODULUS[i], carry);
    u64 c2 = carry;
    carry = 0;
    u[5] = adc(t[6], c2, carry);
    u[6] = carry;  // always fits (reference drops it: src/fp.rs:476)
  }
  Fp r = {{u[0], u[1], u[2], u[3], u[4], u[5]}};
  return fp_subtract_p(r);
}

// src/fp.rs:309-321
static inline Fp fp_pow_vartime(const Fp &a, const u64 by[6]) {
  Fp res = fp_one();
  for (int e = 5; e >= 0; e--)
    for (int i = 63; i >= 0; i--) {
      res = fp_square(res);
      if ((by[e] >> i) & 1) res = fp_mul(res, a);
    }
  return res;
}
// src/fp.rs:346-358  (returns zero for zero input == unwrap_or(zero) used by callers)
static inline Fp fp_invert(const Fp &a, bool *ok = nullptr) {
  static const u64 e[6] = {0xb9feffffffffaaa9ULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                           0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
  if (ok) *ok = !fp_is_zero(a);
  return fp_pow_vartime(a, e);
}
// src/fp.rs:324-343
static inline bool fp_sqrt(const Fp &a, Fp &out) {
  static const u64 e[6] = {0xee7fbfffffffeaabULL, 0x07aaffffac54ffffULL, 0xd9cc34a83dac3d89ULL,
                           0xd91dd2e13ce144afULL, 0x92c6e9ed90d2eb35ULL, 0x0680447a8e5ff9a6ULL};
  out = fp_pow_vartime(a, e);
  return fp_eq(fp_square(out), a);
}
// src/fp.rs:179-208 ; big-endian 48 bytes ; returns false when non-canonical
static inline bool fp_from_bytes(const uint8_t b[48], Fp &out) {
  Fp t;
  for (int i = 0; i < 6; i++) {
    u64 v = 0;
    for (int k = 0; k < 8; k++) v = (v << 8) | b[(5 - i) * 8 + k];
    t.l[i] = v;
  }
  u64 borrow = 0;
  for (int i = 0; i < 6; i++) (void)sbb(t.l[i], FP_MODULUS[i], borrow);
  out = fp_mul(t, FP_R2);
  return (borrow & 1) == 1;
}
// src/fp.rs:211-227
static inline void fp_to_bytes(const Fp &a, uint8_t b[48]) {
  u64 t[12] = {a.l[0], a.l[1], a.l[2], a.l[3], a.l[4], a.l[5], 0, 0, 0, 0, 0, 0};
  Fp c = fp_montgomery_reduce(t);
  for (int i = 0; i < 6; i++)
    for (int k = 0; k < 8; k++) b[(5 - i) * 8 + k] = (uint8_t)(c.l[i] >> (56 - 8 * k));
}
// src/fp.rs:273-299
static inline bool fp_lexicographically_largest(const Fp &a) {
  static const u64 h[6] = {0xdcff7fffffffd556ULL, 0x0f55ffff58a9ffffULL, 0xb39869507b587b12ULL,
                           0xb23ba5c279c2895fULL, 0x258dd3db21a5d66bULL, 0x0d0088f51cbff34dULL};
  u64 t[12] = {a.l[0], a.l[1], a.l[2], a.l[3], a.l[4], a.l[5], 0, 0, 0, 0, 0, 0};
  Fp c = fp_montgomery_reduce(t);
  u64 borrow = 0;
  for (int i = 0; i < 6; i++) (void)sbb(c.l[i], h[i], borrow);
  return !((borrow & 1) == 1);
}

// ================================================================= Fp2 (src/fp2.rs)
struct Fp2 {
  Fp c0, c1;
};
static inline Fp2 fp2_zero() { return Fp2{fp_zero(), fp_zero()}; }
static inline Fp2 fp2_one() { return Fp2{fp_one(), fp_zero()}; }
static inline bool fp2_eq(const Fp2 &a, const Fp2 &b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }
static inline bool fp2_is_zero(const Fp2 &a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
// src/fp2.rs:224-243
static inline Fp2 fp2_add(const Fp2 &a, const Fp2 &b) { return Fp2{fp_add(a.c0, b.c0), fp_add(a.c1, b.c1)}; }
static inline Fp2 fp2_sub(const Fp2 &a, const Fp2 &b) { return Fp2{fp_sub(a.c0, b.c0), fp_sub(a.c1, b.c1)}; }
static inline Fp2 fp2_neg(const Fp2 &a) { return Fp2{fp_neg(a.c0), fp_neg(a.c1)}; }
// src/fp2.rs:141-153
static inline Fp2 fp2_conjugate(const Fp2 &a) { return Fp2{a.c0, fp_neg(a.c1)}; }
static inline Fp2 fp2_frobenius_map(const Fp2 &a) { return fp2_conjugate(a); }
// src/fp2.rs:156-166
static inline Fp2 fp2_mul_by_nonresidue(const Fp2 &a) { return Fp2{fp_sub(a.c0, a.c1), fp_add(a.c0, a.c1)}; }
// src/fp2.rs:182-203
static inline Fp2 fp2_square(const Fp2 &x) {
  Fp a = fp_add(x.c0, x.c1), b = fp_sub(x.c0, x.c1), c = fp_add(x.c0, x.c0);
  return Fp2{fp_mul(a, b), fp_mul(c, x.c1)};
}
// src/fp2.rs:205-222
static inline Fp2 fp2_mul(const Fp2 &x, const Fp2 &y) {
  Fp a0[2] = {x.c0, fp_neg(x.c1)}, b0[2] = {y.c0, y.c1};
  Fp a1[2] = {x.c0, x.c1}, b1[2] = {y.c1, y.c0};
  return Fp2{fp_sum_of_products<2>(a0, b0), fp_sum_of_products<2>(a1, b1)};
}
// src/fp2.rs:300-320
static inline Fp2 fp2_invert(const Fp2 &a, bool *ok = nullptr) {
  Fp t = fp_invert(fp_add(fp_square(a.c0), fp_square(a.c1)), ok);
  return Fp2{fp_mul(a.c0, t), fp_mul(a.c1, fp_neg(t))};
}
// src/fp2.rs:171-180
static inline bool fp2_lexicographically_largest(const Fp2 &a) {
  return fp_lexicographically_largest(a.c1) || (fp_is_zero(a.c1) && fp_lexicographically_largest(a.c0));
}
// src/fp2.rs:322-336
static inline Fp2 fp2_pow_vartime(const Fp2 &a, const u64 by[6]) {
  Fp2 res = fp2_one();
  for (int e = 5; e >= 0; e--)
    for (int i = 63; i >= 0; i--) {
      res = fp2_square(res);
      if ((by[e] >> i) & 1) res = fp2_mul(res, a);
    }
  return res;
}
// src/fp2.rs:245-298  (Alg. 9 of eprint 2012/685)
static inline bool fp2_sqrt(const Fp2 &a, Fp2 &out) {
  if (fp2_is_zero(a)) {
    out = fp2_zero();
    return true;
  }
  static const u64 e1[6] = {0xee7fbfffffffeaaaULL, 0x07aaffffac54ffffULL, 0xd9cc34a83dac3d89ULL,
                            0xd91dd2e13ce144afULL, 0x92c6e9ed90d2eb35ULL, 0x0680447a8e5ff9a6ULL};
  static const u64 e2[6] = {0xdcff7fffffffd555ULL, 0x0f55ffff58a9ffffULL, 0xb39869507b587b12ULL,
                            0xb23ba5c279c2895fULL, 0x258dd3db21a5d66bULL, 0x0d0088f51cbff34dULL};
  Fp2 a1 = fp2_pow_vartime(a, e1);
  Fp2 alpha = fp2_mul(fp2_square(a1), a);
  Fp2 x0 = fp2_mul(a1, a);
  Fp2 cand;
  if (fp2_eq(alpha, fp2_neg(fp2_one()))) {
    cand = Fp2{fp_neg(x0.c1), x0.c0};
  } else {
    cand = fp2_mul(fp2_pow_vartime(fp2_add(alpha, fp2_one()), e2), x0);
  }
  out = cand;
  return fp2_eq(fp2_square(cand), a);
}

// ================================================================= Fp6 (src/fp6.rs)
struct Fp6 {
  Fp2 c0, c1, c2;
};
static inline Fp6 fp6_zero() { return Fp6{fp2_zero(), fp2_zero(), fp2_zero()}; }
static inline Fp6 fp6_one() { return Fp6{fp2_one(), fp2_zero(), fp2_zero()}; }
static inline Fp6 fp6_add(const Fp6 &a, const Fp6 &b) { return Fp6{fp2_add(a.c0, b.c0), fp2_add(a.c1, b.c1), fp2_add(a.c2, b.c2)}; }
static inline Fp6 fp6_sub(const Fp6 &a, const Fp6 &b) { return Fp6{fp2_sub(a.c0, b.c0), fp2_sub(a.c1, b.c1), fp2_sub(a.c2, b.c2)}; }
static inline Fp6 fp6_neg(const Fp6 &a) { return Fp6{fp2_neg(a.c0), fp2_neg(a.c1), fp2_neg(a.c2)}; }
// src/fp6.rs:113-119
static inline Fp6 fp6_mul_by_1(const Fp6 &s, const Fp2 &c1) {
  return Fp6{fp2_mul_by_nonresidue(fp2_mul(s.c2, c1)), fp2_mul(s.c0, c1), fp2_mul(s.c1, c1)};
}
// src/fp6.rs:121-136
static inline Fp6 fp6_mul_by_01(const Fp6 &s, const Fp2 &c0, const Fp2 &c1) {
  Fp2 a_a = fp2_mul(s.c0, c0), b_b = fp2_mul(s.c1, c1);
  Fp2 t1 = fp2_add(fp2_mul_by_nonresidue(fp2_mul(s.c2, c1)), a_a);
  Fp2 t2 = fp2_sub(fp2_sub(fp2_mul(fp2_add(c0, c1), fp2_add(s.c0, s.c1)), a_a), b_b);
  Fp2 t3 = fp2_add(fp2_mul(s.c2, c0), b_b);
  return Fp6{t1, t2, t3};
}
// src/fp6.rs:139-150
static inline Fp6 fp6_mul_by_nonresidue(const Fp6 &a) { return Fp6{fp2_mul_by_nonresidue(a.c2), a.c0, a.c1}; }
// src/fp6.rs:154-188
static const Fp2 FP6_FROB_C1 = {{{0, 0, 0, 0, 0, 0}},
                                {{0xcd03c9e48671f071ULL, 0x5dab22461fcda5d2ULL, 0x587042afd3851b95ULL,
                                  0x8eb60ebe01bacb9eULL, 0x03f97d6e83d050d2ULL, 0x18f0206554638741ULL}}};
static const Fp2 FP6_FROB_C2 = {{{0x890dc9e4867545c3ULL, 0x2af322533285a5d5ULL, 0x50880866309b7e2cULL,
                                  0xa20d1b8c7e881024ULL, 0x14e4f04fe2db9068ULL, 0x14e56d3f1564853aULL}},
                                {{0, 0, 0, 0, 0, 0}}};
static inline Fp6 fp6_frobenius_map(const Fp6 &a) {
  Fp2 c0 = fp2_frobenius_map(a.c0), c1 = fp2_frobenius_map(a.c1), c2 = fp2_frobenius_map(a.c2);
  return Fp6{c0, fp2_mul(c1, FP6_FROB_C1), fp2_mul(c2, FP6_FROB_C2)};
}
// src/fp6.rs:200-274
static inline Fp6 fp6_mul(const Fp6 &a, const Fp6 &b) {
  Fp b10_p_b11 = fp_add(b.c1.c0, b.c1.c1), b10_m_b11 = fp_sub(b.c1.c0, b.c1.c1);
  Fp b20_p_b21 = fp_add(b.c2.c0, b.c2.c1), b20_m_b21 = fp_sub(b.c2.c0, b.c2.c1);
  Fp an[6] = {a.c0.c0, fp_neg(a.c0.c1), a.c1.c0, fp_neg(a.c1.c1), a.c2.c0, fp_neg(a.c2.c1)};
  Fp ap[6] = {a.c0.c0, a.c0.c1, a.c1.c0, a.c1.c1, a.c2.c0, a.c2.c1};
  Fp b00[6] = {b.c0.c0, b.c0.c1, b20_m_b21, b20_p_b21, b10_m_b11, b10_p_b11};
  Fp b01[6] = {b.c0.c1, b.c0.c0, b20_p_b21, b20_m_b21, b10_p_b11, b10_m_b11};
  Fp b10[6] = {b.c1.c0, b.c1.c1, b.c0.c0, b.c0.c1, b20_m_b21, b20_p_b21};
  Fp b11[6] = {b.c1.c1, b.c1.c0, b.c0.c1, b.c0.c0, b20_p_b21, b20_m_b21};
  Fp b20[6] = {b.c2.c0, b.c2.c1, b.c1.c0, b.c1.c1, b.c0.c0, b.c0.c1};
  Fp b21[6] = {b.c2.c1, b.c2.c0, b.c1.c1, b.c1.c0, b.c0.c1, b.c0.c0};
  Fp6 r;
  r.c0.c0 = fp_sum_of_products<6>(an, b00);
  r.c0.c1 = fp_sum_of_products<6>(ap, b01);
  r.c1.c0 = fp_sum_of_products<6>(an, b10);
  r.c1.c1 = fp_sum_of_products<6>(ap, b11);
  r.c2.c0 = fp_sum_of_products<6>(an, b20);
  r.c2.c1 = fp_sum_of_products<6>(ap, b21);
  return r;
}
// src/fp6.rs:277-291
static inline Fp6 fp6_square(const Fp6 &a) {
  Fp2 s0 = fp2_square(a.c0);
  Fp2 ab = fp2_mul(a.c0, a.c1);
  Fp2 s1 = fp2_add(ab, ab);
  Fp2 s2 = fp2_square(fp2_add(fp2_sub(a.c0, a.c1), a.c2));
  Fp2 bc = fp2_mul(a.c1, a.c2);
  Fp2 s3 = fp2_add(bc, bc);
  Fp2 s4 = fp2_square(a.c2);
  return Fp6{fp2_add(fp2_mul_by_nonresidue(s3), s0), fp2_add(fp2_mul_by_nonresidue(s4), s1),
             fp2_sub(fp2_sub(fp2_add(fp2_add(s1, s2), s3), s0), s4)};
}
// src/fp6.rs:294-312
static inline Fp6 fp6_invert(const Fp6 &a, bool *ok = nullptr) {
  Fp2 c0 = fp2_sub(fp2_square(a.c0), fp2_mul_by_nonresidue(fp2_mul(a.c1, a.c2)));
  Fp2 c1 = fp2_sub(fp2_mul_by_nonresidue(fp2_square(a.c2)), fp2_mul(a.c0, a.c1));
  Fp2 c2 = fp2_sub(fp2_square(a.c1), fp2_mul(a.c0, a.c2));
  Fp2 tmp = fp2_mul_by_nonresidue(fp2_add(fp2_mul(a.c1, c2), fp2_mul(a.c2, c1)));
  tmp = fp2_add(tmp, fp2_mul(a.c0, c0));
  Fp2 t = fp2_invert(tmp, ok);
  return Fp6{fp2_mul(t, c0), fp2_mul(t, c1), fp2_mul(t, c2)};
}

// ================================================================= Fp12 (src/fp12.rs)
struct Fp12 {
  Fp6 c0, c1;
};
static inline Fp12 fp12_one() { return Fp12{fp6_one(), fp6_zero()}; }
static inline bool fp12_eq(const Fp12 &a, const Fp12 &b) { return std::memcmp(&a, &b, sizeof(Fp12)) == 0; }
// src/fp12.rs:116-128
static inline Fp12 fp12_mul_by_014(const Fp12 &s, const Fp2 &c0, const Fp2 &c1, const Fp2 &c4) {
  Fp6 aa = fp6_mul_by_01(s.c0, c0, c1);
  Fp6 bb = fp6_mul_by_1(s.c1, c4);
  Fp2 o = fp2_add(c1, c4);
  Fp6 r1 = fp6_add(s.c1, s.c0);
  r1 = fp6_mul_by_01(r1, c0, o);
  r1 = fp6_sub(fp6_sub(r1, aa), bb);
  Fp6 r0 = fp6_add(fp6_mul_by_nonresidue(bb), aa);
  return Fp12{r0, r1};
}
// src/fp12.rs:136-141
static inline Fp12 fp12_conjugate(const Fp12 &a) { return Fp12{a.c0, fp6_neg(a.c1)}; }
// src/fp12.rs:145-171
static const Fp2 FP12_FROB_C1 = {{{0x07089552b319d465ULL, 0xc6695f92b50a8313ULL, 0x97e83cccd117228fULL,
                                   0xa35baecab2dc29eeULL, 0x1ce393ea5daace4dULL, 0x08f2220fb0fb66ebULL}},
                                 {{0xb2f66aad4ce5d646ULL, 0x5842a06bfc497cecULL, 0xcf4895d42599d394ULL,
                                   0xc11b9cba40a8e8d0ULL, 0x2e3813cbe5a0de89ULL, 0x110eefda88847fafULL}}};
// src/fp12.rs:197-214
static inline Fp12 fp12_mul(const Fp12 &a, const Fp12 &b) {
  Fp6 aa = fp6_mul(a.c0, b.c0);
  Fp6 bb = fp6_mul(a.c1, b.c1);
  Fp6 o = fp6_add(b.c0, b.c1);
  Fp6 c1 = fp6_add(a.c1, a.c0);
  c1 = fp6_mul(c1, o);
  c1 = fp6_sub(c1, aa);
  c1 = fp6_sub(c1, bb);
  Fp6 c0 = fp6_add(fp6_mul_by_nonresidue(bb), aa);
  return Fp12{c0, c1};
}
static inline Fp12 fp12_frobenius_map(const Fp12 &a) {
  Fp6 c0 = fp6_frobenius_map(a.c0), c1 = fp6_frobenius_map(a.c1);
  Fp6 k = Fp6{FP12_FROB_C1, fp2_zero(), fp2_zero()};  // Fp6::from(Fp2)
  return Fp12{c0, fp6_mul(c1, k)};
}
// src/fp12.rs:174-185
static inline Fp12 fp12_square(const Fp12 &a) {
  Fp6 ab = fp6_mul(a.c0, a.c1);
  Fp6 c0c1 = fp6_add(a.c0, a.c1);
  Fp6 c0 = fp6_add(fp6_mul_by_nonresidue(a.c1), a.c0);
  c0 = fp6_mul(c0, c0c1);
  c0 = fp6_sub(c0, ab);
  Fp6 c1 = fp6_add(ab, ab);
  c0 = fp6_sub(c0, fp6_mul_by_nonresidue(ab));
  return Fp12{c0, c1};
}
// src/fp12.rs:187-195
static inline Fp12 fp12_invert(const Fp12 &a, bool *ok = nullptr) {
  Fp6 t = fp6_invert(fp6_sub(fp6_square(a.c0), fp6_mul_by_nonresidue(fp6_square(a.c1))), ok);
  return Fp12{fp6_mul(a.c0, t), fp6_mul(a.c1, fp6_neg(t))};
}

// ================================================================= Scalar (src/scalar.rs) — marshalling only
static const u64 FR_MODULUS[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                                  0x73eda753299d7d48ULL};  // src/scalar.rs:76-81
static const u64 FR_INV = 0xfffffffeffffffffULL;           // src/scalar.rs:156
static const u64 FR_R2[4] = {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL,
                             0x0748d9d99f59ff11ULL};  // src/scalar.rs:168-173
static const u64 FR_R3[4] = {0xc62c1807439b73afULL, 0x1b3e0d188cf06990ULL, 0x73d13c71c7b5f418ULL,
                             0x6e2a5bb9c8db33e9ULL};  // src/scalar.rs:176-181
struct Scalar {
  u64 l[4];
};  // Montgomery form, R = 2^256
// src/scalar.rs:506-550
static inline Scalar fr_montgomery_reduce(const u64 tin[8]) {
  u64 t[9];
  for (int i = 0; i < 8; i++) t[i] = tin[i];
  t[8] = 0;
  u64 carry2 = 0;
  for (int i = 0; i < 4; i++) {
    u64 k = t[i] * FR_INV;
    u64 carry = 0;
    (void)mac(t[i], k, FR_MODULUS[0], carry);
    for (int j = 1; j < 4; j++) t[i + j] = mac(t[i + j], k, FR_MODULUS[j], carry);
    t[i + 4] = adc(t[i + 4], carry2, carry);
    carry2 = carry;
  }
  // result = (t4..t7) - q if >= q  (src/scalar.rs:549: (&Scalar([r4..r7])).sub(&MODULUS))
  u64 borrow = 0, d[4];
  for (int i = 0; i < 4; i++) d[i] = sbb(t[4 + i], FR_MODULUS[i], borrow);
  // the reference's sub adds the modulus back when the subtraction underflowed
  Scalar r;
  u64 carry = 0;
  for (int i = 0; i < 4; i++) r.l[i] = adc(d[i], FR_MODULUS[i] & borrow, carry);
  return r;
}
static inline Scalar fr_add(const Scalar &a, const Scalar &b) {  // src/scalar.rs:578-589
  u64 carry = 0, d[4];
  for (int i = 0; i < 4; i++) d[i] = adc(a.l[i], b.l[i], carry);
  u64 borrow = 0, e[4];
  for (int i = 0; i < 4; i++) e[i] = sbb(d[i], FR_MODULUS[i], borrow);
  Scalar r;
  carry = 0;
  for (int i = 0; i < 4; i++) r.l[i] = adc(e[i], FR_MODULUS[i] & borrow, carry);
  return r;
}
static inline Scalar fr_mul(const Scalar &a, const u64 b[4]) {  // src/scalar.rs:554-575
  u64 t[8] = {0};
  for (int i = 0; i < 4; i++) {
    u64 carry = 0;
    for (int j = 0; j < 4; j++) t[i + j] = mac(t[i + j], a.l[i], b[j], carry);
    t[i + 4] = carry;
  }
  return fr_montgomery_reduce(t);
}
// src/scalar.rs:284-296 : Montgomery -> canonical 32-byte little-endian
static inline void fr_to_bytes(const Scalar &s, uint8_t out[32]) {
  u64 t[8] = {s.l[0], s.l[1], s.l[2], s.l[3], 0, 0, 0, 0};
  Scalar c = fr_montgomery_reduce(t);
  for (int i = 0; i < 4; i++)
    for (int k = 0; k < 8; k++) out[i * 8 + k] = (uint8_t)(c.l[i] >> (8 * k));
}
// src/scalar.rs:300-331 : 64 little-endian bytes -> Scalar (Montgomery), d0*R2 + d1*R3
static inline Scalar fr_from_bytes_wide(const uint8_t b[64]) {
  u64 w[8];
  for (int i = 0; i < 8; i++) {
    u64 v = 0;
    for (int k = 7; k >= 0; k--) v = (v << 8) | b[i * 8 + k];
    w[i] = v;
  }
  Scalar d0 = {{w[0], w[1], w[2], w[3]}}, d1 = {{w[4], w[5], w[6], w[7]}};
  return fr_add(fr_mul(d0, FR_R2), fr_mul(d1, FR_R3));
}

// ----------------------------------------------------------------- Scalar field arithmetic
// src/scalar.rs:159-164 (R = Scalar::one()), :183-188 (TWO_INV), :191 (S), :200-205 (ROOT_OF_UNITY),
// :208-213 (ROOT_OF_UNITY_INV), :100-105 (GENERATOR = 7)
static const Scalar FR_ONE = {{0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL}};
static const Scalar FR_TWO_INV = {{0x00000000ffffffffULL, 0xac425bfd0001a401ULL, 0xccc627f7f65e27faULL, 0x0c1258acd66282b7ULL}};
static const Scalar FR_ROOT_OF_UNITY = {{0xb9b58d8c5f0e466aULL, 0x5b1b4c801819d7ecULL, 0x0af53ae352a31e64ULL, 0x5bf3adda19e9b27bULL}};
static const Scalar FR_ROOT_OF_UNITY_INV = {{0x4256481adcf3219aULL, 0x45f37b7f96b6cad3ULL, 0xf9c3f1d75f7a3b27ULL, 0x2d2fc049658afd43ULL}};
static const Scalar FR_GENERATOR = {{0x0000000efffffff1ULL, 0x17e363d300189c0fULL, 0xff9c57876f8457b0ULL, 0x351332208fc5a8c4ULL}};
static const int FR_S = 32;
static inline Scalar fr_zero() { return Scalar{{0, 0, 0, 0}}; }
static inline bool fr_is_zero(const Scalar &a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
static inline bool fr_eq(const Scalar &a, const Scalar &b) {
  return a.l[0] == b.l[0] && a.l[1] == b.l[1] && a.l[2] == b.l[2] && a.l[3] == b.l[3];
}
static inline Scalar fr_sub(const Scalar &a, const Scalar &b) {  // src/scalar.rs:582-597
  u64 borrow = 0, d[4];
  for (int i = 0; i < 4; i++) d[i] = sbb(a.l[i], b.l[i], borrow);
  Scalar r;
  u64 carry = 0;
  for (int i = 0; i < 4; i++) r.l[i] = adc(d[i], FR_MODULUS[i] & borrow, carry);
  return r;
}
static inline Scalar fr_neg(const Scalar &a) {  // src/scalar.rs:613-627
  u64 borrow = 0, d[4];
  for (int i = 0; i < 4; i++) d[i] = sbb(FR_MODULUS[i], a.l[i], borrow);
  u64 mask = fr_is_zero(a) ? 0 : ~(u64)0;
  return Scalar{{d[0] & mask, d[1] & mask, d[2] & mask, d[3] & mask}};
}
static inline Scalar fr_mul(const Scalar &a, const Scalar &b) { return fr_mul(a, b.l); }
// src/scalar.rs:341-370 (dedicated squaring; the same canonical value as mul(a, a), which is what is computed here)
static inline Scalar fr_square(const Scalar &a) { return fr_mul(a, a.l); }
static inline Scalar fr_double(const Scalar &a) { return fr_add(a, a); }  // src/scalar.rs:249-252
// src/scalar.rs:392-404
static inline Scalar fr_pow_vartime(const Scalar &a, const u64 by[4]) {
  Scalar res = FR_ONE;
  for (int w = 3; w >= 0; w--)
    for (int i = 63; i >= 0; i--) {
      res = fr_square(res);
      if ((by[w] >> i) & 1) res = fr_mul(res, a);
    }
  return res;
}
// src/scalar.rs:408-503 computes a^(q-2) with a fixed addition chain; the reference's own test_invert_is_pow
// (:1184-1208) pins invert() == pow_vartime(q - 2), which is what is restated here.  0 -> 0 with ok = false.
static inline Scalar fr_invert(const Scalar &a, bool *ok = nullptr) {
  static const u64 QM2[4] = {0xfffffffeffffffffULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
  if (ok) *ok = !fr_is_zero(a);
  return fr_pow_vartime(a, QM2);
}
// src/scalar.rs:256-281 ; returns false when the encoding is not canonical (>= q)
static inline bool fr_from_bytes(const uint8_t b[32], Scalar &out) {
  Scalar t;
  for (int i = 0; i < 4; i++) {
    u64 v = 0;
    for (int k = 7; k >= 0; k--) v = (v << 8) | b[i * 8 + k];
    t.l[i] = v;
  }
  u64 borrow = 0;
  for (int i = 0; i < 4; i++) (void)sbb(t.l[i], FR_MODULUS[i], borrow);
  out = fr_mul(t, FR_R2);
  return (borrow & 1) != 0;
}
// w_n = ROOT_OF_UNITY^(2^(S - log_n)): the primitive 2^log_n-th root of unity every FFT over this field is built
// on (src/scalar.rs:191-205).  The reference exports the constant, not a transform; the transform below is the
// textbook definition  out[k] = sum_j a[j] w_n^(jk)  that downstream provers (bellman's EvaluationDomain) compute.
static inline Scalar fr_omega(int log_n, bool inverse) {
  Scalar w = inverse ? FR_ROOT_OF_UNITY_INV : FR_ROOT_OF_UNITY;
  for (int i = log_n; i < FR_S; i++) w = fr_square(w);
  return w;
}
static inline void fr_dft_naive(const Scalar *a, Scalar *out, size_t n, const Scalar &w) {
  Scalar wk = FR_ONE;  // w^k
  for (size_t k = 0; k < n; k++) {
    Scalar acc = fr_zero(), x = FR_ONE;  // x = w^(jk)
    for (size_t j = 0; j < n; j++) {
      acc = fr_add(acc, fr_mul(a[j], x));
      x = fr_mul(x, wk);
    }
    out[k] = acc;
    wk = fr_mul(wk, w);
  }
}

// ================================================================= G1 (src/g1.rs)
struct G1Affine {
  Fp x, y;
  uint8_t infinity;
};
struct G1Projective {
  Fp x, y, z;
};
// src/g1.rs:199-214
static const Fp G1_GEN_X = {{0x5cb38790fd530c16ULL, 0x7817fc679976fff5ULL, 0x154f95c7143ba1c1ULL,
                             0xf0ae6acdf3d0e747ULL, 0xedce6ecc21dbf440ULL, 0x120177419e0bfb75ULL}};
static const Fp G1_GEN_Y = {{0xbaac93d50ce72271ULL, 0x8c22631a7918fd8eULL, 0xdd595f13570725ceULL,
                             0x51ac582950405194ULL, 0x0e1c8c3fad0059c0ULL, 0x0bbc3efc5008a26aULL}};
// src/g1.rs:176-183  (B = 4 in Montgomery form)
static const Fp G1_B = {{0xaa270000000cfff3ULL, 0x53cc0032fc34000aULL, 0x478fe97a6b0a807fULL,
                         0xb1d37ebee6ba24d7ULL, 0x8ec9733bbf78ab2fULL, 0x09d645513d83de7eULL}};
static inline G1Affine g1a_identity() { return G1Affine{fp_zero(), fp_one(), 1}; }          // src/g1.rs:187-193
static inline G1Affine g1a_generator() { return G1Affine{G1_GEN_X, G1_GEN_Y, 0}; }
static inline G1Projective g1p_identity() { return G1Projective{fp_zero(), fp_one(), fp_zero()}; }  // :605-611
static inline G1Projective g1p_generator() { return G1Projective{G1_GEN_X, G1_GEN_Y, fp_one()}; }
static inline bool g1p_is_identity(const G1Projective &p) { return fp_is_zero(p.z); }
// src/g1.rs:463-471
static inline G1Projective g1p_from_affine(const G1Affine &p) {
  return G1Projective{p.x, p.y, p.infinity ? fp_zero() : fp_one()};
}
// src/g1.rs:103-114
static inline G1Affine g1a_neg(const G1Affine &p) {
  return G1Affine{p.x, p.infinity ? fp_one() : fp_neg(p.y), p.infinity};
}
static inline G1Projective g1p_neg(const G1Projective &p) { return G1Projective{p.x, fp_neg(p.y), p.z}; }
// src/g1.rs:597-601
static inline Fp g1_mul_by_3b(Fp a) {
  a = fp_add(a, a);
  a = fp_add(a, a);
  return fp_add(fp_add(a, a), a);
}
// src/g1.rs:638-667  (RCB Alg. 9)
static inline G1Projective g1p_double(const G1Projective &s) {
  Fp t0 = fp_square(s.y);
  Fp z3 = fp_add(t0, t0);
  z3 = fp_add(z3, z3);
  z3 = fp_add(z3, z3);
  Fp t1 = fp_mul(s.y, s.z);
  Fp t2 = fp_square(s.z);
  t2 = g1_mul_by_3b(t2);
  Fp x3 = fp_mul(t2, z3);
  Fp y3 = fp_add(t0, t2);
  z3 = fp_mul(t1, z3);
  t1 = fp_add(t2, t2);
  t2 = fp_add(t1, t2);
  t0 = fp_sub(t0, t2);
  y3 = fp_mul(t0, y3);
  y3 = fp_add(x3, y3);
  t1 = fp_mul(s.x, s.y);
  x3 = fp_mul(t0, t1);
  x3 = fp_add(x3, x3);
  if (g1p_is_identity(s)) return g1p_identity();
  return G1Projective{x3, y3, z3};
}
// src/g1.rs:670-712  (RCB Alg. 7)
static inline G1Projective g1p_add(const G1Projective &s, const G1Projective &r) {
  Fp t0 = fp_mul(s.x, r.x);
  Fp t1 = fp_mul(s.y, r.y);
  Fp t2 = fp_mul(s.z, r.z);
  Fp t3 = fp_add(s.x, s.y);
  Fp t4 = fp_add(r.x, r.y);
  t3 = fp_mul(t3, t4);
  t4 = fp_add(t0, t1);
  t3 = fp_sub(t3, t4);
  t4 = fp_add(s.y, s.z);
  Fp x3 = fp_add(r.y, r.z);
  t4 = fp_mul(t4, x3);
  x3 = fp_add(t1, t2);
  t4 = fp_sub(t4, x3);
  x3 = fp_add(s.x, s.z);
  Fp y3 = fp_add(r.x, r.z);
  x3 = fp_mul(x3, y3);
  y3 = fp_add(t0, t2);
  y3 = fp_sub(x3, y3);
  x3 = fp_add(t0, t0);
  t0 = fp_add(x3, t0);
  t2 = g1_mul_by_3b(t2);
  Fp z3 = fp_add(t1, t2);
  t1 = fp_sub(t1, t2);
  y3 = g1_mul_by_3b(y3);
  x3 = fp_mul(t4, y3);
  t2 = fp_mul(t3, t1);
  x3 = fp_sub(t2, x3);
  y3 = fp_mul(y3, t0);
  t1 = fp_mul(t1, z3);
  y3 = fp_add(t1, y3);
  t0 = fp_mul(t0, t3);
  z3 = fp_mul(z3, t4);
  z3 = fp_add(z3, t0);
  return G1Projective{x3, y3, z3};
}
// src/g1.rs:715-752  (RCB Alg. 8)
static inline G1Projective g1p_add_mixed(const G1Projective &s, const G1Affine &r) {
  Fp t0 = fp_mul(s.x, r.x);
  Fp t1 = fp_mul(s.y, r.y);
  Fp t3 = fp_add(r.x, r.y);
  Fp t4 = fp_add(s.x, s.y);
  t3 = fp_mul(t3, t4);
  t4 = fp_add(t0, t1);
  t3 = fp_sub(t3, t4);
  t4 = fp_mul(r.y, s.z);
  t4 = fp_add(t4, s.y);
  Fp y3 = fp_mul(r.x, s.z);
  y3 = fp_add(y3, s.x);
  Fp x3 = fp_add(t0, t0);
  t0 = fp_add(x3, t0);
  Fp t2 = g1_mul_by_3b(s.z);
  Fp z3 = fp_add(t1, t2);
  t1 = fp_sub(t1, t2);
  y3 = g1_mul_by_3b(y3);
  x3 = fp_mul(t4, y3);
  t2 = fp_mul(t3, t1);
  x3 = fp_sub(t2, x3);
  y3 = fp_mul(y3, t0);
  t1 = fp_mul(t1, z3);
  y3 = fp_add(t1, y3);
  t0 = fp_mul(t0, t3);
  z3 = fp_mul(z3, t4);
  z3 = fp_add(z3, t0);
  if (r.infinity) return s;
  return G1Projective{x3, y3, z3};
}
// src/g1.rs:754-774 : constant-time double-and-add over a 32-byte LE scalar, bit 255 skipped
static inline G1Projective g1p_multiply(const G1Projective &s, const uint8_t by[32]) {
  G1Projective acc = g1p_identity();
  bool first = true;
  for (int byte = 31; byte >= 0; byte--)
    for (int i = 7; i >= 0; i--) {
      if (first) {
        first = false;
        continue;
      }
      acc = g1p_double(acc);
      G1Projective sum = g1p_add(acc, s);
      if ((by[byte] >> i) & 1) acc = sum;
    }
  return acc;
}
// src/g1.rs:49-63
static inline G1Affine g1a_from_projective(const G1Projective &p) {
  Fp zinv = fp_invert(p.z);  // 0 for z == 0
  if (fp_is_zero(zinv)) return g1a_identity();
  return G1Affine{fp_mul(p.x, zinv), fp_mul(p.y, zinv), 0};
}
// src/g1.rs:806-839
static inline void g1p_batch_normalize(const G1Projective *p, G1Affine *q, size_t n) {
  Fp acc = fp_one();
  for (size_t i = 0; i < n; i++) {
    q[i].x = acc;
    if (!g1p_is_identity(p[i])) acc = fp_mul(acc, p[i].z);
  }
  acc = fp_invert(acc);
  for (size_t i = n; i-- > 0;) {
    bool skip = g1p_is_identity(p[i]);
    Fp tmp = fp_mul(q[i].x, acc);
    if (!skip) acc = fp_mul(acc, p[i].z);
    q[i].x = fp_mul(p[i].x, tmp);
    q[i].y = fp_mul(p[i].y, tmp);
    q[i].infinity = 0;
    if (skip) q[i] = g1a_identity();
  }
}
// src/g1.rs:414-418 : y^2 - x^3 == 4, or infinity
static inline bool g1a_is_on_curve(const G1Affine &p) {
  return p.infinity || fp_eq(fp_sub(fp_square(p.y), fp_mul(fp_square(p.x), p.x)), G1_B);
}
// src/g1.rs:777-793 : [x]P by double-and-add over BLS_X >> 1 starting from the doubled point, then negate (x < 0)
static inline G1Projective g1p_mul_by_x(const G1Projective &s) {
  G1Projective xself = g1p_identity();
  u64 x = 0xd201000000010000ULL >> 1;
  G1Projective tmp = s;
  while (x != 0) {
    tmp = g1p_double(tmp);
    if (x % 2 == 1) xself = g1p_add(xself, tmp);
    x >>= 1;
  }
  return g1p_neg(xself);
}
// src/g1.rs:479-496 (cross-multiplied comparison)
static inline bool g1p_eq(const G1Projective &a, const G1Projective &b) {
  Fp x1 = fp_mul(a.x, b.z), x2 = fp_mul(b.x, a.z), y1 = fp_mul(a.y, b.z), y2 = fp_mul(b.y, a.z);
  bool az = fp_is_zero(a.z), bz = fp_is_zero(b.z);
  return (az && bz) || (!az && !bz && fp_eq(x1, x2) && fp_eq(y1, y2));
}
// src/g1.rs:421-428 BETA ; :430-437 endomorphism ; :401-410 is_torsion_free: endomorphism(P) == -[x^2]P
static const Fp G1_BETA = {{0x30f1361b798a64e8ULL, 0xf3b8ddab7ece5a2aULL, 0x16a8ca3ac61577f7ULL,
                            0xc26a2ff874fd029bULL, 0x3636b76660701c6eULL, 0x051ba4ab241b6160ULL}};
static inline bool g1a_is_torsion_free(const G1Affine &p) {
  G1Projective m = g1p_neg(g1p_mul_by_x(g1p_mul_by_x(g1p_from_affine(p))));
  G1Affine e = p;
  e.x = fp_mul(e.x, G1_BETA);
  return g1p_eq(m, g1p_from_affine(e));
}
// src/g1.rs:221-260
static inline void g1a_to_compressed(const G1Affine &p, uint8_t out[48]) {
  fp_to_bytes(p.infinity ? fp_zero() : p.x, out);
  out[0] |= 0x80;
  if (p.infinity) out[0] |= 0x40;
  if (!p.infinity && fp_lexicographically_largest(p.y)) out[0] |= 0x20;
}
static inline void g1a_to_uncompressed(const G1Affine &p, uint8_t out[96]) {
  fp_to_bytes(p.infinity ? fp_zero() : p.x, out);
  fp_to_bytes(p.infinity ? fp_zero() : p.y, out + 48);
  if (p.infinity) out[0] |= 0x40;
}
// src/g1.rs:275-320 (from_uncompressed_unchecked + on-curve; torsion check is off-path)
static inline bool g1a_from_uncompressed(const uint8_t in[96], G1Affine &out) {
  bool compression = (in[0] >> 7) & 1, infinity = (in[0] >> 6) & 1, sort = (in[0] >> 5) & 1;
  uint8_t tmp[48];
  std::memcpy(tmp, in, 48);
  tmp[0] &= 0x1f;
  Fp x, y;
  bool okx = fp_from_bytes(tmp, x), oky = fp_from_bytes(in + 48, y);
  if (!okx || !oky) return false;
  if (infinity) {
    if (compression || sort || !fp_is_zero(x) || !fp_is_zero(y)) return false;
    out = g1a_identity();
    return true;
  }
  if (compression || sort) return false;
  out = G1Affine{x, y, 0};
  return g1a_is_on_curve(out);
}
// src/g1.rs:330-390
static inline bool g1a_from_compressed(const uint8_t in[48], G1Affine &out) {
  bool compression = (in[0] >> 7) & 1, infinity = (in[0] >> 6) & 1, sort = (in[0] >> 5) & 1;
  uint8_t tmp[48];
  std::memcpy(tmp, in, 48);
  tmp[0] &= 0x1f;
  Fp x;
  if (!fp_from_bytes(tmp, x)) return false;
  if (infinity) {
    if (!compression || sort || !fp_is_zero(x)) return false;
    out = g1a_identity();
    return true;
  }
  Fp y;
  if (!fp_sqrt(fp_add(fp_mul(fp_square(x), x), G1_B), y)) return false;
  if (fp_lexicographically_largest(y) != sort) y = fp_neg(y);
  out = G1Affine{x, y, 0};
  return compression;
}

// ================================================================= G2 (src/g2.rs)
struct G2Affine {
  Fp2 x, y;
  uint8_t infinity;
};
struct G2Projective {
  Fp2 x, y, z;
};
// src/g2.rs:177-194
static const Fp2 G2_B = {G1_B, G1_B};
// src/g2.rs:212-247
static const Fp2 G2_GEN_X = {{{0xf5f28fa202940a10ULL, 0xb3f5fb2687b4961aULL, 0xa1a893b53e2ae580ULL,
                               0x9894999d1a3caee9ULL, 0x6f67b7631863366bULL, 0x058191924350bcd7ULL}},
                             {{0xa5a9c0759e23f606ULL, 0xaaa0c59dbccd60c3ULL, 0x3bb17e18e2867806ULL,
                               0x1b1ab6cc8541b367ULL, 0xc2b6ed0ef2158547ULL, 0x11922a097360edf3ULL}}};
static const Fp2 G2_GEN_Y = {{{0x4c730af860494c4aULL, 0x597cfa1f5e369c5aULL, 0xe7e6856caa0a635aULL,
                               0xbbefb5e96e0d495fULL, 0x07d3a975f0ef25a2ULL, 0x0083fd8e7e80dae5ULL}},
                             {{0xadc0fc92df64b05dULL, 0x18aa270a2b1461dcULL, 0x86adac6a3be4eba0ULL,
                               0x79495c4ec93da33aULL, 0xe7175850a43ccaedULL, 0x0b2bc2a163de1bf2ULL}}};
static inline G2Affine g2a_identity() { return G2Affine{fp2_zero(), fp2_one(), 1}; }  // src/g2.rs:199-205
static inline G2Affine g2a_generator() { return G2Affine{G2_GEN_X, G2_GEN_Y, 0}; }
static inline G2Projective g2p_identity() { return G2Projective{fp2_zero(), fp2_one(), fp2_zero()}; }
static inline G2Projective g2p_generator() { return G2Projective{G2_GEN_X, G2_GEN_Y, fp2_one()}; }
static inline bool g2p_is_identity(const G2Projective &p) { return fp2_is_zero(p.z); }
static inline G2Projective g2p_from_affine(const G2Affine &p) {  // src/g2.rs:516-525
  return G2Projective{p.x, p.y, p.infinity ? fp2_zero() : fp2_one()};
}
static inline G2Affine g2a_neg(const G2Affine &p) {  // src/g2.rs:104-115
  return G2Affine{p.x, p.infinity ? fp2_one() : fp2_neg(p.y), p.infinity};
}
static inline G2Projective g2p_neg(const G2Projective &p) { return G2Projective{p.x, fp2_neg(p.y), p.z}; }
// src/g2.rs:196 (B3 = B+B+B) and :650-652
static inline Fp2 g2_mul_by_3b(const Fp2 &x) {
  static const Fp2 B3 = fp2_add(fp2_add(G2_B, G2_B), G2_B);
  return fp2_mul(x, B3);
}
// src/g2.rs:709-738
static inline G2Projective g2p_double(const G2Projective &s) {
  Fp2 t0 = fp2_square(s.y);
  Fp2 z3 = fp2_add(t0, t0);
  z3 = fp2_add(z3, z3);
  z3 = fp2_add(z3, z3);
  Fp2 t1 = fp2_mul(s.y, s.z);
  Fp2 t2 = fp2_square(s.z);
  t2 = g2_mul_by_3b(t2);
  Fp2 x3 = fp2_mul(t2, z3);
  Fp2 y3 = fp2_add(t0, t2);
  z3 = fp2_mul(t1, z3);
  t1 = fp2_add(t2, t2);
  t2 = fp2_add(t1, t2);
  t0 = fp2_sub(t0, t2);
  y3 = fp2_mul(t0, y3);
  y3 = fp2_add(x3, y3);
  t1 = fp2_mul(s.x, s.y);
  x3 = fp2_mul(t0, t1);
  x3 = fp2_add(x3, x3);
  if (g2p_is_identity(s)) return g2p_identity();
  return G2Projective{x3, y3, z3};
}
// src/g2.rs:741-783
static inline G2Projective g2p_add(const G2Projective &s, const G2Projective &r) {
  Fp2 t0 = fp2_mul(s.x, r.x);
  Fp2 t1 = fp2_mul(s.y, r.y);
  Fp2 t2 = fp2_mul(s.z, r.z);
  Fp2 t3 = fp2_add(s.x, s.y);
  Fp2 t4 = fp2_add(r.x, r.y);
  t3 = fp2_mul(t3, t4);
  t4 = fp2_add(t0, t1);
  t3 = fp2_sub(t3, t4);
  t4 = fp2_add(s.y, s.z);
  Fp2 x3 = fp2_add(r.y, r.z);
  t4 = fp2_mul(t4, x3);
  x3 = fp2_add(t1, t2);
  t4 = fp2_sub(t4, x3);
  x3 = fp2_add(s.x, s.z);
  Fp2 y3 = fp2_add(r.x, r.z);
  x3 = fp2_mul(x3, y3);
  y3 = fp2_add(t0, t2);
  y3 = fp2_sub(x3, y3);
  x3 = fp2_add(t0, t0);
  t0 = fp2_add(x3, t0);
  t2 = g2_mul_by_3b(t2);
  Fp2 z3 = fp2_add(t1, t2);
  t1 = fp2_sub(t1, t2);
  y3 = g2_mul_by_3b(y3);
  x3 = fp2_mul(t4, y3);
  t2 = fp2_mul(t3, t1);
  x3 = fp2_sub(t2, x3);
  y3 = fp2_mul(y3, t0);
  t1 = fp2_mul(t1, z3);
  y3 = fp2_add(t1, y3);
  t0 = fp2_mul(t0, t3);
  z3 = fp2_mul(z3, t4);
  z3 = fp2_add(z3, t0);
  return G2Projective{x3, y3, z3};
}
// src/g2.rs:786-823
static inline G2Projective g2p_add_mixed(const G2Projective &s, const G2Affine &r) {
  Fp2 t0 = fp2_mul(s.x, r.x);
  Fp2 t1 = fp2_mul(s.y, r.y);
  Fp2 t3 = fp2_add(r.x, r.y);
  Fp2 t4 = fp2_add(s.x, s.y);
  t3 = fp2_mul(t3, t4);
  t4 = fp2_add(t0, t1);
  t3 = fp2_sub(t3, t4);
  t4 = fp2_mul(r.y, s.z);
  t4 = fp2_add(t4, s.y);
  Fp2 y3 = fp2_mul(r.x, s.z);
  y3 = fp2_add(y3, s.x);
  Fp2 x3 = fp2_add(t0, t0);
  t0 = fp2_add(x3, t0);
  Fp2 t2 = g2_mul_by_3b(s.z);
  Fp2 z3 = fp2_add(t1, t2);
  t1 = fp2_sub(t1, t2);
  y3 = g2_mul_by_3b(y3);
  x3 = fp2_mul(t4, y3);
  t2 = fp2_mul(t3, t1);
  x3 = fp2_sub(t2, x3);
  y3 = fp2_mul(y3, t0);
  t1 = fp2_mul(t1, z3);
  y3 = fp2_add(t1, y3);
  t0 = fp2_mul(t0, t3);
  z3 = fp2_mul(z3, t4);
  z3 = fp2_add(z3, t0);
  if (r.infinity) return s;
  return G2Projective{x3, y3, z3};
}
// src/g2.rs:825-845
static inline G2Projective g2p_multiply(const G2Projective &s, const uint8_t by[32]) {
  G2Projective acc = g2p_identity();
  bool first = true;
  for (int byte = 31; byte >= 0; byte--)
    for (int i = 7; i >= 0; i--) {
      if (first) {
        first = false;
        continue;
      }
      acc = g2p_double(acc);
      G2Projective sum = g2p_add(acc, s);
      if ((by[byte] >> i) & 1) acc = sum;
    }
  return acc;
}
// src/g2.rs:50-64
static inline G2Affine g2a_from_projective(const G2Projective &p) {
  Fp2 zinv = fp2_invert(p.z);
  if (fp2_is_zero(zinv)) return g2a_identity();
  return G2Affine{fp2_mul(p.x, zinv), fp2_mul(p.y, zinv), 0};
}
// src/g2.rs:951-984
static inline void g2p_batch_normalize(const G2Projective *p, G2Affine *q, size_t n) {
  Fp2 acc = fp2_one();
  for (size_t i = 0; i < n; i++) {
    q[i].x = acc;
    if (!g2p_is_identity(p[i])) acc = fp2_mul(acc, p[i].z);
  }
  acc = fp2_invert(acc);
  for (size_t i = n; i-- > 0;) {
    bool skip = g2p_is_identity(p[i]);
    Fp2 tmp = fp2_mul(q[i].x, acc);
    if (!skip) acc = fp2_mul(acc, p[i].z);
    q[i].x = fp2_mul(p[i].x, tmp);
    q[i].y = fp2_mul(p[i].y, tmp);
    q[i].infinity = 0;
    if (skip) q[i] = g2a_identity();
  }
}
// src/g2.rs:915-932
static inline G2Projective g2p_mul_by_x(const G2Projective &s) {
  G2Projective xself = g2p_identity();
  u64 x = 0xd201000000010000ULL >> 1;
  G2Projective acc = s;
  while (x != 0) {
    acc = g2p_double(acc);
    if (x % 2 == 1) xself = g2p_add(xself, acc);
    x >>= 1;
  }
  return g2p_neg(xself);
}
// src/g2.rs:847-888
static inline G2Projective g2p_psi(const G2Projective &s) {
  static const Fp2 cx = {{{0, 0, 0, 0, 0, 0}},
                         {{0x890dc9e4867545c3ULL, 0x2af322533285a5d5ULL, 0x50880866309b7e2cULL, 0xa20d1b8c7e881024ULL,
                           0x14e4f04fe2db9068ULL, 0x14e56d3f1564853aULL}}};
  static const Fp2 cy = {{{0x3e2f585da55c9ad1ULL, 0x4294213d86c18183ULL, 0x382844c88b623732ULL, 0x92ad2afd19103e18ULL,
                           0x1d794e4fac7cf0b9ULL, 0x0bd592fc7d825ec8ULL}},
                         {{0x7bcfa7a25aa30fdaULL, 0xdc17dec12a927e7cULL, 0x2f088dd86b4ebef1ULL, 0xd1ca2087da74d4a7ULL,
                           0x2da2596696cebc1dULL, 0x0e2b7eedbbfd87d2ULL}}};
  return G2Projective{fp2_mul(fp2_frobenius_map(s.x), cx), fp2_mul(fp2_frobenius_map(s.y), cy), fp2_frobenius_map(s.z)};
}
// src/g2.rs:537-554
static inline bool g2p_eq(const G2Projective &a, const G2Projective &b) {
  Fp2 x1 = fp2_mul(a.x, b.z), x2 = fp2_mul(b.x, a.z), y1 = fp2_mul(a.y, b.z), y2 = fp2_mul(b.y, a.z);
  bool az = fp2_is_zero(a.z), bz = fp2_is_zero(b.z);
  return (az && bz) || (!az && !bz && fp2_eq(x1, x2) && fp2_eq(y1, y2));
}
// src/g2.rs:475-482 : psi(P) == [x]P
static inline bool g2a_is_torsion_free(const G2Affine &p) {
  G2Projective pp = g2p_from_affine(p);
  return g2p_eq(g2p_psi(pp), g2p_mul_by_x(pp));
}
static inline bool g2a_is_on_curve(const G2Affine &p) {  // src/g2.rs:487-491
  return p.infinity || fp2_eq(fp2_sub(fp2_square(p.y), fp2_mul(fp2_square(p.x), p.x)), G2_B);
}
// src/g2.rs:254-299 ; Fp2 is serialized c1 || c0
static inline void g2a_to_compressed(const G2Affine &p, uint8_t out[96]) {
  Fp2 x = p.infinity ? fp2_zero() : p.x;
  fp_to_bytes(x.c1, out);
  fp_to_bytes(x.c0, out + 48);
  out[0] |= 0x80;
  if (p.infinity) out[0] |= 0x40;
  if (!p.infinity && fp2_lexicographically_largest(p.y)) out[0] |= 0x20;
}
static inline void g2a_to_uncompressed(const G2Affine &p, uint8_t out[192]) {
  Fp2 x = p.infinity ? fp2_zero() : p.x, y = p.infinity ? fp2_zero() : p.y;
  fp_to_bytes(x.c1, out);
  fp_to_bytes(x.c0, out + 48);
  fp_to_bytes(y.c1, out + 96);
  fp_to_bytes(y.c0, out + 144);
  if (p.infinity) out[0] |= 0x40;
}
// src/g2.rs:313-374
static inline bool g2a_from_uncompressed(const uint8_t in[192], G2Affine &out) {
  bool compression = (in[0] >> 7) & 1, infinity = (in[0] >> 6) & 1, sort = (in[0] >> 5) & 1;
  uint8_t tmp[48];
  std::memcpy(tmp, in, 48);
  tmp[0] &= 0x1f;
  Fp xc1, xc0, yc1, yc0;
  bool ok = fp_from_bytes(tmp, xc1);
  ok &= fp_from_bytes(in + 48, xc0);
  ok &= fp_from_bytes(in + 96, yc1);
  ok &= fp_from_bytes(in + 144, yc0);
  if (!ok) return false;
  Fp2 x = {xc0, xc1}, y = {yc0, yc1};
  if (infinity) {
    if (compression || sort || !fp2_is_zero(x) || !fp2_is_zero(y)) return false;
    out = g2a_identity();
    return true;
  }
  if (compression || sort) return false;
  out = G2Affine{x, y, 0};
  return g2a_is_on_curve(out);
}
// src/g2.rs:384-464
static inline bool g2a_from_compressed(const uint8_t in[96], G2Affine &out) {
  bool compression = (in[0] >> 7) & 1, infinity = (in[0] >> 6) & 1, sort = (in[0] >> 5) & 1;
  uint8_t tmp[48];
  std::memcpy(tmp, in, 48);
  tmp[0] &= 0x1f;
  Fp xc1, xc0;
  if (!fp_from_bytes(tmp, xc1) || !fp_from_bytes(in + 48, xc0)) return false;
  Fp2 x = {xc0, xc1};
  if (infinity) {
    if (!compression || sort || !fp2_is_zero(x)) return false;
    out = g2a_identity();
    return true;
  }
  Fp2 y;
  if (!fp2_sqrt(fp2_add(fp2_mul(fp2_square(x), x), G2_B), y)) return false;
  if (fp2_lexicographically_largest(y) != sort) y = fp2_neg(y);
  out = G2Affine{x, y, 0};
  return compression;
}

// ================================================================= pairings (src/pairings.rs, src/lib.rs:72-74)
static const u64 BLS_X = 0xd201000000010000ULL;
static const bool BLS_X_IS_NEGATIVE = true;
struct LineCoeffs {
  Fp2 a, b, c;
};
// src/pairings.rs:709-738
static inline LineCoeffs doubling_step(G2Projective &r) {
  Fp2 tmp0 = fp2_square(r.x);
  Fp2 tmp1 = fp2_square(r.y);
  Fp2 tmp2 = fp2_square(tmp1);
  Fp2 tmp3 = fp2_sub(fp2_sub(fp2_square(fp2_add(tmp1, r.x)), tmp0), tmp2);
  tmp3 = fp2_add(tmp3, tmp3);
  Fp2 tmp4 = fp2_add(fp2_add(tmp0, tmp0), tmp0);
  Fp2 tmp6 = fp2_add(r.x, tmp4);
  Fp2 tmp5 = fp2_square(tmp4);
  Fp2 zsquared = fp2_square(r.z);
  r.x = fp2_sub(fp2_sub(tmp5, tmp3), tmp3);
  r.z = fp2_sub(fp2_sub(fp2_square(fp2_add(r.z, r.y)), tmp1), zsquared);
  r.y = fp2_mul(fp2_sub(tmp3, r.x), tmp4);
  tmp2 = fp2_add(tmp2, tmp2);
  tmp2 = fp2_add(tmp2, tmp2);
  tmp2 = fp2_add(tmp2, tmp2);
  r.y = fp2_sub(r.y, tmp2);
  tmp3 = fp2_mul(tmp4, zsquared);
  tmp3 = fp2_add(tmp3, tmp3);
  tmp3 = fp2_neg(tmp3);
  tmp6 = fp2_sub(fp2_sub(fp2_square(tmp6), tmp0), tmp5);
  tmp1 = fp2_add(tmp1, tmp1);
  tmp1 = fp2_add(tmp1, tmp1);
  tmp6 = fp2_sub(tmp6, tmp1);
  tmp0 = fp2_mul(r.z, zsquared);
  tmp0 = fp2_add(tmp0, tmp0);
  return LineCoeffs{tmp0, tmp3, tmp6};
}
// src/pairings.rs:740-770
static inline LineCoeffs addition_step(G2Projective &r, const G2Affine &q) {
  Fp2 zsquared = fp2_square(r.z);
  Fp2 ysquared = fp2_square(q.y);
  Fp2 t0 = fp2_mul(zsquared, q.x);
  Fp2 t1 = fp2_mul(fp2_sub(fp2_sub(fp2_square(fp2_add(q.y, r.z)), ysquared), zsquared), zsquared);
  Fp2 t2 = fp2_sub(t0, r.x);
  Fp2 t3 = fp2_square(t2);
  Fp2 t4 = fp2_add(t3, t3);
  t4 = fp2_add(t4, t4);
  Fp2 t5 = fp2_mul(t4, t2);
  Fp2 t6 = fp2_sub(fp2_sub(t1, r.y), r.y);
  Fp2 t9 = fp2_mul(t6, q.x);
  Fp2 t7 = fp2_mul(t4, r.x);
  r.x = fp2_sub(fp2_sub(fp2_sub(fp2_square(t6), t5), t7), t7);
  r.z = fp2_sub(fp2_sub(fp2_square(fp2_add(r.z, t2)), zsquared), t3);
  Fp2 t10 = fp2_add(q.y, r.z);
  Fp2 t8 = fp2_mul(fp2_sub(t7, r.x), t6);
  t0 = fp2_mul(r.y, t5);
  t0 = fp2_add(t0, t0);
  r.y = fp2_sub(t8, t0);
  t10 = fp2_sub(fp2_square(t10), ysquared);
  Fp2 ztsquared = fp2_square(r.z);
  t10 = fp2_sub(t10, ztsquared);
  t9 = fp2_sub(fp2_add(t9, t9), t10);
  t10 = fp2_add(r.z, r.z);
  t6 = fp2_neg(t6);
  t1 = fp2_add(t6, t6);
  return LineCoeffs{t10, t1, t9};
}
// src/pairings.rs:696-707
static inline Fp12 ell(const Fp12 &f, const LineCoeffs &co, const G1Affine &p) {
  Fp2 c0 = co.a, c1 = co.b;
  c0.c0 = fp_mul(c0.c0, p.y);
  c0.c1 = fp_mul(c0.c1, p.y);
  c1.c0 = fp_mul(c1.c0, p.x);
  c1.c1 = fp_mul(c1.c1, p.x);
  return fp12_mul_by_014(f, co.c, c1, c0);
}
// src/pairings.rs:668-694, generic driver; D supplies doubling_step/addition_step on f
template <class D>
static inline Fp12 miller_loop_generic(D &d) {
  Fp12 f = fp12_one();
  bool found_one = false;
  for (int b = 63; b >= 0; b--) {
    bool i = (((BLS_X >> 1) >> b) & 1) == 1;
    if (!found_one) {
      found_one = i;
      continue;
    }
    f = d.doubling(f);
    if (i) f = d.addition(f);
    f = fp12_square(f);
  }
  f = d.doubling(f);
  if (BLS_X_IS_NEGATIVE) f = fp12_conjugate(f);
  return f;
}
// src/pairings.rs:504-546
struct G2Prepared {
  uint8_t infinity;
  std::vector<LineCoeffs> coeffs;
};
static inline G2Prepared g2_prepare(const G2Affine &qin) {
  struct Drv {
    G2Projective cur;
    G2Affine base;
    std::vector<LineCoeffs> *out;
    Fp12 doubling(const Fp12 &f) {
      out->push_back(doubling_step(cur));
      return f;
    }
    Fp12 addition(const Fp12 &f) {
      out->push_back(addition_step(cur, base));
      return f;
    }
  };
  G2Prepared r;
  r.infinity = qin.infinity;
  G2Affine q = qin.infinity ? g2a_generator() : qin;
  Drv d{g2p_from_affine(q), q, &r.coeffs};
  (void)miller_loop_generic(d);
  return r;
}
// src/pairings.rs:554-603
static inline Fp12 multi_miller_loop(const G1Affine *ps, const G2Prepared *qs, size_t n) {
  struct Drv {
    const G1Affine *ps;
    const G2Prepared *qs;
    size_t n, index;
    Fp12 step(Fp12 f) {
      for (size_t t = 0; t < n; t++) {
        bool either = ps[t].infinity || qs[t].infinity;
        Fp12 nf = ell(f, qs[t].coeffs[index], ps[t]);
        if (!either) f = nf;
      }
      index++;
      return f;
    }
    Fp12 doubling(const Fp12 &f) { return step(f); }
    Fp12 addition(const Fp12 &f) { return step(f); }
  };
  Drv d{ps, qs, n, 0};
  return miller_loop_generic(d);
}
// the Miller loop of src/pairings.rs:607-646 (unprepared, single pair; identity handling included)
static inline Fp12 miller_loop_pair(const G1Affine &pin, const G2Affine &qin) {
  struct Drv {
    G2Projective cur;
    G2Affine base;
    G1Affine p;
    Fp12 doubling(const Fp12 &f) {
      LineCoeffs c = doubling_step(cur);
      return ell(f, c, p);
    }
    Fp12 addition(const Fp12 &f) {
      LineCoeffs c = addition_step(cur, base);
      return ell(f, c, p);
    }
  };
  bool either = pin.infinity || qin.infinity;
  G1Affine p = either ? g1a_generator() : pin;
  G2Affine q = either ? g2a_generator() : qin;
  Drv d{g2p_from_affine(q), q, p};
  Fp12 tmp = miller_loop_generic(d);
  return either ? fp12_one() : tmp;
}
// src/pairings.rs:48-176
static inline void fp4_square(const Fp2 &a, const Fp2 &b, Fp2 &c0, Fp2 &c1) {
  Fp2 t0 = fp2_square(a), t1 = fp2_square(b);
  Fp2 t2 = fp2_mul_by_nonresidue(t1);
  c0 = fp2_add(t2, t0);
  t2 = fp2_add(a, b);
  t2 = fp2_square(t2);
  t2 = fp2_sub(t2, t0);
  c1 = fp2_sub(t2, t1);
}
static inline Fp12 cyclotomic_square(const Fp12 &f) {
  Fp2 z0 = f.c0.c0, z4 = f.c0.c1, z3 = f.c0.c2, z2 = f.c1.c0, z1 = f.c1.c1, z5 = f.c1.c2;
  Fp2 t0, t1, t2, t3;
  fp4_square(z0, z1, t0, t1);
  z0 = fp2_sub(t0, z0);
  z0 = fp2_add(fp2_add(z0, z0), t0);
  z1 = fp2_add(t1, z1);
  z1 = fp2_add(fp2_add(z1, z1), t1);
  fp4_square(z2, z3, t0, t1);
  fp4_square(z4, z5, t2, t3);
  z4 = fp2_sub(t0, z4);
  z4 = fp2_add(fp2_add(z4, z4), t0);
  z5 = fp2_add(t1, z5);
  z5 = fp2_add(fp2_add(z5, z5), t1);
  t0 = fp2_mul_by_nonresidue(t3);
  z2 = fp2_add(t0, z2);
  z2 = fp2_add(fp2_add(z2, z2), t0);
  z3 = fp2_sub(t2, z3);
  z3 = fp2_add(fp2_add(z3, z3), t2);
  return Fp12{Fp6{z0, z4, z3}, Fp6{z2, z1, z5}};
}
static inline Fp12 cyclotomic_exp(const Fp12 &f) {
  Fp12 tmp = fp12_one();
  bool found_one = false;
  for (int b = 63; b >= 0; b--) {
    bool i = ((BLS_X >> b) & 1) == 1;
    if (found_one)
      tmp = cyclotomic_square(tmp);
    else
      found_one = i;
    if (i) tmp = fp12_mul(tmp, f);
  }
  return fp12_conjugate(tmp);
}
static inline Fp12 final_exponentiation(const Fp12 &fin) {
  Fp12 f = fin;
  Fp12 t0 = f;
  for (int i = 0; i < 6; i++) t0 = fp12_frobenius_map(t0);
  Fp12 t1 = fp12_invert(f);
  Fp12 t2 = fp12_mul(t0, t1);
  t1 = t2;
  t2 = fp12_frobenius_map(fp12_frobenius_map(t2));
  t2 = fp12_mul(t2, t1);
  t1 = fp12_conjugate(cyclotomic_square(t2));
  Fp12 t3 = cyclotomic_exp(t2);
  Fp12 t4 = cyclotomic_square(t3);
  Fp12 t5 = fp12_mul(t1, t3);
  t1 = cyclotomic_exp(t5);
  t0 = cyclotomic_exp(t1);
  Fp12 t6 = cyclotomic_exp(t0);
  t6 = fp12_mul(t6, t4);
  t4 = cyclotomic_exp(t6);
  t5 = fp12_conjugate(t5);
  t4 = fp12_mul(t4, fp12_mul(t5, t2));
  t5 = fp12_conjugate(t2);
  t1 = fp12_mul(t1, t2);
  t1 = fp12_frobenius_map(fp12_frobenius_map(fp12_frobenius_map(t1)));
  t6 = fp12_mul(t6, t5);
  t6 = fp12_frobenius_map(t6);
  t3 = fp12_mul(t3, t0);
  t3 = fp12_frobenius_map(fp12_frobenius_map(t3));
  t3 = fp12_mul(t3, t1);
  t3 = fp12_mul(t3, t6);
  f = fp12_mul(t3, t4);
  return f;
}
// src/pairings.rs:607-653
static inline Fp12 pairing(const G1Affine &p, const G2Affine &q) {
  return final_exponentiation(miller_loop_pair(p, q));
}

}  // namespace bls_oracle
