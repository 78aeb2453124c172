// Fp6 = Fp2[v]/(v^3-(u+1)) and Fp12 = Fp6[w]/(w^2-v) on the device.
// Replaces src/fp6.rs (mul :200, square :277, mul_by_1 :113, mul_by_01 :121, mul_by_nonresidue :139,
// frobenius_map :154, invert :294) and src/fp12.rs (Mul :197, square :174, mul_by_014 :116,
// conjugate :136, frobenius_map :145, invert :187).  All values canonical => bit-identical.
//
// An Fp12 is 144 32-bit words — it cannot live in registers next to anything else — so tower values
// are addressed through pointers (per-thread local memory or shared memory) and only Fp2 operands are
// pulled into registers; the Fp2 mul/sqr are deliberately NOT inlined (one copy of the 3x305-IMAD
// body in the instruction cache instead of one per call site).
#pragma once
#include "constants.cuh"
#include "fp2.cuh"

namespace b200 {

struct fp6 {
  fp2 c0, c1, c2;
};
struct fp12 {
  fp6 c0, c1;
};

// ---------------------------------------------------------------- Fp6
B200_DEV void fp6_add(fp6 *r, const fp6 *a, const fp6 *b) {
  r->c0 = fp2_add(a->c0, b->c0);
  r->c1 = fp2_add(a->c1, b->c1);
  r->c2 = fp2_add(a->c2, b->c2);
}
B200_DEV void fp6_sub(fp6 *r, const fp6 *a, const fp6 *b) {
  r->c0 = fp2_sub(a->c0, b->c0);
  r->c1 = fp2_sub(a->c1, b->c1);
  r->c2 = fp2_sub(a->c2, b->c2);
}
B200_DEV void fp6_neg(fp6 *r, const fp6 *a) {
  r->c0 = fp2_neg(a->c0);
  r->c1 = fp2_neg(a->c1);
  r->c2 = fp2_neg(a->c2);
}
// (c0,c1,c2) -> (xi*c2, c0, c1)      src/fp6.rs:139-150   (r may alias a)
B200_DEV void fp6_mul_by_nonresidue(fp6 *r, const fp6 *a) {
  fp2 t = fp2_mul_by_nonresidue(a->c2);
  fp2 a0 = a->c0, a1 = a->c1;
  r->c0 = t;
  r->c1 = a0;
  r->c2 = a1;
}
// Karatsuba/Toom over Fp2: 6 Fp2 muls (the reference's interleaved schoolbook computes the same element)
B200_NOINL void fp6_mul(fp6 *r, const fp6 *a, const fp6 *b) {
  fp2 v0 = M2(a->c0, b->c0), v1 = M2(a->c1, b->c1), v2 = M2(a->c2, b->c2);
  fp2 t0 = M2(fp2_add(a->c1, a->c2), fp2_add(b->c1, b->c2));
  fp2 t1 = M2(fp2_add(a->c0, a->c1), fp2_add(b->c0, b->c1));
  fp2 t2 = M2(fp2_add(a->c0, a->c2), fp2_add(b->c0, b->c2));
  t0 = fp2_sub(fp2_sub(t0, v1), v2);
  t1 = fp2_sub(fp2_sub(t1, v0), v1);
  t2 = fp2_sub(fp2_sub(t2, v0), v2);
  r->c0 = fp2_add(v0, fp2_mul_by_nonresidue(t0));
  r->c1 = fp2_add(t1, fp2_mul_by_nonresidue(v2));
  r->c2 = fp2_add(t2, v1);
}
// src/fp6.rs:277-291
B200_NOINL void fp6_sqr(fp6 *r, const fp6 *a) {
  fp2 s0 = S2(a->c0);
  fp2 ab = M2(a->c0, a->c1);
  fp2 s1 = fp2_dbl(ab);
  fp2 s2 = S2(fp2_add(fp2_sub(a->c0, a->c1), a->c2));
  fp2 bc = M2(a->c1, a->c2);
  fp2 s3 = fp2_dbl(bc);
  fp2 s4 = S2(a->c2);
  r->c0 = fp2_add(fp2_mul_by_nonresidue(s3), s0);
  r->c1 = fp2_add(fp2_mul_by_nonresidue(s4), s1);
  r->c2 = fp2_sub(fp2_sub(fp2_add(fp2_add(s1, s2), s3), s0), s4);
}
// src/fp6.rs:113-119
B200_DEV void fp6_mul_by_1(fp6 *r, const fp6 *a, const fp2 *c1) {
  fp2 t0 = fp2_mul_by_nonresidue(M2(a->c2, *c1));
  fp2 t1 = M2(a->c0, *c1);
  fp2 t2 = M2(a->c1, *c1);
  r->c0 = t0;
  r->c1 = t1;
  r->c2 = t2;
}
// src/fp6.rs:121-136
B200_DEV void fp6_mul_by_01(fp6 *r, const fp6 *a, const fp2 *c0, const fp2 *c1) {
  fp2 a_a = M2(a->c0, *c0);
  fp2 b_b = M2(a->c1, *c1);
  fp2 t1 = fp2_add(fp2_mul_by_nonresidue(M2(a->c2, *c1)), a_a);
  fp2 t2 = fp2_sub(fp2_sub(M2(fp2_add(*c0, *c1), fp2_add(a->c0, a->c1)), a_a), b_b);
  fp2 t3 = fp2_add(M2(a->c2, *c0), b_b);
  r->c0 = t1;
  r->c1 = t2;
  r->c2 = t3;
}
// src/fp6.rs:154-188 ; the two coefficients are (0 + k u) and (k + 0 u): 2 Fp muls each
B200_DEV void fp6_frobenius(fp6 *r, const fp6 *a) {
  fp2 c0 = fp2_conj(a->c0), c1 = fp2_conj(a->c1), c2 = fp2_conj(a->c2);
  fp k1 = fp_const(K_FROB6_C1_U), k2 = fp_const(K_FROB6_C2_R);
  fp t0, t1;
  // (x + y u)(k u) = -y k + x k u
  fp_mul_ni(&t0, &c1.c1, &k1);
  fp_mul_ni(&t1, &c1.c0, &k1);
  r->c0 = c0;
  r->c1 = fp2{fp_neg(t0), t1};
  fp_mul_ni(&t0, &c2.c0, &k2);
  fp_mul_ni(&t1, &c2.c1, &k2);
  r->c2 = fp2{t0, t1};
}
// src/fp6.rs:294-312
B200_NOINL void fp6_inv(fp6 *r, const fp6 *a) {
  fp2 c0 = fp2_sub(S2(a->c0), fp2_mul_by_nonresidue(M2(a->c1, a->c2)));
  fp2 c1 = fp2_sub(fp2_mul_by_nonresidue(S2(a->c2)), M2(a->c0, a->c1));
  fp2 c2 = fp2_sub(S2(a->c1), M2(a->c0, a->c2));
  fp2 tmp = fp2_mul_by_nonresidue(fp2_add(M2(a->c1, c2), M2(a->c2, c1)));
  tmp = fp2_add(tmp, M2(a->c0, c0));
  fp2 t = fp2_inv_ni(tmp);
  r->c0 = M2(t, c0);
  r->c1 = M2(t, c1);
  r->c2 = M2(t, c2);
}

// ---------------------------------------------------------------- Fp12
B200_DEV void fp12_set_one(fp12 *r) {
  r->c0.c0 = fp2_one();
  r->c0.c1 = fp2_zero();
  r->c0.c2 = fp2_zero();
  r->c1.c0 = fp2_zero();
  r->c1.c1 = fp2_zero();
  r->c1.c2 = fp2_zero();
}
// src/fp12.rs:136-141  (r may alias a)
B200_DEV void fp12_conj(fp12 *r, const fp12 *a) {
  r->c0 = a->c0;
  fp6_neg(&r->c1, &a->c1);
}
// src/fp12.rs:197-214   (r may alias a or b)
B200_NOINL void fp12_mul(fp12 *r, const fp12 *a, const fp12 *b) {
  fp6 aa, bb, o, c1;
  fp6_mul(&aa, &a->c0, &b->c0);
  fp6_mul(&bb, &a->c1, &b->c1);
  fp6_add(&o, &b->c0, &b->c1);
  fp6_add(&c1, &a->c1, &a->c0);
  fp6_mul(&c1, &c1, &o);
  fp6_sub(&c1, &c1, &aa);
  fp6_sub(&c1, &c1, &bb);
  fp6_mul_by_nonresidue(&bb, &bb);
  fp6_add(&r->c0, &bb, &aa);
  r->c1 = c1;
}
// src/fp12.rs:174-185
B200_NOINL void fp12_sqr(fp12 *r, const fp12 *a) {
  fp6 ab, c0c1, c0, t;
  fp6_mul(&ab, &a->c0, &a->c1);
  fp6_add(&c0c1, &a->c0, &a->c1);
  fp6_mul_by_nonresidue(&c0, &a->c1);
  fp6_add(&c0, &c0, &a->c0);
  fp6_mul(&c0, &c0, &c0c1);
  fp6_sub(&c0, &c0, &ab);
  fp6_mul_by_nonresidue(&t, &ab);
  fp6_sub(&r->c0, &c0, &t);
  fp6_add(&r->c1, &ab, &ab);
}
// src/fp12.rs:116-128   (r may alias f)
B200_NOINL void fp12_mul_by_014(fp12 *r, const fp12 *f, const fp2 *c0, const fp2 *c1, const fp2 *c4) {
  fp6 aa, bb, t;
  fp6_mul_by_01(&aa, &f->c0, c0, c1);
  fp6_mul_by_1(&bb, &f->c1, c4);
  fp2 o = fp2_add(*c1, *c4);
  fp6_add(&t, &f->c1, &f->c0);
  fp6_mul_by_01(&t, &t, c0, &o);
  fp6_sub(&t, &t, &aa);
  fp6_sub(&r->c1, &t, &bb);
  fp6_mul_by_nonresidue(&bb, &bb);
  fp6_add(&r->c0, &bb, &aa);
}
// src/fp12.rs:145-171   (r may alias a)
B200_NOINL void fp12_frobenius(fp12 *r, const fp12 *a) {
  fp6 c0, c1;
  fp6_frobenius(&c0, &a->c0);
  fp6_frobenius(&c1, &a->c1);
  fp2 k = fp2{fp_const(K_FROB12_C1_R), fp_const(K_FROB12_C1_U)};
  // c1 * Fp6::from(k): every Fp2 coefficient times k
  r->c0 = c0;
  r->c1.c0 = M2(c1.c0, k);
  r->c1.c1 = M2(c1.c1, k);
  r->c1.c2 = M2(c1.c2, k);
}
// src/fp12.rs:187-195
B200_NOINL void fp12_inv(fp12 *r, const fp12 *a) {
  fp6 s0, s1, t, nt;
  fp6_sqr(&s0, &a->c0);
  fp6_sqr(&s1, &a->c1);
  fp6_mul_by_nonresidue(&s1, &s1);
  fp6_sub(&s0, &s0, &s1);
  fp6_inv(&t, &s0);
  fp6_neg(&nt, &t);
  fp6 r0, r1;
  fp6_mul(&r0, &a->c0, &t);
  fp6_mul(&r1, &a->c1, &nt);
  r->c0 = r0;
  r->c1 = r1;
}

// memory layout: 12 Fp in the order c0.c0.c0, c0.c0.c1, c0.c1.c0, ... c1.c2.c1 (struct order)
B200_DEV void fp6_load(fp6 *r, const void *p) {
  const char *q = reinterpret_cast<const char *>(p);
  r->c0 = fp2_load(q);
  r->c1 = fp2_load(q + 96);
  r->c2 = fp2_load(q + 192);
}
B200_DEV void fp6_store(void *p, const fp6 *a) {
  char *q = reinterpret_cast<char *>(p);
  fp2_store(q, a->c0);
  fp2_store(q + 96, a->c1);
  fp2_store(q + 192, a->c2);
}
B200_DEV void fp12_load(fp12 *r, const void *p) {
  fp6_load(&r->c0, p);
  fp6_load(&r->c1, reinterpret_cast<const char *>(p) + 288);
}
B200_DEV void fp12_store(void *p, const fp12 *a) {
  fp6_store(p, &a->c0);
  fp6_store(reinterpret_cast<char *>(p) + 288, &a->c1);
}

}  // namespace b200
