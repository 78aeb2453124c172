// Short-Weierstrass a=0 curve arithmetic over F = fp (G1) or fp2 (G2), homogeneous projective
// coordinates with the COMPLETE Renes–Costello–Batina formulas — the same formulas, in the same
// operation order, as the reference (src/g1.rs: double :638-667, add :670-712, add_mixed :715-752,
// multiply :754-774; src/g2.rs: :709-738, :741-783, :786-823, :825-845), so raw (x,y,z) limbs are
// bit-identical to the reference's G1Projective/G2Projective, not just the affine value.
#pragma once
#include "fp2.cuh"

namespace b200 {

template <class F>
struct affine {
  F x, y;
  bool inf;
};
template <class F>
struct proj {
  F x, y, z;
};

template <class F>
B200_DEV proj<F> proj_identity() {  // src/g1.rs:605-611
  return proj<F>{field_traits<F>::zero(), field_traits<F>::one(), field_traits<F>::zero()};
}
template <class F>
B200_DEV affine<F> affine_identity() {  // src/g1.rs:187-193
  return affine<F>{field_traits<F>::zero(), field_traits<F>::one(), true};
}
template <class F>
B200_DEV bool proj_is_identity(const proj<F> &p) { return f_is_zero(p.z); }
template <class F>
B200_DEV proj<F> proj_from_affine(const affine<F> &a) {  // src/g1.rs:463-471
  return proj<F>{a.x, a.y, a.inf ? field_traits<F>::zero() : field_traits<F>::one()};
}
template <class F>
B200_DEV proj<F> proj_select(const proj<F> &a, const proj<F> &b, bool choose_b) {
  return proj<F>{f_select(a.x, b.x, choose_b), f_select(a.y, b.y, choose_b), f_select(a.z, b.z, choose_b)};
}
template <class F>
B200_DEV proj<F> proj_neg(const proj<F> &a) { return proj<F>{a.x, f_neg(a.y), a.z}; }

// RCB Alg. 9
template <class F>
B200_DEV proj<F> proj_double(const proj<F> &s) {
  typedef field_traits<F> T;
  F t0 = f_sqr(s.y);
  F z3 = f_dbl(t0);
  z3 = f_dbl(z3);
  z3 = f_dbl(z3);
  F t1 = f_mul(s.y, s.z);
  F t2 = f_sqr(s.z);
  t2 = T::mul_by_3b(t2);
  F x3 = f_mul(t2, z3);
  F y3 = f_add(t0, t2);
  z3 = f_mul(t1, z3);
  t1 = f_dbl(t2);
  t2 = f_add(t1, t2);
  t0 = f_sub(t0, t2);
  y3 = f_mul(t0, y3);
  y3 = f_add(x3, y3);
  t1 = f_mul(s.x, s.y);
  x3 = f_mul(t0, t1);
  x3 = f_dbl(x3);
  proj<F> r{x3, y3, z3};
  return proj_select(r, proj_identity<F>(), proj_is_identity(s));
}

// RCB Alg. 7
template <class F>
B200_DEV proj<F> proj_add(const proj<F> &s, const proj<F> &r) {
  typedef field_traits<F> T;
  F t0 = f_mul(s.x, r.x);
  F t1 = f_mul(s.y, r.y);
  F t2 = f_mul(s.z, r.z);
  F t3 = f_add(s.x, s.y);
  F t4 = f_add(r.x, r.y);
  t3 = f_mul(t3, t4);
  t4 = f_add(t0, t1);
  t3 = f_sub(t3, t4);
  t4 = f_add(s.y, s.z);
  F x3 = f_add(r.y, r.z);
  t4 = f_mul(t4, x3);
  x3 = f_add(t1, t2);
  t4 = f_sub(t4, x3);
  x3 = f_add(s.x, s.z);
  F y3 = f_add(r.x, r.z);
  x3 = f_mul(x3, y3);
  y3 = f_add(t0, t2);
  y3 = f_sub(x3, y3);
  x3 = f_dbl(t0);
  t0 = f_add(x3, t0);
  t2 = T::mul_by_3b(t2);
  F z3 = f_add(t1, t2);
  t1 = f_sub(t1, t2);
  y3 = T::mul_by_3b(y3);
  x3 = f_mul(t4, y3);
  t2 = f_mul(t3, t1);
  x3 = f_sub(t2, x3);
  y3 = f_mul(y3, t0);
  t1 = f_mul(t1, z3);
  y3 = f_add(t1, y3);
  t0 = f_mul(t0, t3);
  z3 = f_mul(z3, t4);
  z3 = f_add(z3, t0);
  return proj<F>{x3, y3, z3};
}

// RCB Alg. 8 (rhs affine, NOT the identity: callers handle rhs.inf, see src/g1.rs:751)
template <class F>
B200_DEV proj<F> proj_add_mixed_nz(const proj<F> &s, const F &rx, const F &ry) {
  typedef field_traits<F> T;
  F t0 = f_mul(s.x, rx);
  F t1 = f_mul(s.y, ry);
  F t3 = f_add(rx, ry);
  F t4 = f_add(s.x, s.y);
  t3 = f_mul(t3, t4);
  t4 = f_add(t0, t1);
  t3 = f_sub(t3, t4);
  t4 = f_mul(ry, s.z);
  t4 = f_add(t4, s.y);
  F y3 = f_mul(rx, s.z);
  y3 = f_add(y3, s.x);
  F x3 = f_dbl(t0);
  t0 = f_add(x3, t0);
  F t2 = T::mul_by_3b(s.z);
  F z3 = f_add(t1, t2);
  t1 = f_sub(t1, t2);
  y3 = T::mul_by_3b(y3);
  x3 = f_mul(t4, y3);
  t2 = f_mul(t3, t1);
  x3 = f_sub(t2, x3);
  y3 = f_mul(y3, t0);
  t1 = f_mul(t1, z3);
  y3 = f_add(t1, y3);
  t0 = f_mul(t0, t3);
  z3 = f_mul(z3, t4);
  z3 = f_add(z3, t0);
  return proj<F>{x3, y3, z3};
}
template <class F>
B200_DEV proj<F> proj_add_mixed(const proj<F> &s, const affine<F> &r) {
  proj<F> t = proj_add_mixed_nz(s, r.x, r.y);
  return proj_select(t, s, r.inf);
}

// G::multiply: 255 iterations of double + selected full add, MSB(bit 254) -> LSB, bit 255 skipped.
// `by` = canonical little-endian scalar as 8 x u32.
template <class F>
B200_DEV proj<F> proj_multiply(const proj<F> &s, const uint32_t by[8]) {
  proj<F> acc = proj_identity<F>();
#pragma unroll 1
  for (int bit = 254; bit >= 0; bit--) {
    acc = proj_double(acc);
    proj<F> sum = proj_add(acc, s);
    bool b = (by[bit >> 5] >> (bit & 31)) & 1;
    acc = proj_select(acc, sum, b);
  }
  return acc;
}

// ---- XYZZ accumulator (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity: ZZ = 0) for the MSM buckets.
// Not a reference type: the reference only has the complete projective formulas above (11 FpM + ~17 field
// additions per mixed add).  Inside a bucket only the GROUP ELEMENT matters, so buckets use
// the cheaper madd-2008-s (8M + 2S, 6 field additions) and fall back to explicit handling of the three
// exceptional cases the incomplete formula has — accumulator = identity, P + P, P + (-P) — which the parity
// suite exercises (duplicate points, P and -P with equal scalars, all-equal scalars).
template <class F>
struct xyzz {
  F x, y, zz, zzz;
};
template <class F>
B200_DEV xyzz<F> xyzz_identity() {
  return xyzz<F>{field_traits<F>::zero(), field_traits<F>::one(), field_traits<F>::zero(), field_traits<F>::zero()};
}
// acc + (px, py), (px, py) affine and not the identity
template <class F>
B200_DEV xyzz<F> xyzz_add_mixed(const xyzz<F> &a, const F &px, const F &py) {
  if (f_is_zero(a.zz)) return xyzz<F>{px, py, field_traits<F>::one(), field_traits<F>::one()};
  F u2 = f_mul(px, a.zz);
  F s2 = f_mul(py, a.zzz);
  F p = f_sub(u2, a.x);
  F r = f_sub(s2, a.y);
  if (f_is_zero(p)) {
    if (!f_is_zero(r)) return xyzz_identity<F>();  // P + (-P)
    // P + P: double the affine point (mdbl-2008-s-1 with Z = 1); y = 0 cannot happen on a prime-order subgroup
    // but is still the identity if it does
    if (f_is_zero(py)) return xyzz_identity<F>();
    F u = f_dbl(py);
    F v = f_sqr(u);
    F w = f_mul(u, v);
    F s = f_mul(px, v);
    F xx = f_sqr(px);
    F m = f_add(f_dbl(xx), xx);
    F x3 = f_sub(f_sqr(m), f_dbl(s));
    F y3 = f_sub(f_mul(m, f_sub(s, x3)), f_mul(w, py));
    return xyzz<F>{x3, y3, v, w};
  }
  F pp = f_sqr(p);
  F ppp = f_mul(p, pp);
  F q = f_mul(a.x, pp);
  F x3 = f_sub(f_sub(f_sqr(r), ppp), f_dbl(q));
  F y3 = f_sub(f_mul(r, f_sub(q, x3)), f_mul(a.y, ppp));
  return xyzz<F>{x3, y3, f_mul(a.zz, pp), f_mul(a.zzz, ppp)};
}
// same group element in the reference's homogeneous projective form: (X*ZZZ : Y*ZZ : ZZ*ZZZ)
template <class F>
B200_DEV proj<F> xyzz_to_proj(const xyzz<F> &a) {
  if (f_is_zero(a.zz)) return proj_identity<F>();
  return proj<F>{f_mul(a.x, a.zzz), f_mul(a.y, a.zz), f_mul(a.zz, a.zzz)};
}

// ---- memory layouts (include/bls12381_b200.h): affine = x||y, projective = x||y||z, limbs as Fp
template <class F>
B200_DEV proj<F> proj_load(const void *p) {
  const char *q = reinterpret_cast<const char *>(p);
  constexpr int B = field_traits<F>::bytes;
  return proj<F>{field_traits<F>::load(q), field_traits<F>::load(q + B), field_traits<F>::load(q + 2 * B)};
}
template <class F>
B200_DEV void proj_store(void *p, const proj<F> &a) {
  char *q = reinterpret_cast<char *>(p);
  constexpr int B = field_traits<F>::bytes;
  f_store(q, a.x);
  f_store(q + B, a.y);
  f_store(q + 2 * B, a.z);
}
// flag set => G*Affine::identity() whatever the coordinate bytes hold (marshalling rule of the C ABI)
template <class F>
B200_DEV affine<F> affine_load(const void *xy, const uint8_t *inf, size_t i) {
  constexpr int B = field_traits<F>::bytes;
  const char *q = reinterpret_cast<const char *>(xy) + (size_t)2 * B * i;
  bool is_inf = inf != nullptr && inf[i] != 0;
  if (is_inf) return affine_identity<F>();
  return affine<F>{field_traits<F>::load_ro(q), field_traits<F>::load_ro(q + B), false};
}
template <class F>
B200_DEV void affine_store(void *xy, uint8_t *inf, size_t i, const affine<F> &a) {
  constexpr int B = field_traits<F>::bytes;
  char *q = reinterpret_cast<char *>(xy) + (size_t)2 * B * i;
  f_store(q, a.x);
  f_store(q + B, a.y);
  if (inf) inf[i] = a.inf ? 1 : 0;
}
// src/g1.rs:49-63
template <class F>
B200_DEV affine<F> proj_to_affine(const proj<F> &p) {
  F zinv = f_inv(p.z);
  bool is_inf = f_is_zero(zinv);
  affine<F> a{f_mul(p.x, zinv), f_mul(p.y, zinv), false};
  affine<F> id = affine_identity<F>();
  return affine<F>{f_select(a.x, id.x, is_inf), f_select(a.y, id.y, is_inf), is_inf};
}

}  // namespace b200
