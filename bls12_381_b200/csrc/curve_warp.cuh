// Lane-parallel group operations for the latency-bound serial tails (the Horner chain of an MSM).
//
// A single thread needs ~8 (double) / ~12 (add) field multiplications one after the other, each a
// ~300-IMAD dependent stream that cannot use more than one SMSP's multiplier.  Here ONE WARP performs one
// group operation: the point is replicated in every lane, the independent multiplications of a formula
// level are computed by different lanes in the same SIMT instruction stream (a warp-wide fp_mul costs the
// same whether 1 or 32 lanes are active), the products are broadcast with shuffles, and the cheap linear
// steps are recomputed redundantly by all lanes.  RCB doubling (src/g1.rs:638-667) has 3 dependent levels
// (4, 2, 2 products) instead of 8 sequential multiplications; the complete addition (src/g1.rs:670-712)
// has 2 levels (6, 6) instead of 12.  Same formulas, same field elements as curve.cuh — only the schedule
// differs, so results stay bit-identical.
#pragma once
#include "curve.cuh"

namespace b200 {

B200_DEV fp f_shfl(const fp &a, int src) {
  fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src);
  return r;
}
B200_DEV fp2 f_shfl(const fp2 &a, int src) { return fp2{f_shfl(a.c0, src), f_shfl(a.c1, src)}; }

// lane-indexed choice among up to 6 candidates (lanes >= count reuse candidate 0: their product is ignored)
template <class F>
B200_DEV F lane_pick(int lane, const F &c0, const F &c1, const F &c2, const F &c3, const F &c4, const F &c5) {
  F r = c0;
  r = f_select(r, c1, lane == 1);
  r = f_select(r, c2, lane == 2);
  r = f_select(r, c3, lane == 3);
  r = f_select(r, c4, lane == 4);
  r = f_select(r, c5, lane == 5);
  return r;
}

// GROUP form: the six lanes base .. base+5 of a warp perform the operation (sub = lane - base in 0..5), so a warp carries
// five independent operations at once (lanes 30/31 run along on garbage).  All 32 lanes must call; the lanes of a group hold
// the same `s` and return the same result.  warp_double / warp_add below are the one-group-per-warp special case.
template <class F>
B200_DEV proj<F> grp_double(const proj<F> &s, int lane, int base) {
  typedef field_traits<F> T;
  // level 1: y^2, y*z, z^2, x*y
  F a = lane_pick<F>(lane, s.y, s.y, s.z, s.x, s.y, s.y);
  F b = lane_pick<F>(lane, s.y, s.z, s.z, s.y, s.y, s.y);
  F p = f_mul(a, b);
  F t0 = f_shfl(p, base), t1 = f_shfl(p, base + 1), t2 = f_shfl(p, base + 2), xy = f_shfl(p, base + 3);
  F z3 = f_dbl(f_dbl(f_dbl(t0)));
  t2 = T::mul_by_3b(t2);
  F y3 = f_add(t0, t2);
  // level 2: t2*z3, t1*z3
  a = lane_pick<F>(lane, t2, t1, t2, t2, t2, t2);
  p = f_mul(a, z3);
  F x3 = f_shfl(p, base);
  z3 = f_shfl(p, base + 1);
  t1 = f_dbl(t2);
  t2 = f_add(t1, t2);
  t0 = f_sub(t0, t2);
  // level 3: t0*y3, t0*xy
  b = lane_pick<F>(lane, y3, xy, y3, y3, y3, y3);
  p = f_mul(t0, b);
  y3 = f_add(x3, f_shfl(p, base));
  x3 = f_dbl(f_shfl(p, base + 1));
  proj<F> r{x3, y3, z3};
  return proj_select(r, proj_identity<F>(), proj_is_identity(s));
}

template <class F>
B200_DEV proj<F> warp_double(const proj<F> &s, int lane) { return grp_double(s, lane, 0); }

template <class F>
B200_DEV proj<F> grp_add(const proj<F> &s, const proj<F> &r, int lane, int base) {
  typedef field_traits<F> T;
  // level 1: x1x2, y1y2, z1z2, (x1+y1)(x2+y2), (y1+z1)(y2+z2), (x1+z1)(x2+z2)
  F a = lane_pick<F>(lane, s.x, s.y, s.z, f_add(s.x, s.y), f_add(s.y, s.z), f_add(s.x, s.z));
  F b = lane_pick<F>(lane, r.x, r.y, r.z, f_add(r.x, r.y), f_add(r.y, r.z), f_add(r.x, r.z));
  F p = f_mul(a, b);
  F t0 = f_shfl(p, base), t1 = f_shfl(p, base + 1), t2 = f_shfl(p, base + 2), t3 = f_shfl(p, base + 3), t4 = f_shfl(p, base + 4), x3 = f_shfl(p, base + 5);
  t3 = f_sub(t3, f_add(t0, t1));
  t4 = f_sub(t4, f_add(t1, t2));
  F y3 = f_sub(x3, f_add(t0, t2));
  x3 = f_dbl(t0);
  t0 = f_add(x3, t0);
  t2 = T::mul_by_3b(t2);
  F z3 = f_add(t1, t2);
  t1 = f_sub(t1, t2);
  y3 = T::mul_by_3b(y3);
  // level 2: t4*y3, t3*t1, y3*t0, t1*z3, t0*t3, z3*t4
  a = lane_pick<F>(lane, t4, t3, y3, t1, t0, z3);
  b = lane_pick<F>(lane, y3, t1, t0, z3, t3, t4);
  p = f_mul(a, b);
  x3 = f_sub(f_shfl(p, base + 1), f_shfl(p, base));
  y3 = f_add(f_shfl(p, base + 3), f_shfl(p, base + 2));
  z3 = f_add(f_shfl(p, base + 5), f_shfl(p, base + 4));
  return proj<F>{x3, y3, z3};
}

template <class F>
B200_DEV proj<F> warp_add(const proj<F> &s, const proj<F> &r, int lane) { return grp_add(s, r, lane, 0); }

}  // namespace b200
