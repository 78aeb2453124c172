// Batched-affine tree levels for the MSM buckets (included by capi_msm.cu).
//
// A bucket with m points is summed as a tree: level 1 adds the points pairwise (floor(m/2) independent additions),
// level 2 adds those sums pairwise, ...  All pair additions of a level — across all buckets of the window group —
// are independent, so every thread takes K consecutive pairs of the flattened pair list and adds them in AFFINE
// coordinates with Montgomery's simultaneous-inversion trick:
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1
// i.e. 1 (prefix product) + 2 (individual inverse, running inverse) + 3 (lambda, lambda^2, lambda*(x1-x3)) = 6
// multiplications per addition, plus ONE inversion per K additions done by the binary GCD of fp_inv.cuh on the ALU
// pipe.  The XYZZ mixed addition the bucket kernel otherwise uses costs 10.  After LV levels (3 for the 2^20 config:
// ~32 points per bucket -> ~4) the remaining entries go through the XYZZ bucket kernel.
// Exceptional pairs are classified before the inversion: P = Q -> tangent (denominator 2y), P = -Q -> the group
// identity (stored as the marker x = y = 0, which is not on the curve), identity operand -> copy of the other one.
// Not a reference algorithm (the reference has no MSM); the group element is identical, parity is on the result.
#pragma once
#include "curve.cuh"
#include "fp_inv.cuh"

namespace b200 {

constexpr int AFF_KMAX = 64;

B200_DEV fp f_inv_fast(const fp &a, const uint32_t *pow2) { return fp_inv_fast(a, pow2); }
B200_DEV fp2 f_inv_fast(const fp2 &a, const uint32_t *pow2) {  // src/fp2.rs:300-320 with the binary-GCD Fp inverse
  fp t = fp_inv_fast(fp_add(fp_mul_c(a.c0, a.c0), fp_mul_c(a.c1, a.c1)), pow2);
  return fp2{fp_mul_c(a.c0, t), fp_mul_c(a.c1, fp_neg(t))};
}

template <class F>
struct aff_pt {
  F x, y;
};
template <class F>
B200_DEV bool aff_is_inf(const aff_pt<F> &p) { return f_is_zero(p.x) && f_is_zero(p.y); }

// entry `idx` of a slot's list at the current level
template <class F, bool L0>
B200_DEV aff_pt<F> aff_load_entry(const char *points, const char *bx, const uint32_t *sorted_w, const char *inbuf_w,
                                  uint32_t pos) {
  constexpr size_t FB = field_traits<F>::bytes, AB = 2 * FB;
  if (L0) {
    uint32_t e = __ldg(sorted_w + pos);
    const char *pp = points + AB * (size_t)(e & 0x3fffffffu);
    const char *px = (e & 0x40000000u) ? bx + FB * (size_t)(e & 0x3fffffffu) : pp;
    aff_pt<F> r{field_traits<F>::load_ro(px), field_traits<F>::load_ro(pp + FB)};
    if (e >> 31) r.y = f_neg(r.y);
    return r;
  }
  const char *q = inbuf_w + AB * (size_t)pos;
  return aff_pt<F>{field_traits<F>::load(q), field_traits<F>::load(q + FB)};
}

// kind: 0 chord, 1 tangent, 2 result is the identity, 3 result = p, 4 result = q
template <class F>
B200_DEV int aff_classify(const aff_pt<F> &p, const aff_pt<F> &q, F &den) {
  bool ip = aff_is_inf(p), iq = aff_is_inf(q);
  den = field_traits<F>::one();
  if (ip && iq) return 2;
  if (iq) return 3;
  if (ip) return 4;
  F dx = f_sub(q.x, p.x);
  if (!f_is_zero(dx)) {
    den = dx;
    return 0;
  }
  if (f_eq(p.y, q.y) && !f_is_zero(p.y)) {
    den = f_dbl(p.y);
    return 1;
  }
  return 2;
}

// per-slot counts of the next level; giant slots (>= giant points at level 0) stay out of the affine path
__global__ void __launch_bounds__(256) k_aff_counts(size_t total, const uint32_t *cprev, const uint32_t *hist0, uint32_t giant,
                                                  uint32_t *npairs, uint32_t *cnew) {
  size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k >= total) return;
  uint32_t c = hist0[k] >= giant ? 0u : cprev[k];
  npairs[k] = c >> 1;
  cnew[k] = (c + 1) >> 1;
}

// grid = (ceil(max_pairs_per_window / (K * 128)), windows).  Arrays are per slot, offsets are window-relative.
template <class F, bool L0>
__global__ void __launch_bounds__(128) k_aff_level(int nbuckets, int K, const char *points, const char *bx, size_t sstride,
                                                 const uint32_t *in_off, const uint32_t *sorted, const char *inbuf,
                                                 size_t in_cap, const uint32_t *npairs, const uint32_t *pairoff,
                                                 const uint32_t *outoff, char *outbuf, size_t out_cap,
                                                 const uint32_t *pow2) {
  constexpr size_t FB = field_traits<F>::bytes, AB = 2 * FB;
  const int j = blockIdx.y;
  const uint32_t *np = npairs + (size_t)j * nbuckets, *po = pairoff + (size_t)j * nbuckets;
  const uint32_t *oo = outoff + (size_t)j * nbuckets, *io = in_off + (size_t)j * nbuckets;
  const uint32_t *sorted_w = L0 ? sorted + (size_t)j * sstride : nullptr;
  const char *inbuf_w = L0 ? nullptr : inbuf + AB * (size_t)j * in_cap;
  char *out_w = outbuf + AB * (size_t)j * out_cap;
  const uint32_t P = po[nbuckets - 1] + np[nbuckets - 1];
  const uint32_t p0 = (blockIdx.x * 128u + threadIdx.x) * (uint32_t)K;
  if (p0 >= P) return;
  const uint32_t p1 = min(p0 + (uint32_t)K, P);
  int lo = 0, hi = nbuckets;  // last bucket b with pairoff[b] <= p0
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (po[mid] <= p0) lo = mid; else hi = mid;
  }
  F pre[AFF_KMAX];
  uint32_t bk[AFF_KMAX];
  uint16_t ik[AFF_KMAX];  // pair index inside the bucket (< 512 since giants are excluded)
  // pass 1: prefix products of the denominators
  F acc = field_traits<F>::one();
  {
    uint32_t bb = (uint32_t)lo, i = p0 - po[lo];
    for (uint32_t p = p0; p < p1; p++) {
      while (i >= np[bb]) {
        bb++;
        i = 0;
      }
      uint32_t base = io[bb] + 2 * i;
      aff_pt<F> a = aff_load_entry<F, L0>(points, bx, sorted_w, inbuf_w, base);
      aff_pt<F> b = aff_load_entry<F, L0>(points, bx, sorted_w, inbuf_w, base + 1);
      F den;
      aff_classify(a, b, den);
      pre[p - p0] = acc;
      bk[p - p0] = bb;
      ik[p - p0] = (uint16_t)i;
      acc = f_mul(acc, den);
      i++;
    }
  }
  F inv = f_inv_fast(acc, pow2);
  // pass 2, backwards: individual inverses and the additions
  for (uint32_t p = p1; p-- > p0;) {
    uint32_t bb = bk[p - p0], i = ik[p - p0];
    uint32_t base = io[bb] + 2 * i;
    aff_pt<F> a = aff_load_entry<F, L0>(points, bx, sorted_w, inbuf_w, base);
    aff_pt<F> b = aff_load_entry<F, L0>(points, bx, sorted_w, inbuf_w, base + 1);
    F den;
    int kind = aff_classify(a, b, den);
    F dinv = f_mul(inv, pre[p - p0]);
    inv = f_mul(inv, den);
    aff_pt<F> r;
    if (kind <= 1) {
      F num;
      if (kind == 0) {
        num = f_sub(b.y, a.y);
      } else {
        F xx = f_sqr(a.x);
        num = f_add(f_dbl(xx), xx);
      }
      F lam = f_mul(num, dinv);
      r.x = f_sub(f_sub(f_sqr(lam), a.x), b.x);
      r.y = f_sub(f_mul(lam, f_sub(a.x, r.x)), a.y);
    } else if (kind == 2) {
      r.x = field_traits<F>::zero();
      r.y = field_traits<F>::zero();
    } else {
      r = kind == 3 ? a : b;
    }
    char *dst = out_w + AB * (size_t)(oo[bb] + i);
    f_store(dst, r.x);
    f_store(dst + FB, r.y);
  }
}

// odd entry of a slot passes through to the next level (one thread per slot)
template <class F, bool L0>
__global__ void __launch_bounds__(256) k_aff_leftover(int nbuckets, size_t total, const char *points, const char *bx,
                                                    size_t sstride, const uint32_t *cprev, const uint32_t *hist0,
                                                    uint32_t giant, const uint32_t *in_off, const uint32_t *sorted,
                                                    const char *inbuf, size_t in_cap, const uint32_t *npairs,
                                                    const uint32_t *outoff, char *outbuf, size_t out_cap) {
  constexpr size_t FB = field_traits<F>::bytes, AB = 2 * FB;
  size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k >= total) return;
  uint32_t c = cprev[k];
  if (hist0[k] >= giant || (c & 1u) == 0) return;
  size_t j = k / nbuckets;
  const uint32_t *sorted_w = L0 ? sorted + j * sstride : nullptr;
  const char *inbuf_w = L0 ? nullptr : inbuf + AB * j * in_cap;
  aff_pt<F> a = aff_load_entry<F, L0>(points, bx, sorted_w, inbuf_w, in_off[k] + c - 1);
  char *dst = outbuf + AB * (j * out_cap + outoff[k] + npairs[k]);
  f_store(dst, a.x);
  f_store(dst + FB, a.y);
}

// the bucket kernel over the entries that are left after the affine levels (affine points in `buf`, identity
// markers skipped); slots visited in `order`
template <class F, int MINB>
__global__ void __launch_bounds__(128, MINB) k_msm_accumulate_buf(int nbuckets, size_t total, size_t slot0, const char *buf,
                                                                size_t cap, const uint32_t *cnt_final,
                                                                const uint32_t *off_final, const uint32_t *hist0,
                                                                uint32_t giant, const uint32_t *order, char *buckets) {
  constexpr size_t FB = field_traits<F>::bytes, AB = 2 * FB, PB = 3 * FB;
  size_t t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t0 >= total) return;
  size_t k = slot0 + order[t0];
  if (hist0[k] >= giant) return;  // handled by the giant kernels on the original lists
  size_t j = (k - slot0) / nbuckets;  // window index inside the group: buf / cnt / off are group-relative
  size_t kr = k - slot0;
  uint32_t cnt = cnt_final[kr];
  const char *src = buf + AB * (j * cap + off_final[kr]);
  xyzz<F> acc = xyzz_identity<F>();
  for (uint32_t t = 0; t < cnt; t++) {
    F x = field_traits<F>::load(src + AB * t), y = field_traits<F>::load(src + AB * t + FB);
    if (f_is_zero(x) && f_is_zero(y)) continue;
    acc = xyzz_add_mixed(acc, x, y);
  }
  proj_store<F>(buckets + PB * k, xyzz_to_proj(acc));
}

}  // namespace b200
