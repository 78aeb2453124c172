// Lane-cooperative pairing kernels (pairing_variant = 7): six lanes per pairing, the Fp12 accumulator of the Miller loop and
// every temporary of the final exponentiation distributed one Fp2 coefficient per lane (coop12.cuh) — no Fp12 in local
// memory.  Replaces, with bit-identical results, src/pairings.rs miller_loop :668-694 + ell :696-707 (over G2Prepared
// coefficients :498-546), multi_miller_loop :554-603 and final_exponentiation :48-176.
//
// Pairs are handed to warps in a grid-stride loop (5 pairs per warp and turn; the groups of a warp synchronise independently); a block is `coop_warps` warps, one block
// per SM.  The G2 line coefficients come from memory (prepared by k_g2_prepare, 68 x 288 B per Q, in the order the loop
// consumes them); the lanes 0..3 of a group multiply them by P.x / P.y, all six then apply the sparse line to f.
#define B200_FP2_KCALL 1  // (the few plain Fp2 products in the inversion: Karatsuba over fp_mul_c)
#include "ctx.cuh"
#include "coop12.cuh"
#include "pairing.cuh"

using namespace b200;

namespace {

constexpr int S_C1 = 18, S_C4 = 19;  // board slots of the scaled line coefficients (coeffs.b * P.x, coeffs.a * P.y)

// f <- f * line(coeffs, P)   (ell, src/pairings.rs:696-707: f.mul_by_014(c.2, c.1 * p.x, c.0 * p.y))
// co = one coefficient triple (a, b, c) of 3 x 96 bytes; pc = P.x on lanes 0,1 and P.y on lanes 2,3 of the group
B200_NOINL fp2 co_ell(cgrp g, fp2 f, const char *co, fp pc) {
  const int k = g.k;
  const int off = k == 0 ? 96 : (k == 1 ? 144 : (k == 3 ? 48 : 0));  // b.c0, b.c1, a.c0, a.c1
  fp m = fp_mul_c(fp_load(co + off), pc);
  if (k < 4) co_put_half(g, k < 2 ? S_C1 : S_C4, k & 1, m);
  return co_mul_sparse(g, f, reinterpret_cast<const uint32_t *>(co + 192), co_slot(g, S_C1), co_slot(g, S_C4));
}

// Miller loop of ONE term over prepared coefficients (src/pairings.rs:554-603 with one term / miller_loop :668-694)
B200_DEV fp2 co_miller_prepared(const cgrp &g, const char *coeffs, const fp &pc) {
  fp2 f = co_one(g);
  const unsigned long long x = B200_BLS_X >> 1;
  int idx = 0;
#pragma unroll 1
  for (int b = 61; b >= 0; b--) {  // bit 62 is the leading one of x >> 1
    f = co_ell(g, f, coeffs + 288 * idx++, pc);
    if ((x >> b) & 1) f = co_ell(g, f, coeffs + 288 * idx++, pc);
    f = co_sqr(g, f);
  }
  f = co_ell(g, f, coeffs + 288 * idx, pc);
  return co_conj(g, f);  // BLS_X_IS_NEGATIVE
}

// multi_miller_loop over T prepared terms with ONE squaring of f per bit for all terms (src/pairings.rs:554-603):
//   f <- f^2 * prod_t line_t   per bit.   Terms with an identity on either side are skipped (:566-569, :578-581).
// coeffs / pxy / pinf / qinf are indexed by the absolute term number first + t.
B200_DEV fp2 co_miller_prepared_multi(const cgrp &g, const char *coeffs, const char *pxy, const uint8_t *pinf,
                                      const uint8_t *qinf, size_t first, int T) {
  fp2 f = co_one(g);
  const unsigned long long x = B200_BLS_X >> 1;
  const int poff = (g.k & 2) ? 48 : 0;
  int idx = 0;
#pragma unroll 1
  for (int b = 61; b >= -1; b--) {       // b = -1: the final doubling step after the loop
    const bool bit = b >= 0 && ((x >> b) & 1);
#pragma unroll 1
    for (int t = 0; t < T; t++) {
      const size_t i = first + t;
      if ((pinf != nullptr && pinf[i] != 0) || (qinf != nullptr && qinf[i] != 0)) continue;   // uniform over the group
      const char *co = coeffs + (size_t)19584 * i + 288 * idx;
      fp pc = fp_load_ro(pxy + 96 * i + poff);
      f = co_ell(g, f, co, pc);
      if (bit) f = co_ell(g, f, co + 288, pc);
    }
    idx += bit ? 2 : 1;
    if (b >= 0) f = co_sqr(g, f);
  }
  return co_conj(g, f);
}

// f^|x| then conjugate (cycolotomic_exp, src/pairings.rs:115-132); the first multiplication (1 * f) is a copy.  The squarings
// between two set bits of x run as ONE call (x = 0xd201000000010000: runs of 1, 2, 3, 9, 32 and 16).
B200_NOINL fp2 co_cyclotomic_exp(cgrp g, fp2 f) {
  fp2 r = f;
  int b = 62;
#pragma unroll 1
  while (b >= 0) {
    int n = 0;
    bool bit;
    do {
      n++;
      bit = (B200_BLS_X >> b) & 1;
      b--;
    } while (!bit && b >= 0);
    r = co_cyclotomic_sqr_n(g, r, n);
    if (bit) r = co_mul(g, r, f);
  }
  return co_conj(g, r);
}

// src/pairings.rs:134-176, same sequence of operations (frobenius_map^6 = conjugation)
B200_DEV fp2 co_final_exponentiation(const cgrp &g, const fp2 &f, const uint32_t *pow2) {
  fp2 t0 = co_conj(g, f);
  fp2 t1 = co_inv(g, f, pow2);
  fp2 t2 = co_mul(g, t0, t1);
  t1 = t2;
  t2 = co_frobenius(g, t2, 2);
  t2 = co_mul(g, t2, t1);
  t1 = co_conj(g, co_cyclotomic_sqr(g, t2));
  fp2 t3 = co_cyclotomic_exp(g, t2);
  fp2 t4 = co_cyclotomic_sqr(g, t3);
  fp2 t5 = co_mul(g, t1, t3);
  t1 = co_cyclotomic_exp(g, t5);
  t0 = co_cyclotomic_exp(g, t1);
  fp2 t6 = co_cyclotomic_exp(g, t0);
  t6 = co_mul(g, t6, t4);
  t4 = co_cyclotomic_exp(g, t6);
  t5 = co_conj(g, t5);
  t4 = co_mul(g, t4, co_mul(g, t5, t2));
  t5 = co_conj(g, t2);
  t1 = co_frobenius(g, co_mul(g, t1, t2), 3);
  t6 = co_frobenius(g, co_mul(g, t6, t5), 1);
  t3 = co_frobenius(g, co_mul(g, t3, t0), 2);
  t3 = co_mul(g, t3, t1);
  t3 = co_mul(g, t3, t6);
  return co_mul(g, t3, t4);
}

// ---- G2Prepared with six lanes per Q (src/pairings.rs:498-546, doubling_step :709-738, addition_step :740-770).  One thread
// per Q (k_g2_prepare of pairing_v4.cu) walks 68 dependent steps of ~11 Fp2 products each: 2.3 ms however small the batch.
// Here the independent Fp2 products of a formula level run in different lanes (doubling: levels of 4, 6 and 1 products;
// addition: 2, 2, 3, 3, 4), the lanes exchange them through the group's board, and every lane keeps the running point R and
// repeats the cheap linear steps.  Same formulas on the same values -> the same coefficients, limb for limb.
// lane-indexed choice among six candidates
B200_DEV fp2 co_pick6(int k, const fp2 &c0, const fp2 &c1, const fp2 &c2, const fp2 &c3, const fp2 &c4, const fp2 &c5) {
  fp2 r = c0;
  r = fp2_select(r, c1, k == 1);
  r = fp2_select(r, c2, k == 2);
  r = fp2_select(r, c3, k == 3);
  r = fp2_select(r, c4, k == 4);
  r = fp2_select(r, c5, k == 5);
  return r;
}
// one level: lane k publishes a * b in slot k; co_lvl_done() after the products have been read
B200_DEV void co_lvl(const cgrp &g, const fp2 &a, const fp2 &b) {
  co_put(g, g.k, M2(a, b));
  co_sync(g);
}
B200_DEV fp2 co_lvl_get(const cgrp &g, int slot) { return co_ld2(g.bd + CO_SLOT * slot); }
B200_DEV void co_lvl_done(const cgrp &g) { co_sync(g); }

struct co_r2 {
  fp2 x, y, z;
};
B200_DEV void co_store_line(const cgrp &g, char *dst, const fp2 &a, const fp2 &b, const fp2 &c) {
  if (g.k == 0) fp2_store(dst, a);
  if (g.k == 1) fp2_store(dst + 96, b);
  if (g.k == 2) fp2_store(dst + 192, c);
}
B200_NOINL co_r2 co_doubling_step(cgrp g, co_r2 r, char *dst) {
  const int k = g.k;
  // level 1: x^2, y^2, z^2, (z + y)^2
  fp2 zy = fp2_add(r.z, r.y);
  fp2 a = co_pick6(k, r.x, r.y, r.z, zy, r.x, r.x);
  co_lvl(g, a, a);
  fp2 tmp0 = co_lvl_get(g, 0), tmp1 = co_lvl_get(g, 1), zsq = co_lvl_get(g, 2), p3 = co_lvl_get(g, 3);
  co_lvl_done(g);
  fp2 rz = fp2_sub(fp2_sub(p3, tmp1), zsq);
  fp2 tmp4 = fp2_add(fp2_dbl(tmp0), tmp0);
  fp2 tmp6 = fp2_add(r.x, tmp4);
  // level 2: tmp1^2, (tmp1 + x)^2, tmp4^2, tmp4 * zsq, tmp6^2, rz * zsq
  fp2 t1x = fp2_add(tmp1, r.x);
  a = co_pick6(k, tmp1, t1x, tmp4, tmp4, tmp6, rz);
  fp2 b = co_pick6(k, tmp1, t1x, tmp4, zsq, tmp6, zsq);
  co_lvl(g, a, b);
  fp2 tmp2 = co_lvl_get(g, 0), q1 = co_lvl_get(g, 1), tmp5 = co_lvl_get(g, 2);
  fp2 tmp3 = fp2_dbl(fp2_sub(fp2_sub(q1, tmp0), tmp2));
  fp2 rx = fp2_sub(fp2_sub(tmp5, tmp3), tmp3);
  {
    fp2 q3 = co_lvl_get(g, 3), q4 = co_lvl_get(g, 4), q5 = co_lvl_get(g, 5);
    fp2 t1q = fp2_dbl(fp2_dbl(tmp1));
    co_store_line(g, dst, fp2_dbl(q5), fp2_neg(fp2_dbl(q3)), fp2_sub(fp2_sub(fp2_sub(q4, tmp0), tmp5), t1q));
  }
  co_lvl_done(g);
  // level 3: (tmp3 - rx) * tmp4
  co_lvl(g, fp2_sub(tmp3, rx), tmp4);
  fp2 ry = fp2_sub(co_lvl_get(g, 0), fp2_dbl(fp2_dbl(fp2_dbl(tmp2))));
  co_lvl_done(g);
  return co_r2{rx, ry, rz};
}
B200_NOINL co_r2 co_addition_step(cgrp g, co_r2 r, fp2 qx, fp2 qy, fp2 ysq, char *dst) {
  const int k = g.k;
  // level 1: z^2, (qy + z)^2
  fp2 qz = fp2_add(qy, r.z);
  fp2 a = fp2_select(r.z, qz, k == 1);
  co_lvl(g, a, a);
  fp2 zsq = co_lvl_get(g, 0), p1 = co_lvl_get(g, 1);
  co_lvl_done(g);
  // level 2: zsq * qx, ((qy + z)^2 - ysq - zsq) * zsq
  fp2 u = fp2_sub(fp2_sub(p1, ysq), zsq);
  co_lvl(g, fp2_select(zsq, u, k == 1), fp2_select(qx, zsq, k == 1));
  fp2 t0 = co_lvl_get(g, 0), t1 = co_lvl_get(g, 1);
  co_lvl_done(g);
  fp2 t2 = fp2_sub(t0, r.x);
  fp2 t6 = fp2_sub(fp2_sub(t1, r.y), r.y);
  // level 3: t2^2, t6 * qx, t6^2
  co_lvl(g, co_pick6(k, t2, t6, t6, t2, t2, t2), co_pick6(k, t2, qx, t6, t2, t2, t2));
  fp2 t3 = co_lvl_get(g, 0), t9 = co_lvl_get(g, 1), t6sq = co_lvl_get(g, 2);
  co_lvl_done(g);
  fp2 t4 = fp2_dbl(fp2_dbl(t3));
  // level 4: t4 * t2, t4 * x, (z + t2)^2
  fp2 zt = fp2_add(r.z, t2);
  co_lvl(g, co_pick6(k, t4, t4, zt, t4, t4, t4), co_pick6(k, t2, r.x, zt, t2, t2, t2));
  fp2 t5 = co_lvl_get(g, 0), t7 = co_lvl_get(g, 1), p2 = co_lvl_get(g, 2);
  co_lvl_done(g);
  fp2 rx = fp2_sub(fp2_sub(fp2_sub(t6sq, t5), t7), t7);
  fp2 rz = fp2_sub(fp2_sub(p2, zsq), t3);
  fp2 t10 = fp2_add(qy, rz);
  // level 5: (t7 - rx) * t6, y * t5, t10^2, rz^2
  fp2 d = fp2_sub(t7, rx);
  co_lvl(g, co_pick6(k, d, r.y, t10, rz, d, d), co_pick6(k, t6, t5, t10, rz, t6, t6));
  fp2 t8 = co_lvl_get(g, 0), yt5 = co_lvl_get(g, 1), t10sq = co_lvl_get(g, 2), rzsq = co_lvl_get(g, 3);
  co_lvl_done(g);
  fp2 ry = fp2_sub(t8, fp2_dbl(yt5));
  t10 = fp2_sub(fp2_sub(t10sq, ysq), rzsq);
  co_store_line(g, dst, fp2_dbl(rz), fp2_dbl(fp2_neg(t6)), fp2_sub(fp2_dbl(t9), t10));
  return co_r2{rx, ry, rz};
}
// the 68 coefficient triples of one Q in the order the Miller loop consumes them (g2_prepare of pairing.cuh)
B200_DEV void co_g2_prepare(const cgrp &g, const fp2 &qx, const fp2 &qy, char *coeffs) {
  co_r2 r{qx, qy, fp2_one()};
  co_lvl(g, qy, qy);
  fp2 ysq = co_lvl_get(g, 0);
  co_lvl_done(g);
  const unsigned long long x = B200_BLS_X >> 1;
  int idx = 0;
#pragma unroll 1
  for (int b = 61; b >= -1; b--) {   // bit 62 is the leading one of x >> 1; b = -1: the doubling step after the loop
    r = co_doubling_step(g, r, coeffs + 288 * idx++);
    if (b >= 0 && ((x >> b) & 1)) r = co_addition_step(g, r, qx, qy, ysq, coeffs + 288 * idx++);
  }
}
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_coop_g2_prepare(const char *qxy, const uint8_t *qinf, size_t n, char *coeffs) {
  B200_DYN_SMEM(uint32_t, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int grp = lane / CO_LANES;
  if (grp >= CO_GROUPS) return;   // lanes 30, 31: no group
  cgrp g;
  g.k = lane - CO_LANES * grp;
  g.mask = 0x3fu << (CO_LANES * grp);
  g.bd = smem + (size_t)(warp * CO_GROUPS + grp) * CO_BOARD;
  const size_t stride = (size_t)gridDim.x * nwarp * CO_GROUPS;
#pragma unroll 1
  for (size_t base = ((size_t)blockIdx.x * nwarp + warp) * CO_GROUPS; base < n; base += stride) {
    const size_t i = base + grp;
    if (i >= n) continue;
    affine<fp2> q = affine_load<fp2>(qxy, qinf, i);
    // the identity is prepared as the generator's coefficients (src/pairings.rs:528-544; the caller keeps the flag)
    fp2 qx = q.inf ? fp2{fp_const(K_G2_GEN_X0), fp_const(K_G2_GEN_X1)} : q.x;
    fp2 qy = q.inf ? fp2{fp_const(K_G2_GEN_Y0), fp_const(K_G2_GEN_Y1)} : q.y;
    co_g2_prepare(g, qx, qy, coeffs + (size_t)19584 * i);
  }
}

constexpr int CO_FLAG_MILLER = 1, CO_FLAG_FINAL_EXP = 2;

// Product mode: item i = the product over the `terms` consecutive pairs [i * terms, (i + 1) * terms) — ONE Miller value
// (and, with final_exp, one Gt) per item.  Groth16 / BLS batch verification: terms = 3..4, n_items = number of proofs;
// a single large product: n_items = number of term chunks, the partial products are multiplied by k_coop_product.
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_coop_pairing_product(int final_exp, const char *pxy, const uint8_t *pinf,
                                                                 const char *coeffs, const uint8_t *qinf, int terms,
                                                                 size_t n_terms, size_t n_items, char *out,
                                                                 const uint32_t *pow2) {
  B200_DYN_SMEM(uint32_t, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int grp = lane / CO_LANES;
  if (grp >= CO_GROUPS) return;   // lanes 30, 31: no group
  cgrp g;
  g.k = lane - CO_LANES * grp;
  g.mask = 0x3fu << (CO_LANES * grp);
  g.bd = smem + (size_t)(warp * CO_GROUPS + grp) * CO_BOARD;
  const size_t stride = (size_t)gridDim.x * nwarp * CO_GROUPS;
#pragma unroll 1
  for (size_t base = ((size_t)blockIdx.x * nwarp + warp) * CO_GROUPS; base < n_items; base += stride) {
    const size_t i = base + grp;
    if (i >= n_items) continue;
    const bool valid = true;
    size_t first = i * (size_t)terms;
    int T = first + terms <= n_terms ? terms : (int)(n_terms - first);   // the last chunk of a large product may be short
    fp2 f = co_miller_prepared_multi(g, coeffs, pxy, pinf, qinf, first, T);
    if (final_exp) f = co_final_exponentiation(g, f, pow2);
    if (valid) co_store12(g, out + 576 * i, f);
  }
}

// out[b] = product of in[b * per .. (b + 1) * per) (clipped to n), one item per GROUP; used to fold the partial products of
// a large multi_miller_loop: launch with n_out = ceil(n / per) groups, repeat until one value is left.
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_coop_product(const char *in, size_t n, int per, size_t n_out, char *out) {
  B200_DYN_SMEM(uint32_t, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int grp = lane / CO_LANES;
  if (grp >= CO_GROUPS) return;   // lanes 30, 31: no group
  cgrp g;
  g.k = lane - CO_LANES * grp;
  g.mask = 0x3fu << (CO_LANES * grp);
  g.bd = smem + (size_t)(warp * CO_GROUPS + grp) * CO_BOARD;
  const size_t stride = (size_t)gridDim.x * nwarp * CO_GROUPS;
#pragma unroll 1
  for (size_t base = ((size_t)blockIdx.x * nwarp + warp) * CO_GROUPS; base < n_out; base += stride) {
    const size_t i = base + grp;
    if (i >= n_out) continue;
    const bool valid = true;
    fp2 f = co_one(g);
#pragma unroll 1
    for (int t = 0; t < per; t++) {
      size_t j = i * (size_t)per + t;
      if (j >= n) break;
      f = co_mul(g, f, co_load12(g, in + 576 * j));
    }
    if (valid) co_store12(g, out + 576 * i, f);
  }
}

// flags & 1: f = Miller loop of (P_i, prepared Q_i) else f = in[i];  flags & 2: f = final_exponentiation(f).
// One warp = 5 pairs; grid-stride over groups of 5.
template <int MAXT, int FLAGS>
__global__ void __launch_bounds__(MAXT, 1) k_coop_pairing(int, const char *pxy, const uint8_t *pinf, const char *coeffs,
                                                        const uint8_t *qinf, const char *in, size_t n, char *out,
                                                        const uint32_t *pow2) {
  B200_DYN_SMEM(uint32_t, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int grp = lane / CO_LANES;
  if (grp >= CO_GROUPS) return;   // lanes 30, 31: no group
  cgrp g;
  g.k = lane - CO_LANES * grp;
  g.mask = 0x3fu << (CO_LANES * grp);
  g.bd = smem + (size_t)(warp * CO_GROUPS + grp) * CO_BOARD;
  const size_t stride = (size_t)gridDim.x * nwarp * CO_GROUPS;
#pragma unroll 1
  for (size_t base = ((size_t)blockIdx.x * nwarp + warp) * CO_GROUPS; base < n; base += stride) {
    const size_t i = base + grp;
    if (i >= n) continue;   // groups are independent: nothing to do for this one
    const bool valid = true;
    fp2 f;
    bool ident = false;
    if (FLAGS & CO_FLAG_MILLER) {
      ident = (pinf != nullptr && pinf[i] != 0) || (qinf != nullptr && qinf[i] != 0);
      // lanes 0,1 of a group keep P.x, lanes 2,3 P.y (the others never use theirs)
      fp pc = fp_load_ro(pxy + 96 * i + ((g.k & 2) ? 48 : 0));
      f = co_miller_prepared(g, coeffs + (size_t)19584 * i, pc);
      if (ident) f = co_one(g);  // src/pairings.rs:566-569 (term skipped) / :636-651 (pairing of an identity)
    } else {
      f = co_load12(g, in + 576 * i);
    }
    if (FLAGS & CO_FLAG_FINAL_EXP) f = co_final_exponentiation(g, f, pow2);
    if (valid) co_store12(g, out + 576 * i, f);
  }
}

}  // namespace

// host side: called by capi_pairing.cu when pairing_variant == 7
int b200_pair_coop_launch(b200_ctx *ctx, cudaStream_t strm, int flags, const void *p, const void *pi, const void *coeffs,
                          const void *qi, const void *in, size_t n, void *out) {
  if (n == 0) return B200_OK;
  int warps = ctx->tune_coop_warps;
  if (warps < 1) warps = 1;
  if (warps > 12) warps = 12;
  size_t turns = (n + CO_GROUPS - 1) / CO_GROUPS;                      // warp-turns of 5 pairs
  // a batch of less than one wave is spread over ALL SMs first (as few warps per scheduler as possible): a lone warp runs its
  // dependent chain ~2x faster than three that share a multiplier, and round 2's first version packed 12 warps into the first
  // ceil(turns / 12) SMs (1024 pairs: 18 of 148 SMs busy)
  {
    size_t per_sm = (turns + (size_t)ctx->sm_count - 1) / (size_t)ctx->sm_count;
    if (per_sm < 1) per_sm = 1;
    if (per_sm < (size_t)warps) warps = (int)per_sm;
  }
  unsigned grid = (unsigned)((turns + warps - 1) / warps);
  if (grid > (unsigned)ctx->sm_count) grid = (unsigned)ctx->sm_count;  // one block per SM, grid-stride beyond that
  size_t smem = (size_t)warps * CO_WARP_SMEM;
  // the kernel is compiled per phase (1 = Miller loop, 2 = final exponentiation, 3 = both in one launch) and per register
  // budget (255 registers for <= 8 warps per SM, 168 for <= 12): the instruction cache, not the multiplier, limits these
  // kernels (ncu: no_instruction is the second stall), and a phase-specific kernel keeps every warp of an SM in the same
  // few functions
#ifdef B200_HOST_EMUL
#define CO_SET_ATTR(MAXT, FL)
#else
#define CO_SET_ATTR(MAXT, FL)                                                                                                \
  if (!ctx->coop_attr_done[(MAXT == 256 ? 0 : 3) + FL - 1]) {                                                                \
    cudaError_t e = cudaFuncSetAttribute(k_coop_pairing<MAXT, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                                         (MAXT / 32) * CO_WARP_SMEM);                                                        \
    if (e != cudaSuccess) return b200::set_err(ctx, e, "cudaFuncSetAttribute(k_coop_pairing)");                              \
    ctx->coop_attr_done[(MAXT == 256 ? 0 : 3) + FL - 1] = true;                                                              \
  }
#endif
#define CO_LAUNCH(MAXT, FL)                                                                                                  \
  do {                                                                                                                       \
    CO_SET_ATTR(MAXT, FL)                                                                                                    \
    B200_LAUNCH_ON(ctx, strm, (k_coop_pairing<MAXT, FL>), grid, 32 * warps, smem, flags, (const char *)p,                    \
                   (const uint8_t *)pi, (const char *)coeffs, (const uint8_t *)qi, (const char *)in, n, (char *)out,        \
                   (const uint32_t *)ctx->inv_pow2);                                                                         \
  } while (0)
  if (warps > 12) warps = 12;
  smem = (size_t)warps * CO_WARP_SMEM;
  if (warps <= 8) {
    if (flags == 1) CO_LAUNCH(256, 1); else if (flags == 2) CO_LAUNCH(256, 2); else CO_LAUNCH(256, 3);
  } else {
    if (flags == 1) CO_LAUNCH(384, 1); else if (flags == 2) CO_LAUNCH(384, 2); else CO_LAUNCH(384, 3);
  }
#undef CO_LAUNCH
#undef CO_SET_ATTR
  return B200_OK;
}

// G2 line coefficients with six lanes per Q (small batches)
int b200_pair_coop_prepare_launch(b200_ctx *ctx, cudaStream_t strm, const void *q, const void *qi, size_t n, void *coeffs) {
  if (n == 0) return B200_OK;
  int warps = ctx->tune_coop_warps;
  if (warps < 1) warps = 1;
  if (warps > 8) warps = 8;
  size_t turns = (n + CO_GROUPS - 1) / CO_GROUPS;
  if (turns < (size_t)warps * ctx->sm_count) warps = (int)((turns + ctx->sm_count - 1) / ctx->sm_count);   // spread over the SMs first
  if (warps < 1) warps = 1;
  unsigned grid = (unsigned)((turns + warps - 1) / warps);
  if (grid > (unsigned)ctx->sm_count) grid = (unsigned)ctx->sm_count;
  size_t smem = (size_t)warps * CO_WARP_SMEM;
#ifndef B200_HOST_EMUL
  if (!ctx->coop_attr_done[7]) {
    cudaError_t e = cudaFuncSetAttribute(k_coop_g2_prepare<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * CO_WARP_SMEM);
    if (e != cudaSuccess) return b200::set_err(ctx, e, "cudaFuncSetAttribute(k_coop_g2_prepare)");
    ctx->coop_attr_done[7] = true;
  }
#endif
  B200_LAUNCH_ON(ctx, strm, k_coop_g2_prepare<256>, grid, 32 * warps, smem, (const char *)q, (const uint8_t *)qi, n, (char *)coeffs);
  return B200_OK;
}

namespace {
// launch geometry shared by the product kernels: `items` groups of work, `warps` warps per block, one block per SM
inline void coop_geometry(b200_ctx *ctx, size_t items, int *warps, unsigned *grid, size_t *smem) {
  int w = ctx->tune_coop_warps;
  if (w < 1) w = 1;
  if (w > 12) w = 12;
  size_t turns = (items + CO_GROUPS - 1) / CO_GROUPS;
  {                                                      // small jobs: spread over all SMs, as few warps per SM as possible
    size_t per_sm = (turns + (size_t)ctx->sm_count - 1) / (size_t)ctx->sm_count;
    if (per_sm < 1) per_sm = 1;
    if (per_sm < (size_t)w) w = (int)per_sm;
  }
  unsigned g = (unsigned)((turns + w - 1) / w);
  if (g > (unsigned)ctx->sm_count) g = (unsigned)ctx->sm_count;
  if (g < 1) g = 1;
  *warps = w;
  *grid = g;
  *smem = (size_t)w * CO_WARP_SMEM;
}
}  // namespace

// one Miller value (final_exp = 0) or Gt (final_exp = 1) per item = product over `terms` consecutive prepared pairs
int b200_pair_coop_product_launch(b200_ctx *ctx, cudaStream_t strm, int final_exp, const void *p, const void *pi,
                                  const void *coeffs, const void *qi, int terms, size_t n_terms, size_t n_items, void *out) {
  if (n_items == 0) return B200_OK;
  int warps;
  unsigned grid;
  size_t smem;
  coop_geometry(ctx, n_items, &warps, &grid, &smem);
#ifndef B200_HOST_EMUL
  if (!ctx->coop_attr_done[6]) {
    cudaError_t e = cudaFuncSetAttribute(k_coop_pairing_product<384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * CO_WARP_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_coop_product<384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * CO_WARP_SMEM);
    if (e != cudaSuccess) return b200::set_err(ctx, e, "cudaFuncSetAttribute(k_coop_pairing_product)");
    ctx->coop_attr_done[6] = true;
  }
#endif
  B200_LAUNCH_ON(ctx, strm, k_coop_pairing_product<384>, grid, 32 * warps, smem, final_exp, (const char *)p, (const uint8_t *)pi,
                 (const char *)coeffs, (const uint8_t *)qi, terms, n_terms, n_items, (char *)out, (const uint32_t *)ctx->inv_pow2);
  return B200_OK;
}
// out[0] = product of in[0..n) (n = 0: one()); `scratch` holds ceil(n / 16) Fp12 values; in is not modified
int b200_pair_coop_fold_launch(b200_ctx *ctx, cudaStream_t strm, const void *in, size_t n, void *scratch, void *out) {
  constexpr int PER = 16;
  const char *src = (const char *)in;
  char *bufs[2] = {(char *)scratch, (char *)scratch + 576 * ((n + PER - 1) / PER + 1)};
  int which = 0;
  for (;;) {
    size_t n_out = n <= (size_t)PER ? 1 : (n + PER - 1) / PER;
    char *dst = n_out == 1 ? (char *)out : bufs[which];
    int warps;
    unsigned grid;
    size_t smem;
    coop_geometry(ctx, n_out, &warps, &grid, &smem);
#ifndef B200_HOST_EMUL
    if (!ctx->coop_attr_done[6]) {
      cudaError_t e = cudaFuncSetAttribute(k_coop_pairing_product<384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * CO_WARP_SMEM);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k_coop_product<384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * CO_WARP_SMEM);
      if (e != cudaSuccess) return b200::set_err(ctx, e, "cudaFuncSetAttribute(k_coop_product)");
      ctx->coop_attr_done[6] = true;
    }
#endif
    B200_LAUNCH_ON(ctx, strm, k_coop_product<384>, grid, 32 * warps, smem, src, n, PER, n_out, dst);
    if (n_out == 1) return B200_OK;
    src = dst;
    n = n_out;
    which ^= 1;
  }
}
