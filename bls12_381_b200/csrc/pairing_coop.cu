// Lane-cooperative pairing kernels (pairing_variant = 7): six lanes per pairing, the Fp12 accumulator of the Miller loop and
// every temporary of the final exponentiation distributed one Fp2 coefficient per lane (coop12.cuh) — no Fp12 in local
// memory.  Replaces, with bit-identical results, src/pairings.rs miller_loop :668-694 + ell :696-707 (over G2Prepared
// coefficients :498-546), multi_miller_loop :554-603 and final_exponentiation :48-176.
//
// Pairs are handed to warps in a grid-stride loop (5 pairs per warp and turn); a block is BLOCK_WARPS warps, one block
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
B200_DEV fp2 co_ell(const cgrp &g, const fp2 &f, const char *co, const fp &pc) {
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

// f^|x| then conjugate (cycolotomic_exp, src/pairings.rs:115-132); the first multiplication (1 * f) is a copy
B200_DEV fp2 co_cyclotomic_exp(const cgrp &g, const fp2 &f) {
  fp2 r = f;
#pragma unroll 1
  for (int b = 62; b >= 0; b--) {
    r = co_cyclotomic_sqr(g, r);
    if ((B200_BLS_X >> b) & 1) r = co_mul(g, r, f);
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

constexpr int CO_FLAG_MILLER = 1, CO_FLAG_FINAL_EXP = 2;

// flags & 1: f = Miller loop of (P_i, prepared Q_i) else f = in[i];  flags & 2: f = final_exponentiation(f).
// One warp = 5 pairs; grid-stride over groups of 5.
__global__ void __launch_bounds__(512, 1) k_coop_pairing(int flags, const char *pxy, const uint8_t *pinf, const char *coeffs,
                                                        const uint8_t *qinf, const char *in, size_t n, char *out,
                                                        const uint32_t *pow2) {
  B200_DYN_SMEM(uint32_t, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  int grp = lane / CO_LANES;
  cgrp g;
  g.live = grp < CO_GROUPS;
  g.k = lane - CO_LANES * grp;
  if (!g.live) grp = 0;
  g.bd = smem + (size_t)(warp * CO_GROUPS + grp) * CO_BOARD;
  const size_t stride = (size_t)gridDim.x * nwarp * CO_GROUPS;
#pragma unroll 1
  for (size_t base = ((size_t)blockIdx.x * nwarp + warp) * CO_GROUPS; base < n; base += stride) {
    size_t i = base + grp;
    const bool valid = g.live && i < n;
    if (i >= n) i = n - 1;
    fp2 f;
    bool ident = false;
    if (flags & CO_FLAG_MILLER) {
      ident = (pinf != nullptr && pinf[i] != 0) || (qinf != nullptr && qinf[i] != 0);
      // lanes 0,1 of a group keep P.x, lanes 2,3 P.y (the others never use theirs)
      fp pc = fp_load_ro(pxy + 96 * i + ((g.k & 2) ? 48 : 0));
      f = co_miller_prepared(g, coeffs + (size_t)19584 * i, pc);
      if (ident) f = co_one(g);  // src/pairings.rs:566-569 (term skipped) / :636-651 (pairing of an identity)
    } else {
      f = co_load12(g, in + 576 * i);
    }
    if (flags & CO_FLAG_FINAL_EXP) f = co_final_exponentiation(g, f, pow2);
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
  if (warps > 16) warps = 16;
  size_t turns = (n + CO_GROUPS - 1) / CO_GROUPS;                      // warp-turns of 5 pairs
  unsigned grid = (unsigned)((turns + warps - 1) / warps);
  if (grid > (unsigned)ctx->sm_count) grid = (unsigned)ctx->sm_count;  // one block per SM, grid-stride beyond that
  size_t smem = (size_t)warps * CO_WARP_SMEM;
#ifndef B200_HOST_EMUL
  if (!ctx->coop_attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k_coop_pairing, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * CO_WARP_SMEM);
    if (e != cudaSuccess) return b200::set_err(ctx, e, "cudaFuncSetAttribute(k_coop_pairing)");
    ctx->coop_attr_done = true;
  }
#endif
  B200_LAUNCH_ON(ctx, strm, k_coop_pairing, grid, 32 * warps, smem, flags, (const char *)p, (const uint8_t *)pi,
                 (const char *)coeffs, (const uint8_t *)qi, (const char *)in, n, (char *)out, (const uint32_t *)ctx->inv_pow2);
  return B200_OK;
}
