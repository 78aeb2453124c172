// C ABI, part 2: batched Miller loop / final exponentiation / pairing (BASELINE config 4) and the
// product mode behind multi_miller_loop.  One thread per pair; pairs are independent, so batches shard
// by pair index with no collective.
#include "ctx.cuh"

using namespace b200;

#define DECL_VARIANT(v)                                                                                              \
  int b200_pair_miller_##v(b200_ctx *, cudaStream_t, const void *, const void *, const void *, const void *, size_t,   \
                           void *);                                                                                    \
  int b200_pair_final_exp_##v(b200_ctx *, cudaStream_t, const void *, size_t, void *);                                 \
  int b200_pair_product_##v(b200_ctx *, const void *, size_t, void *);                                                 \
  int b200_pair_g2_prepare_##v(b200_ctx *, cudaStream_t, const void *, const void *, size_t, void *);                  \
  int b200_pair_miller_prepared_##v(b200_ctx *, cudaStream_t, const void *, const void *, const void *, const void *,  \
                                    size_t, void *);
DECL_VARIANT(v4)
// pairing_coop.cu: flags 1 = Miller loop over prepared coefficients, 2 = final exponentiation (of `in` when bit 0 is clear)
int b200_pair_coop_launch(b200_ctx *, cudaStream_t, int flags, const void *p, const void *pi, const void *coeffs,
                          const void *qi, const void *in, size_t n, void *out);
int b200_pair_coop_prepare_launch(b200_ctx *, cudaStream_t, const void *q, const void *qi, size_t n, void *coeffs);
int b200_pair_coop_product_launch(b200_ctx *, cudaStream_t, int final_exp, const void *p, const void *pi, const void *coeffs,
                                  const void *qi, int terms, size_t n_terms, size_t n_items, void *out);
int b200_pair_coop_fold_launch(b200_ctx *, cudaStream_t, const void *in, size_t n, void *scratch, void *out);

namespace {
// G2 line coefficients for the six-lane kernels: small batches with six lanes per Q (latency), large ones one thread per Q
int g2_prepare_on(b200_ctx *ctx, cudaStream_t st, const void *q, const void *qi, size_t n, void *coeffs) {
  if (n <= (size_t)ctx->tune_coop_prepare_max) return b200_pair_coop_prepare_launch(ctx, st, q, qi, n, coeffs);
  return b200_pair_g2_prepare_v4(ctx, st, q, qi, n, coeffs);
}

// The pairing kernels live in their own translation unit (pairing_v4.cu: 255 registers, 4 resident 64-thread blocks
// per SM, Fp2 multiply = Karatsuba over fp_mul_c calls).  Lower register budgets (168 / 128) and the inlined /
// lazily reduced Fp2 multiplies were built the same way, measured slower, and removed.
// variant 7: G2 line coefficients of all pairs (68 x 288 B each) into the scratch arena, then the six-lane Miller loop
// (+ final exponentiation when `with_final_exp`) in ONE kernel
int coop_from_affine(b200_ctx *ctx, cudaStream_t st, const void *p, const void *pi, const void *q, const void *qi, size_t n,
                     void *out, bool with_final_exp) {
  if (n == 0) return B200_OK;
  int rc = arena_reserve(ctx, (size_t)19584 * n + 256);
  if (rc != B200_OK) return rc;
  char *co = arena_take<char>(ctx, (size_t)19584 * n);
  rc = g2_prepare_on(ctx, st, q, qi, n, co);
  if (rc != B200_OK) return rc;
  if (with_final_exp && ctx->tune_coop_split) {
    rc = b200_pair_coop_launch(ctx, st, 1, p, pi, co, qi, nullptr, n, out);
    return rc != B200_OK ? rc : b200_pair_coop_launch(ctx, st, 2, nullptr, nullptr, nullptr, nullptr, out, n, out);
  }
  return b200_pair_coop_launch(ctx, st, with_final_exp ? 3 : 1, p, pi, co, qi, nullptr, n, out);
}
int miller_on(b200_ctx *ctx, cudaStream_t st, const void *p, const void *pi, const void *q, const void *qi, size_t n,
              void *out) {
  if (ctx->coop_for(n)) return coop_from_affine(ctx, st, p, pi, q, qi, n, out, false);
  return b200_pair_miller_v4(ctx, st, p, pi, q, qi, n, out);
}
int final_exp_on(b200_ctx *ctx, cudaStream_t st, const void *in, size_t n, void *out) {
  if (ctx->coop_for(n)) return b200_pair_coop_launch(ctx, st, 2, nullptr, nullptr, nullptr, nullptr, in, n, out);
  return b200_pair_final_exp_v4(ctx, st, in, n, out);
}
int miller_dev(b200_ctx *ctx, const void *p, const void *pi, const void *q, const void *qi, size_t n, void *out) {
  return miller_on(ctx, ctx->stream, p, pi, q, qi, n, out);
}
int final_exp_dev(b200_ctx *ctx, const void *in, size_t n, void *out) { return final_exp_on(ctx, ctx->stream, in, n, out); }

// Full pairings of a batch: Miller loop then final exponentiation per CHUNK, chunks alternating between the two
// streams of the ctx.  Every thread of either kernel runs for milliseconds, so a single launch per kernel pays a
// whole extra wave for the last, partially filled one (2^16 pairs = 1.73 waves of 148 SMs x 4 x 64 threads);
// with independent chunks in flight the block scheduler back-fills the tail of one kernel with blocks of another.
// Six-lane kernels, batch larger than one wave of groups (148 SMs x 12 warps x 5 = 8 880 pairs): the batch is cut into
// `tune_coop_chunks` chunks that alternate between the two streams of the ctx.  Each kernel is one persistent block per SM and
// every warp-turn runs for milliseconds, so a single launch per phase idles the SMs whose warps have one turn less (2^16 pairs =
// 7.38 turns per warp: 91 of 148 blocks wait out the 8th turn, in both phases); with independent chunks in flight the blocks of
// the other stream's next kernel take over an SM the moment a block exits.
int coop_chunked(b200_ctx *ctx, const void *p, const void *pi, const void *q, const void *qi, size_t n, void *out, int chunks) {
  int rc = arena_reserve(ctx, (size_t)19584 * n + 256);
  if (rc != B200_OK) return rc;
  char *co = arena_take<char>(ctx, (size_t)19584 * n);
  B200_CUDA(ctx, cudaEventRecord(ctx->ev_sync[0], ctx->stream));
  B200_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_sync[0], 0));
  struct join_guard {   // every exit path (error returns included) re-joins the main stream with the side stream
    b200_ctx *c;
    ~join_guard() {
      cudaEventRecord(c->ev_sync[1], c->stream2);
      cudaStreamWaitEvent(c->stream, c->ev_sync[1], 0);
    }
  } join_on_exit{ctx};
  // every chunk but the first is a whole number of waves, so that the kernels that finish LAST end on full rounds; the first
  // chunk takes the remainder (its partial last round is back-filled by the blocks of the following chunk)
  const size_t wave = (size_t)ctx->sm_count * (ctx->tune_coop_warps < 1 ? 1 : ctx->tune_coop_warps) * 5;
  size_t per = (n / chunks) / wave * wave;
  if (per < wave) per = wave;
  const size_t first = n - (size_t)(chunks - 1) * per;   // n > 2 waves and chunks <= n / wave (caller) keep this positive
  for (int c = 0; c < chunks; c++) {
    size_t lo = c == 0 ? 0 : first + (size_t)(c - 1) * per, cnt = c == 0 ? first : per;
    cudaStream_t st = (c & 1) ? ctx->stream2 : ctx->stream;
    const char *cp = (const char *)p + 96 * lo, *cq = (const char *)q + 192 * lo;
    const uint8_t *cpi = pi ? (const uint8_t *)pi + lo : nullptr, *cqi = qi ? (const uint8_t *)qi + lo : nullptr;
    char *cout = (char *)out + 576 * lo, *cco = co + (size_t)19584 * lo;
    rc = b200_pair_g2_prepare_v4(ctx, st, cq, cqi, cnt, cco);
    if (rc == B200_OK) rc = b200_pair_coop_launch(ctx, st, 1, cp, cpi, cco, cqi, nullptr, cnt, cout);
    if (rc == B200_OK) rc = b200_pair_coop_launch(ctx, st, 2, nullptr, nullptr, nullptr, nullptr, cout, cnt, cout);
    if (rc != B200_OK) return rc;
  }
  return B200_OK;
}
int pairing_dev(b200_ctx *ctx, const void *p, const void *pi, const void *q, const void *qi, size_t n, void *out) {
  if (ctx->coop_for(n)) {
    const size_t wave = (size_t)ctx->sm_count * (ctx->tune_coop_warps < 1 ? 1 : ctx->tune_coop_warps) * 5;
    int chunks = ctx->tune_coop_chunks;
    if ((size_t)chunks > n / wave) chunks = (int)(n / wave);
    if (chunks > 1 && n > 2 * wave && ctx->tune_coop_split) return coop_chunked(ctx, p, pi, q, qi, n, out, chunks);
    return coop_from_affine(ctx, ctx->stream, p, pi, q, qi, n, out, true);
  }
  int chunks = ctx->tune_pairing_chunks;
  if (chunks < 1) chunks = 1;
  // chunking only pays when the batch exceeds one wave of resident threads (148 SMs x 4 blocks x 64 = 37 888): below
  // that a single launch per kernel is already one (partial) wave, and smaller launches only add latency
  if (n <= (size_t)ctx->sm_count * 4 * 64 + 2048) chunks = 1;
  if (chunks == 1) {
    int rc = b200_pair_miller_v4(ctx, ctx->stream, p, pi, q, qi, n, out);
    return rc != B200_OK ? rc : b200_pair_final_exp_v4(ctx, ctx->stream, out, n, out);
  }
  B200_CUDA(ctx, cudaEventRecord(ctx->ev_sync[0], ctx->stream));
  B200_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_sync[0], 0));
  struct join_guard {   // every exit path (error returns included) re-joins the main stream with the side stream
    b200_ctx *c;
    ~join_guard() {
      cudaEventRecord(c->ev_sync[1], c->stream2);
      cudaStreamWaitEvent(c->stream, c->ev_sync[1], 0);
    }
  } join_on_exit{ctx};
  size_t per = ((n + chunks - 1) / chunks + 63) & ~(size_t)63;
  for (int c = 0; c < chunks; c++) {
    size_t lo = (size_t)c * per, cnt = lo >= n ? 0 : (n - lo < per ? n - lo : per);
    if (cnt == 0) break;
    cudaStream_t st = (c & 1) ? ctx->stream2 : ctx->stream;
    const char *cp = (const char *)p + 96 * lo, *cq = (const char *)q + 192 * lo;
    const uint8_t *cpi = pi ? (const uint8_t *)pi + lo : nullptr, *cqi = qi ? (const uint8_t *)qi + lo : nullptr;
    char *co = (char *)out + 576 * lo;
    // the chunked two-stream schedule belongs to the one-thread-per-pairing kernels: call them directly (the dispatchers
    // would pick the six-lane kernels for a chunk-sized batch, and those use the ctx arena, which two streams cannot share)
    int rc = b200_pair_miller_v4(ctx, st, cp, cpi, cq, cqi, cnt, co);
    if (rc == B200_OK) rc = b200_pair_final_exp_v4(ctx, st, co, cnt, co);
    if (rc != B200_OK) return rc;
  }
  return B200_OK;
}
int product_dev(b200_ctx *ctx, const void *in, size_t n, void *out) {
  if (ctx->coop_products()) {     // multi-block fold, 16 values per group and level
    int rc = arena_reserve(ctx, 2 * 576 * (n / 16 + 4) + 256);
    if (rc != B200_OK) return rc;
    char *scratch = arena_take<char>(ctx, 2 * 576 * (n / 16 + 4));
    return b200_pair_coop_fold_launch(ctx, ctx->stream, in, n, scratch, out);
  }
  return b200_pair_product_v4(ctx, in, n, out);
}
// Products of pairings with ONE squaring of the accumulator per bit for all terms of a product (src/pairings.rs:554-603).
// n_products items of `terms` consecutive pairs each; out[i] = MillerLoopResult (final_exp = 0) or Gt (final_exp = 1).
int pairing_products_dev(b200_ctx *ctx, const void *p, const void *pi, const void *q, const void *qi, size_t terms,
                         size_t n_products, int final_exp, void *out) {
  size_t n = terms * n_products;
  if (n == 0) return B200_OK;
  int rc = arena_reserve(ctx, (size_t)19584 * n + 256);
  if (rc != B200_OK) return rc;
  char *co = arena_take<char>(ctx, (size_t)19584 * n);
  rc = g2_prepare_on(ctx, ctx->stream, q, qi, n, co);
  if (rc != B200_OK) return rc;
  return b200_pair_coop_product_launch(ctx, ctx->stream, final_exp, p, pi, co, qi, (int)terms, n, n_products, out);
}
// multi_miller_loop over n terms -> ONE value: the terms are cut into chunks of T (shared squaring inside a chunk, one group
// of six lanes per chunk; T grows with n so that one wave of groups covers the batch), the chunk values are folded by
// k_coop_product.  `coeffs` = prepared coefficients on the device, or nullptr (then q/qi are affine points to prepare).
int multi_miller_dev(b200_ctx *ctx, const void *p, const void *pi, const void *q, const void *qi, const void *coeffs, size_t n,
                     void *out) {
  size_t capacity = (size_t)ctx->sm_count * (ctx->tune_coop_warps < 1 ? 1 : ctx->tune_coop_warps) * 5;
  size_t T = (n + capacity - 1) / capacity;
  if (T < 1) T = 1;
  if (T > 64) T = 64;
  size_t items = n == 0 ? 0 : (n + T - 1) / T;
  size_t need = (coeffs ? 0 : (size_t)19584 * n) + 576 * (items + 1) + 2 * 576 * (items / 16 + 4) + 1024;
  int rc = arena_reserve(ctx, need);
  if (rc != B200_OK) return rc;
  const char *co = (const char *)coeffs;
  if (!co && n) {
    char *c2 = arena_take<char>(ctx, (size_t)19584 * n);
    rc = g2_prepare_on(ctx, ctx->stream, q, qi, n, c2);
    if (rc != B200_OK) return rc;
    co = c2;
  }
  char *partial = arena_take<char>(ctx, 576 * (items + 1));
  char *scratch = arena_take<char>(ctx, 2 * 576 * (items / 16 + 4));
  rc = b200_pair_coop_product_launch(ctx, ctx->stream, 0, p, pi, co, qi, (int)T, n, items, partial);
  if (rc != B200_OK) return rc;
  return b200_pair_coop_fold_launch(ctx, ctx->stream, partial, items, scratch, out);
}
int sync(b200_ctx *ctx) {
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
struct host_pairs {
  const void *dp = nullptr, *dpi = nullptr, *dq = nullptr, *dqi = nullptr;
  int rc = B200_OK;
  host_pairs(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *pi, const b200_g2_affine *q, const uint8_t *qi,
             size_t n, size_t extra) {
    rc = stage_reserve(ctx, 96 * n + 192 * n + 2 * n + extra + 16 * 256);
    if (rc != B200_OK) return;
    auto up = [&](const void *h, size_t bytes) -> const void * {
      if (!h) return nullptr;
      void *d = stage_take(ctx, bytes);
      cudaError_t e = cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream);
      if (e != cudaSuccess) rc = set_err(ctx, e, "H2D");
      return d;
    };
    dp = up(p, 96 * n);
    dpi = up(pi, n);
    dq = up(q, 192 * n);
    dqi = up(qi, n);
  }
};

}  // namespace

#define CHECK_CTX(ctx)                      \
  if ((ctx) == nullptr) return B200_EINVAL; \
  ctx_guard guard__(ctx);                   \
  if (!guard__.ok) return B200_ENODEV

extern "C" {

int b200_miller_loop_batch_dev(b200_ctx *ctx, const void *p, const void *p_inf, const void *q, const void *q_inf,
                               size_t n, void *out) {
  CHECK_CTX(ctx);
  if (n && (!p || !q || !out)) return B200_EINVAL;
  int rc = miller_dev(ctx, p, p_inf, q, q_inf, n, out);
  return rc != B200_OK ? rc : sync(ctx);
}
int b200_final_exponentiation_batch_dev(b200_ctx *ctx, const void *in, size_t n, void *out) {
  CHECK_CTX(ctx);
  if (n && (!in || !out)) return B200_EINVAL;
  int rc = final_exp_dev(ctx, in, n, out);
  return rc != B200_OK ? rc : sync(ctx);
}
int b200_pairing_batch_dev(b200_ctx *ctx, const void *p, const void *p_inf, const void *q, const void *q_inf, size_t n,
                           void *gt_out) {
  CHECK_CTX(ctx);
  if (n && (!p || !q || !gt_out)) return B200_EINVAL;
  int rc = pairing_dev(ctx, p, p_inf, q, q_inf, n, gt_out);
  return rc != B200_OK ? rc : sync(ctx);
}
int b200_fp12_product_dev(b200_ctx *ctx, const void *in, size_t n, void *out) {
  CHECK_CTX(ctx);
  if (!out || (n && !in)) return B200_EINVAL;
  int rc = product_dev(ctx, in, n, out);
  return rc != B200_OK ? rc : sync(ctx);
}

// ---- G2Prepared: prepared coefficients stay resident on the device for fixed verifying keys
int b200_g2_prepare_dev(b200_ctx *ctx, const void *q, const void *q_inf, size_t n, void *coeffs) {
  CHECK_CTX(ctx);
  if (n && (!q || !coeffs)) return B200_EINVAL;
  int rc = g2_prepare_on(ctx, ctx->stream, q, q_inf, n, coeffs);
  return rc != B200_OK ? rc : sync(ctx);
}
int b200_miller_loop_prepared_batch_dev(b200_ctx *ctx, const void *p, const void *p_inf, const void *coeffs,
                                        const void *q_inf, size_t n, void *out) {
  CHECK_CTX(ctx);
  if (n && (!p || !coeffs || !out)) return B200_EINVAL;
  int rc = ctx->coop_for(n) ? b200_pair_coop_launch(ctx, ctx->stream, 1, p, p_inf, coeffs, q_inf, nullptr, n, out)
                                          : b200_pair_miller_prepared_v4(ctx, ctx->stream, p, p_inf, coeffs, q_inf, n, out);
  return rc != B200_OK ? rc : sync(ctx);
}
int b200_g2_prepare(b200_ctx *ctx, const b200_g2_affine *q, const uint8_t *q_inf, size_t n, b200_fp2 *coeffs) {
  CHECK_CTX(ctx);
  if (n && (!q || !coeffs)) return B200_EINVAL;
  if (n == 0) return B200_OK;
  int rc = stage_reserve(ctx, 192 * n + n + 19584 * n + 1024);
  if (rc != B200_OK) return rc;
  void *dq = stage_take(ctx, 192 * n), *dqi = q_inf ? stage_take(ctx, n) : nullptr, *dc = stage_take(ctx, 19584 * n);
  B200_CUDA(ctx, cudaMemcpyAsync(dq, q, 192 * n, cudaMemcpyHostToDevice, ctx->stream));
  if (q_inf) B200_CUDA(ctx, cudaMemcpyAsync(dqi, q_inf, n, cudaMemcpyHostToDevice, ctx->stream));
  rc = g2_prepare_on(ctx, ctx->stream, dq, dqi, n, dc);
  if (rc != B200_OK) return rc;
  B200_CUDA(ctx, cudaMemcpyAsync(coeffs, dc, 19584 * n, cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}
// multi_miller_loop(&[(&p_i, &prepared_i)]) -> ONE MillerLoopResult  (src/pairings.rs:554-603)
int b200_multi_miller_loop_prepared(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *p_inf, const b200_fp2 *coeffs,
                                    const uint8_t *q_inf, size_t n, b200_fp12 *out) {
  CHECK_CTX(ctx);
  if (!out || (n && (!p || !coeffs))) return B200_EINVAL;
  int rc = stage_reserve(ctx, 96 * n + 2 * n + 19584 * n + 576 * (n + 2) + 2048);
  if (rc != B200_OK) return rc;
  void *dp = stage_take(ctx, 96 * (n ? n : 1)), *dpi = p_inf ? stage_take(ctx, n ? n : 1) : nullptr;
  void *dqi = q_inf ? stage_take(ctx, n ? n : 1) : nullptr, *dc = stage_take(ctx, 19584 * (n ? n : 1));
  void *dml = stage_take(ctx, 576 * (n ? n : 1)), *dout = stage_take(ctx, 576);
  if (n) {
    B200_CUDA(ctx, cudaMemcpyAsync(dp, p, 96 * n, cudaMemcpyHostToDevice, ctx->stream));
    if (p_inf) B200_CUDA(ctx, cudaMemcpyAsync(dpi, p_inf, n, cudaMemcpyHostToDevice, ctx->stream));
    if (q_inf) B200_CUDA(ctx, cudaMemcpyAsync(dqi, q_inf, n, cudaMemcpyHostToDevice, ctx->stream));
    B200_CUDA(ctx, cudaMemcpyAsync(dc, coeffs, 19584 * n, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (ctx->coop_products()) {
    rc = multi_miller_dev(ctx, dp, dpi, nullptr, dqi, dc, n, dout);
  } else {
    rc = b200_pair_miller_prepared_v4(ctx, ctx->stream, dp, dpi, dc, dqi, n, dml);
    if (rc == B200_OK) rc = product_dev(ctx, dml, n, dout);
  }
  if (rc != B200_OK) return rc;
  B200_CUDA(ctx, cudaMemcpyAsync(out, dout, 576, cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}

int b200_miller_loop_batch(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *p_inf, const b200_g2_affine *q,
                           const uint8_t *q_inf, size_t n, b200_fp12 *out) {
  CHECK_CTX(ctx);
  if (n && (!p || !q || !out)) return B200_EINVAL;
  if (n == 0) return B200_OK;
  host_pairs h(ctx, p, p_inf, q, q_inf, n, 576 * n);
  if (h.rc != B200_OK) return h.rc;
  void *dout = stage_take(ctx, 576 * n);
  int rc = miller_dev(ctx, h.dp, h.dpi, h.dq, h.dqi, n, dout);
  if (rc != B200_OK) return rc;
  B200_CUDA(ctx, cudaMemcpyAsync(out, dout, 576 * n, cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}
int b200_final_exponentiation_batch(b200_ctx *ctx, const b200_fp12 *in, size_t n, b200_fp12 *out) {
  CHECK_CTX(ctx);
  if (n && (!in || !out)) return B200_EINVAL;
  if (n == 0) return B200_OK;
  int rc = stage_reserve(ctx, 576 * n + 256);
  if (rc != B200_OK) return rc;
  void *d = stage_take(ctx, 576 * n);
  B200_CUDA(ctx, cudaMemcpyAsync(d, in, 576 * n, cudaMemcpyHostToDevice, ctx->stream));
  rc = final_exp_dev(ctx, d, n, d);
  if (rc != B200_OK) return rc;
  B200_CUDA(ctx, cudaMemcpyAsync(out, d, 576 * n, cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}
int b200_pairing_batch(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *p_inf, const b200_g2_affine *q,
                       const uint8_t *q_inf, size_t n, b200_fp12 *gt_out) {
  CHECK_CTX(ctx);
  if (n && (!p || !q || !gt_out)) return B200_EINVAL;
  if (n == 0) return B200_OK;
  host_pairs h(ctx, p, p_inf, q, q_inf, n, 576 * n);
  if (h.rc != B200_OK) return h.rc;
  void *dout = stage_take(ctx, 576 * n);
  int rc = pairing_dev(ctx, h.dp, h.dpi, h.dq, h.dqi, n, dout);
  if (rc != B200_OK) return rc;
  B200_CUDA(ctx, cudaMemcpyAsync(gt_out, dout, 576 * n, cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}
int b200_multi_miller_loop(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *p_inf, const b200_g2_affine *q,
                           const uint8_t *q_inf, size_t n, b200_fp12 *out) {
  CHECK_CTX(ctx);
  if (!out || (n && (!p || !q))) return B200_EINVAL;
  host_pairs h(ctx, p, p_inf, q, q_inf, n, 576 * n + 576 + 512);
  if (h.rc != B200_OK) return h.rc;
  void *dml = stage_take(ctx, 576 * (n ? n : 1));
  void *dout = stage_take(ctx, 576);
  int rc;
  if (ctx->coop_products()) {
    rc = multi_miller_dev(ctx, h.dp, h.dpi, h.dq, h.dqi, nullptr, n, dout);
  } else {
    rc = miller_dev(ctx, h.dp, h.dpi, h.dq, h.dqi, n, dml);
    if (rc == B200_OK) rc = product_dev(ctx, dml, n, dout);
  }
  if (rc != B200_OK) return rc;
  B200_CUDA(ctx, cudaMemcpyAsync(out, dout, 576, cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}
// ---- products of pairings (Groth16 / BLS batch verification shape): n_products x `terms` pairs -> n_products values
int b200_pairing_product_batch_dev(b200_ctx *ctx, const void *p, const void *p_inf, const void *q, const void *q_inf,
                                   size_t terms, size_t n_products, int final_exp, void *out) {
  CHECK_CTX(ctx);
  if (terms == 0 || (n_products && (!p || !q || !out))) return B200_EINVAL;
  int rc = pairing_products_dev(ctx, p, p_inf, q, q_inf, terms, n_products, final_exp, out);
  return rc != B200_OK ? rc : sync(ctx);
}
int b200_multi_miller_loop_dev(b200_ctx *ctx, const void *p, const void *p_inf, const void *q, const void *q_inf, size_t n,
                               void *out) {
  CHECK_CTX(ctx);
  if (!out || (n && (!p || !q))) return B200_EINVAL;
  int rc = multi_miller_dev(ctx, p, p_inf, q, q_inf, nullptr, n, out);
  return rc != B200_OK ? rc : sync(ctx);
}
int b200_pairing_product_batch(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *p_inf, const b200_g2_affine *q,
                               const uint8_t *q_inf, size_t terms, size_t n_products, int final_exp, b200_fp12 *out) {
  CHECK_CTX(ctx);
  if (terms == 0 || (n_products && (!p || !q || !out))) return B200_EINVAL;
  size_t n = terms * n_products;
  if (n == 0) return B200_OK;
  host_pairs h(ctx, p, p_inf, q, q_inf, n, 576 * n_products + 512);
  if (h.rc != B200_OK) return h.rc;
  void *dout = stage_take(ctx, 576 * n_products);
  int rc = pairing_products_dev(ctx, h.dp, h.dpi, h.dq, h.dqi, terms, n_products, final_exp, dout);
  if (rc != B200_OK) return rc;
  B200_CUDA(ctx, cudaMemcpyAsync(out, dout, 576 * n_products, cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}

}  // extern "C"
