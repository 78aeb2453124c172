// C ABI, part 7: Gt * Scalar for a batch (src/pairings.rs:296-323).  Own translation unit with the pairing units' Fp2
// multiply (Karatsuba over fp_mul_c calls) — the kernel has the pairing kernels' shape: one thread per element, Fp12 in
// local memory, latency-bound.
#define B200_FP2_KCALL 1
#include "ctx.cuh"
#include "gt.cuh"

using namespace b200;

#define CHECK_CTX(ctx)                      \
  if ((ctx) == nullptr) return B200_EINVAL; \
  ctx_guard guard__(ctx);                   \
  if (!guard__.ok) return B200_ENODEV

extern "C" {

int b200_gt_mul_batch_dev(b200_ctx *ctx, const void *g, const void *scalars, size_t n, void *out) {
  CHECK_CTX(ctx);
  if (n && (!g || !scalars || !out)) return B200_EINVAL;
  if (n == 0) return B200_OK;
  B200_LAUNCH(ctx, k_gt_mul_batch, (unsigned)((n + 63) / 64), 64, 0, (const char *)g, (const uint32_t *)scalars, n, (char *)out);
  return B200_OK;
}
int b200_gt_mul_batch(b200_ctx *ctx, const b200_fp12 *g, const b200_scalar *scalars, size_t n, b200_fp12 *out) {
  CHECK_CTX(ctx);
  if (n && (!g || !scalars || !out)) return B200_EINVAL;
  if (n == 0) return B200_OK;
  int rc = stage_reserve(ctx, 2 * (576 * n + 256) + 32 * n + 256);
  if (rc != B200_OK) return rc;
  char *dg = (char *)stage_take(ctx, 576 * n), *dout = (char *)stage_take(ctx, 576 * n);
  uint32_t *ds = (uint32_t *)stage_take(ctx, 32 * n);
  B200_CUDA(ctx, cudaMemcpyAsync(dg, g, 576 * n, cudaMemcpyHostToDevice, ctx->stream));
  B200_CUDA(ctx, cudaMemcpyAsync(ds, scalars, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
  B200_LAUNCH(ctx, k_gt_mul_batch, (unsigned)((n + 63) / 64), 64, 0, (const char *)dg, (const uint32_t *)ds, n, dout);
  B200_CUDA(ctx, cudaMemcpyAsync(out, dout, 576 * n, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

}  // extern "C"
