// C ABI, part 2: batched Miller loop / final exponentiation / pairing (BASELINE config 4) and the
// product mode behind multi_miller_loop.  One thread per pair; pairs are independent, so batches shard
// by pair index with no collective.
#include "ctx.cuh"
#include "pairing.cuh"

using namespace b200;

namespace {

constexpr unsigned PAIR_BLOCK = 64;

__global__ void __launch_bounds__(PAIR_BLOCK) k_miller_loop(const char *pxy, const uint8_t *pinf, const char *qxy,
                                                          const uint8_t *qinf, size_t n, char *out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  fp12 f;
  miller_loop_pair(&f, affine_load<fp>(pxy, pinf, i), affine_load<fp2>(qxy, qinf, i));
  fp12_store(out + 576 * i, &f);
}
__global__ void __launch_bounds__(PAIR_BLOCK) k_final_exp(const char *in, size_t n, char *out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  fp12 f;
  fp12_load(&f, in + 576 * i);
  final_exponentiation(&f);
  fp12_store(out + 576 * i, &f);
}
// product of n Fp12 values: per-thread strided partial products, then a shared-memory tree
__global__ void __launch_bounds__(PAIR_BLOCK) k_fp12_product(const char *in, size_t n, char *out) {
  extern __shared__ char smem[];
  fp12 acc, t;
  fp12_set_one(&acc);
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
    fp12_load(&t, in + 576 * i);
    fp12_mul(&acc, &acc, &t);
  }
  fp12_store(smem + 576 * threadIdx.x, &acc);
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      fp12_load(&acc, smem + 576 * threadIdx.x);
      fp12_load(&t, smem + 576 * (threadIdx.x + s));
      fp12_mul(&acc, &acc, &t);
      fp12_store(smem + 576 * threadIdx.x, &acc);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    fp12_load(&acc, smem);
    fp12_store(out, &acc);
  }
}

inline unsigned nblk(size_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

int miller_dev(b200_ctx *ctx, const void *p, const void *pi, const void *q, const void *qi, size_t n, void *out) {
  if (n == 0) return B200_OK;
  B200_LAUNCH(ctx, k_miller_loop, nblk(n, PAIR_BLOCK), PAIR_BLOCK, 0, (const char *)p, (const uint8_t *)pi,
              (const char *)q, (const uint8_t *)qi, n, (char *)out);
  return B200_OK;
}
int final_exp_dev(b200_ctx *ctx, const void *in, size_t n, void *out) {
  if (n == 0) return B200_OK;
  B200_LAUNCH(ctx, k_final_exp, nblk(n, PAIR_BLOCK), PAIR_BLOCK, 0, (const char *)in, n, (char *)out);
  return B200_OK;
}
int product_dev(b200_ctx *ctx, const void *in, size_t n, void *out) {
  // two levels: up to 148*? blocks would need a second pass; n partial products are tiny, one block is enough
  B200_LAUNCH(ctx, k_fp12_product, 1, PAIR_BLOCK, 576 * PAIR_BLOCK, (const char *)in, n, (char *)out);
  return B200_OK;
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
  int rc = miller_dev(ctx, p, p_inf, q, q_inf, n, gt_out);
  if (rc == B200_OK) rc = final_exp_dev(ctx, gt_out, n, gt_out);
  return rc != B200_OK ? rc : sync(ctx);
}
int b200_fp12_product_dev(b200_ctx *ctx, const void *in, size_t n, void *out) {
  CHECK_CTX(ctx);
  if (!out || (n && !in)) return B200_EINVAL;
  int rc = product_dev(ctx, in, n, out);
  return rc != B200_OK ? rc : sync(ctx);
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
  int rc = miller_dev(ctx, h.dp, h.dpi, h.dq, h.dqi, n, dout);
  if (rc == B200_OK) rc = final_exp_dev(ctx, dout, n, dout);
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
  int rc = miller_dev(ctx, h.dp, h.dpi, h.dq, h.dqi, n, dml);
  if (rc == B200_OK) rc = product_dev(ctx, dml, n, dout);
  if (rc != B200_OK) return rc;
  B200_CUDA(ctx, cudaMemcpyAsync(out, dout, 576, cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}

}  // extern "C"
