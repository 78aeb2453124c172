// C ABI, part 6: batched hash to curve.  Device functions, kernels and the launch plan live in
// h2c.cuh (shared with the CPU test harness); this file stages the buffers.
#include <vector>

#include "ctx.cuh"
#include "h2c.cuh"

using namespace b200;

namespace {

struct gpu_launcher {
  b200_ctx *ctx;
  template <class K, class... A>
  int operator()(K k, unsigned grid, unsigned block, A... a) {
    int tr = timing_begin(ctx, "k_h2c");
    B200_KERNEL_LAUNCH(k, grid, block, 0, ctx->stream, a...);
    timing_end(ctx, tr);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(ctx, e, "launch h2c kernel");
    return 0;
  }
};

// offsets must be ascending; returns the byte range [lo, hi) they cover
bool offsets_ok(const uint64_t *off, size_t n, uint64_t *lo, uint64_t *hi) {
  for (size_t i = 0; i < n; i++)
    if (off[i + 1] < off[i]) return false;
  *lo = off[0];
  *hi = off[n];
  return true;
}

// copies the messages, the rebased offsets and DST_prime to the device staging area
int stage_inputs(b200_ctx *ctx, const uint8_t *msgs, const uint64_t *off, size_t n, const uint8_t *dst, size_t dst_len,
                 size_t extra_bytes, uint8_t **d_msgs, uint64_t **d_off, uint8_t **d_dp, int *dp_len, uint8_t **d_extra,
                 std::vector<uint64_t> &rebased) {
  uint64_t lo, hi;
  if (!offsets_ok(off, n, &lo, &hi)) return B200_EINVAL;
  const size_t mbytes = (size_t)(hi - lo);
  if (mbytes && !msgs) return B200_EINVAL;
  uint8_t dp[256];
  *dp_len = h2c_dst_prime(dst, dst_len, dp);
  rebased.resize(n + 1);
  for (size_t i = 0; i <= n; i++) rebased[i] = off[i] - lo;
  int rc = stage_reserve(ctx, mbytes + 8 * (n + 1) + 256 + extra_bytes + 8 * 256);
  if (rc != B200_OK) return rc;
  *d_msgs = (uint8_t *)stage_take(ctx, mbytes ? mbytes : 1);
  *d_off = (uint64_t *)stage_take(ctx, 8 * (n + 1));
  *d_dp = (uint8_t *)stage_take(ctx, 256);
  *d_extra = (uint8_t *)stage_take(ctx, extra_bytes ? extra_bytes : 1);
  if (mbytes) B200_CUDA(ctx, cudaMemcpyAsync(*d_msgs, msgs + lo, mbytes, cudaMemcpyHostToDevice, ctx->stream));
  B200_CUDA(ctx, cudaMemcpyAsync(*d_off, rebased.data(), 8 * (n + 1), cudaMemcpyHostToDevice, ctx->stream));
  B200_CUDA(ctx, cudaMemcpyAsync(*d_dp, dp, 256, cudaMemcpyHostToDevice, ctx->stream));
  // dp / rebased are pageable host memory: the async copies above have consumed them when they return
  return B200_OK;
}

int hash_host(b200_ctx *ctx, int group, const uint8_t *msgs, const uint64_t *off, size_t n, const uint8_t *dst, size_t dst_len,
              int encode, void *out) {
  const int count = encode ? 1 : 2;
  const size_t okm_bytes = (size_t)h2c_okm_bytes(group, count) * n, pb = (size_t)144 * group * n;
  uint8_t *d_msgs, *d_dp, *d_extra;
  uint64_t *d_off;
  int dp_len;
  std::vector<uint64_t> rebased;
  int rc = stage_inputs(ctx, msgs, off, n, dst, dst_len, okm_bytes + 256 + pb, &d_msgs, &d_off, &d_dp, &dp_len, &d_extra, rebased);
  if (rc != B200_OK) return rc;
  uint8_t *d_okm = d_extra;
  char *d_out = (char *)d_extra + ((okm_bytes + 255) & ~(size_t)255);
  gpu_launcher l{ctx};
  rc = h2c_hash_run(l, group, (const uint8_t *)d_msgs, (const uint64_t *)d_off, n, (const uint8_t *)d_dp, dp_len, count, d_okm, d_out);
  if (rc < 0) return rc;
  B200_CUDA(ctx, cudaMemcpyAsync(out, d_out, pb, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

}  // namespace

#define CHECK_CTX(ctx)                      \
  if ((ctx) == nullptr) return B200_EINVAL; \
  ctx_guard guard__(ctx);                   \
  if (!guard__.ok) return B200_ENODEV

extern "C" {

int b200_expand_message_xmd_sha256(b200_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst,
                                   size_t dst_len, size_t len_in_bytes, uint8_t *out) {
  CHECK_CTX(ctx);
  if ((len_in_bytes + 31) / 32 > 255 || len_in_bytes > 65535) return B200_EINVAL;  // the reference panics (expand_msg.rs:243-248)
  if (n && (!offsets || !out)) return B200_EINVAL;
  if (dst_len && !dst) return B200_EINVAL;
  if (n == 0 || len_in_bytes == 0) return B200_OK;
  uint8_t *d_msgs, *d_dp, *d_out;
  uint64_t *d_off;
  int dp_len;
  std::vector<uint64_t> rebased;
  int rc = stage_inputs(ctx, msgs, offsets, n, dst, dst_len, len_in_bytes * n, &d_msgs, &d_off, &d_dp, &dp_len, &d_out, rebased);
  if (rc != B200_OK) return rc;
  B200_LAUNCH(ctx, k_h2c_expand, (unsigned)((n + 127) / 128), 128, 0, (const uint8_t *)d_msgs, (const uint64_t *)d_off, n,
              (const uint8_t *)d_dp, dp_len, (uint32_t)len_in_bytes, d_out);
  B200_CUDA(ctx, cudaMemcpyAsync(out, d_out, len_in_bytes * n, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
int b200_g1_hash_to_curve(b200_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len,
                          int encode, b200_g1_projective *out) {
  CHECK_CTX(ctx);
  if ((n && (!offsets || !out)) || (dst_len && !dst)) return B200_EINVAL;
  return n ? hash_host(ctx, 1, msgs, offsets, n, dst, dst_len, encode, out) : B200_OK;
}
int b200_g2_hash_to_curve(b200_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len,
                          int encode, b200_g2_projective *out) {
  CHECK_CTX(ctx);
  if ((n && (!offsets || !out)) || (dst_len && !dst)) return B200_EINVAL;
  return n ? hash_host(ctx, 2, msgs, offsets, n, dst, dst_len, encode, out) : B200_OK;
}
int b200_fr_from_okm(b200_ctx *ctx, const uint8_t *okm, size_t n, b200_fr *out) {
  CHECK_CTX(ctx);
  if (n && (!okm || !out)) return B200_EINVAL;
  if (n == 0) return B200_OK;
  int rc = stage_reserve(ctx, 48 * n + 32 * n + 4 * 256);
  if (rc != B200_OK) return rc;
  uint8_t *din = (uint8_t *)stage_take(ctx, 48 * n);
  char *dout = (char *)stage_take(ctx, 32 * n);
  B200_CUDA(ctx, cudaMemcpyAsync(din, okm, 48 * n, cudaMemcpyHostToDevice, ctx->stream));
  B200_LAUNCH(ctx, k_h2c_fr_from_okm, (unsigned)((n + 127) / 128), 128, 0, (const uint8_t *)din, n, dout);
  B200_CUDA(ctx, cudaMemcpyAsync(out, dout, 32 * n, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
int b200_fr_hash_to_field(b200_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len,
                          int count, b200_fr *out) {
  CHECK_CTX(ctx);
  if (count < 1 || count > 170 || (n && (!offsets || !out)) || (dst_len && !dst)) return B200_EINVAL;  // 48 * count <= 255 * 32
  if (n == 0) return B200_OK;
  const size_t len = (size_t)48 * count, ob = 32 * (size_t)count * n;
  uint8_t *d_msgs, *d_dp, *d_extra;
  uint64_t *d_off;
  int dp_len;
  std::vector<uint64_t> rebased;
  int rc = stage_inputs(ctx, msgs, offsets, n, dst, dst_len, len * n + 256 + ob, &d_msgs, &d_off, &d_dp, &dp_len, &d_extra, rebased);
  if (rc != B200_OK) return rc;
  uint8_t *d_okm = d_extra;
  char *d_out = (char *)d_extra + ((len * n + 255) & ~(size_t)255);
  B200_LAUNCH(ctx, k_h2c_expand, (unsigned)((n + 127) / 128), 128, 0, (const uint8_t *)d_msgs, (const uint64_t *)d_off, n,
              (const uint8_t *)d_dp, dp_len, (uint32_t)len, d_okm);
  const size_t m = n * (size_t)count;
  B200_LAUNCH(ctx, k_h2c_fr_from_okm, (unsigned)((m + 127) / 128), 128, 0, (const uint8_t *)d_okm, m, d_out);
  B200_CUDA(ctx, cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
int b200_h2c_stage(b200_ctx *ctx, int group, int kind, const void *in, size_t n, void *out) {
  CHECK_CTX(ctx);
  if ((group != 1 && group != 2) || kind < 0 || kind > 3 || (n && (!in || !out))) return B200_EINVAL;
  if (n == 0) return B200_OK;
  const size_t fb = 48 * (size_t)group, ib = ((kind == 0 || kind == 2) ? fb : 3 * fb) * n, ob = 3 * fb * n;
  int rc = stage_reserve(ctx, ib + ob + 4 * 256);
  if (rc != B200_OK) return rc;
  char *din = (char *)stage_take(ctx, ib), *dout = (char *)stage_take(ctx, ob);
  B200_CUDA(ctx, cudaMemcpyAsync(din, in, ib, cudaMemcpyHostToDevice, ctx->stream));
  if (group == 1)
    B200_LAUNCH(ctx, k_h2c_stage<fp>, (unsigned)((n + 127) / 128), 128, 0, kind, (const char *)din, n, dout);
  else
    B200_LAUNCH(ctx, k_h2c_stage<fp2>, (unsigned)((n + 127) / 128), 128, 0, kind, (const char *)din, n, dout);
  B200_CUDA(ctx, cudaMemcpyAsync(out, dout, ob, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

}  // extern "C"
