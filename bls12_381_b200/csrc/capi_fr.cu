// C ABI, part 5: scalar-field (Fr) batch arithmetic and the NTT.  Kernels and the launch
// plan live in fr_ntt.cuh (shared with the CPU test harness); this file owns the device memory and the streams.
#include <new>

#include "ctx.cuh"
#include "fr_ntt.cuh"

using namespace b200;

namespace {

struct fr_state {
  int log_n = -1;
  char *mem = nullptr;
  fr_ntt_tables tb{};
};
void fr_state_free(void *p) {
  fr_state *st = static_cast<fr_state *>(p);
  if (st->mem) cudaFree(st->mem);
  delete st;
}

// launcher of fr_ntt.cuh's plans: a kernel launch on the ctx stream with the ctx's timing / launch accounting
struct gpu_launcher {
  b200_ctx *ctx;
  template <class K, class... A>
  int operator()(K k, unsigned grid, unsigned block, A... a) {
    int tr = timing_begin(ctx, "k_fr");
    B200_KERNEL_LAUNCH(k, grid, block, 0, ctx->stream, a...);
    timing_end(ctx, tr);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(ctx, e, "launch fr kernel");
    return 0;
  }
};

int ensure_tables(b200_ctx *ctx, int log_n, fr_state **out) {
  fr_state *st = static_cast<fr_state *>(ctx->fr_state);
  if (!st) {
    st = new (std::nothrow) fr_state();
    if (!st) return B200_ENOMEM;
    ctx->fr_state = st;
    ctx->fr_state_free = fr_state_free;
  }
  if (st->log_n != log_n) {
    B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (st->mem) cudaFree(st->mem);
    st->mem = nullptr;
    st->log_n = -1;
    B200_CUDA(ctx, cudaMalloc((void **)&st->mem, fr_tables_bytes(log_n)));
    gpu_launcher l{ctx};
    int rc = fr_ntt_build_tables(l, st->mem, log_n, &st->tb);
    if (rc < 0) return rc;
    st->log_n = log_n;
  }
  *out = st;
  return B200_OK;
}

inline unsigned nblk(size_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

bool fr_op_valid(int op) {
  return op == B200_OP_MUL || op == B200_OP_ADD || op == B200_OP_SUB || op == B200_OP_SQUARE || op == B200_OP_NEG ||
         op == B200_OP_INVERT || op == B200_OP_DOUBLE;
}
bool fr_op_binary(int op) { return op == B200_OP_MUL || op == B200_OP_ADD || op == B200_OP_SUB; }

int ntt_dev(b200_ctx *ctx, const char *in, int log_n, bool inverse, bool coset, char *out) {
  fr_state *st = nullptr;
  int rc = ensure_tables(ctx, log_n, &st);
  if (rc != B200_OK) return rc;
  const size_t bytes = (size_t)32 << log_n;
  if (in == out) {  // pass 0 gathers from bit-reversed positions: it cannot run in place
    rc = arena_reserve(ctx, bytes + 256);
    if (rc != B200_OK) return rc;
    char *tmp = arena_take<char>(ctx, bytes);
    B200_CUDA(ctx, cudaMemcpyAsync(tmp, in, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    in = tmp;
  }
  gpu_launcher l{ctx};
  rc = fr_ntt_run(l, in, out, log_n, inverse, coset, st->tb);
  return rc < 0 ? rc : B200_OK;
}

}  // namespace

#define CHECK_CTX(ctx)                      \
  if ((ctx) == nullptr) return B200_EINVAL; \
  ctx_guard guard__(ctx);                   \
  if (!guard__.ok) return B200_ENODEV

extern "C" {

int b200_fr_op_dev(b200_ctx *ctx, int op, const void *a, const void *b, size_t n, void *out) {
  CHECK_CTX(ctx);
  if (!fr_op_valid(op) || (n && (!a || !out || (fr_op_binary(op) && !b)))) return B200_EINVAL;
  if (n == 0) return B200_OK;
  B200_LAUNCH(ctx, k_fr_op, nblk(n, 256), 256, 0, op, (const char *)a, fr_op_binary(op) ? (const char *)b : nullptr,
              (char *)out, n);
  return B200_OK;
}
int b200_fr_op(b200_ctx *ctx, int op, const b200_fr *a, const b200_fr *b, size_t n, b200_fr *out) {
  CHECK_CTX(ctx);
  if (!fr_op_valid(op) || (n && (!a || !out || (fr_op_binary(op) && !b)))) return B200_EINVAL;
  if (n == 0) return B200_OK;
  const bool bin = fr_op_binary(op);
  int rc = stage_reserve(ctx, 3 * (32 * n + 256));
  if (rc != B200_OK) return rc;
  char *da = (char *)stage_take(ctx, 32 * n), *db = bin ? (char *)stage_take(ctx, 32 * n) : nullptr,
       *dout = (char *)stage_take(ctx, 32 * n);
  B200_CUDA(ctx, cudaMemcpyAsync(da, a, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
  if (bin) B200_CUDA(ctx, cudaMemcpyAsync(db, b, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
  B200_LAUNCH(ctx, k_fr_op, nblk(n, 256), 256, 0, op, (const char *)da, (const char *)db, dout, n);
  B200_CUDA(ctx, cudaMemcpyAsync(out, dout, 32 * n, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
int b200_fr_to_bytes(b200_ctx *ctx, const b200_fr *a, size_t n, b200_scalar *out) {
  CHECK_CTX(ctx);
  if (n && (!a || !out)) return B200_EINVAL;
  if (n == 0) return B200_OK;
  int rc = stage_reserve(ctx, 2 * (32 * n + 256));
  if (rc != B200_OK) return rc;
  char *da = (char *)stage_take(ctx, 32 * n), *dout = (char *)stage_take(ctx, 32 * n);
  B200_CUDA(ctx, cudaMemcpyAsync(da, a, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
  B200_LAUNCH(ctx, k_fr_to_bytes, nblk(n, 256), 256, 0, (const char *)da, dout, n);
  B200_CUDA(ctx, cudaMemcpyAsync(out, dout, 32 * n, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
int b200_fr_from_bytes(b200_ctx *ctx, const b200_scalar *in, size_t n, b200_fr *out, uint8_t *ok) {
  CHECK_CTX(ctx);
  if (n && (!in || !out || !ok)) return B200_EINVAL;
  if (n == 0) return B200_OK;
  int rc = stage_reserve(ctx, 2 * (32 * n + 256) + n + 256);
  if (rc != B200_OK) return rc;
  char *din = (char *)stage_take(ctx, 32 * n), *dout = (char *)stage_take(ctx, 32 * n);
  uint8_t *dok = (uint8_t *)stage_take(ctx, n);
  B200_CUDA(ctx, cudaMemcpyAsync(din, in, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
  B200_LAUNCH(ctx, k_fr_from_bytes, nblk(n, 256), 256, 0, (const char *)din, dout, dok, n);
  B200_CUDA(ctx, cudaMemcpyAsync(out, dout, 32 * n, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaMemcpyAsync(ok, dok, n, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
int b200_fr_ntt_dev(b200_ctx *ctx, const void *in, int log_n, int inverse, int coset, void *out) {
  CHECK_CTX(ctx);
  if (log_n < 0 || log_n > FR_NTT_MAX_LOG_N || !in || !out) return B200_EINVAL;
  return ntt_dev(ctx, (const char *)in, log_n, inverse != 0, coset != 0, (char *)out);
}
int b200_fr_ntt(b200_ctx *ctx, const b200_fr *in, int log_n, int inverse, int coset, b200_fr *out) {
  CHECK_CTX(ctx);
  if (log_n < 0 || log_n > FR_NTT_MAX_LOG_N || !in || !out) return B200_EINVAL;
  const size_t bytes = (size_t)32 << log_n;
  int rc = stage_reserve(ctx, 2 * (bytes + 256));
  if (rc != B200_OK) return rc;
  char *din = (char *)stage_take(ctx, bytes), *dout = (char *)stage_take(ctx, bytes);
  B200_CUDA(ctx, cudaMemcpyAsync(din, in, bytes, cudaMemcpyHostToDevice, ctx->stream));
  rc = ntt_dev(ctx, din, log_n, inverse != 0, coset != 0, dout);
  if (rc != B200_OK) return rc;
  B200_CUDA(ctx, cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

}  // extern "C"
