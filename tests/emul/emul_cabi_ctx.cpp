// TEST INFRASTRUCTURE.  Context + error helpers for the mock-runtime build of the C ABI units (see mock/cuda_runtime.h).
#include "cuda_host_shim.h"

#include "ctx.cuh"

extern "C" {
int b200_ctx_create(int, b200_ctx **out) {
  *out = new b200_ctx();
  return B200_OK;
}
void b200_ctx_destroy(b200_ctx *ctx) {
  if (!ctx) return;
  if (ctx->fr_state && ctx->fr_state_free) ctx->fr_state_free(ctx->fr_state);
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->stage) cudaFree(ctx->stage);
  delete ctx;
}
const char *b200_strerror(int code) { return code == B200_OK ? "ok" : code == B200_EINVAL ? "invalid argument" : "error"; }
const char *b200_last_error(const b200_ctx *ctx) { return ctx ? ctx->err : ""; }
uint64_t b200_ctx_launch_count(const b200_ctx *ctx) { return ctx->launches; }
}
