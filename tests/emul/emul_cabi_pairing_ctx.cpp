// TEST INFRASTRUCTURE.  Context helpers for the mock-runtime build of the PAIRING units (capi_pairing.cu +
// pairing_v4/v5/v6.cu): ctx create/destroy and the tuning keys those units read.
#include "cuda_host_shim.h"

#include "ctx.cuh"

extern "C" {
int b200_ctx_create(int, b200_ctx **out) {
  *out = new b200_ctx();
  (*out)->sm_count = 1;   // chunking threshold of pairing_dev = sm_count * 256 + 2048 pairs
  return B200_OK;
}
void b200_ctx_destroy(b200_ctx *ctx) {
  if (!ctx) return;
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->stage) cudaFree(ctx->stage);
  delete ctx;
}
const char *b200_strerror(int code) { return code == B200_OK ? "ok" : code == B200_EINVAL ? "invalid argument" : "error"; }
const char *b200_last_error(const b200_ctx *ctx) { return ctx ? ctx->err : ""; }
// the two keys the pairing units read, with the value checks of capi_basic.cu
int b200_ctx_set_tuning(b200_ctx *ctx, const char *key, int value) {
  if (!ctx || !key) return B200_EINVAL;
  if (!strcmp(key, "pairing_variant")) {
    if (value < 4 || value > 6) return B200_EINVAL;
    ctx->tune_pairing_variant = value;
  } else if (!strcmp(key, "pairing_chunks")) {
    if (value < 1 || value > 64) return B200_EINVAL;
    ctx->tune_pairing_chunks = value;
  } else {
    return B200_EINVAL;
  }
  return B200_OK;
}
}
