// b200_ctx: one CUDA device, one stream, grow-only scratch arena, launch counter, last-error text.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/bls12381_b200.h"

struct b200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;   // false: the caller's stream (b200_ctx_create_on_stream), never destroyed by the ctx
  // side streams of an MSM: bucket reduction (stream2) and the serial Horner chain (stream3) of one window
  // group overlap the bucket accumulation of the next group on `stream`
  cudaStream_t stream2 = nullptr, stream3 = nullptr;
  static constexpr int N_SYNC_EVENTS = 40;
  cudaEvent_t ev_sync[N_SYNC_EVENTS] = {};
  std::mutex mu;
  char err[256] = {0};
  uint64_t launches = 0;
  int msm_c = 0;
  // host-pointer MSM: the point upload runs on stream3 while the scalars (uploaded first) are being sorted on the main
  // stream; the first kernel that reads points waits for ev_sync[30] when this is set (msm_host in capi_msm.cu)
  bool msm_points_event_pending = false;
  uint32_t *msm_bad_flag = nullptr;   // device word set by the MSM digit kernel when a scalar is not canonical (>= q)
  // G1 MSM: GLV split k = k1 + k2*lambda (8 windows of 2n entries instead of 16 of n).  0 off (default), 1 on,
  // 2 = on for window-sharded calls only.  Implemented, parity-tested, and measured NEGATIVE at 2^20 on B200:
  // 1 GPU 9.34 vs 9.02 ms, 8 GPUs 3.66 vs 3.19 ms — halving the windows halves the (window x bucket) slots, i.e. the
  // thread-level parallelism of the bucket kernel (one 2^15-bucket window = 0.58 of a wave), which costs more than
  // the shorter reduction/Horner tail saves.  Kept as a knob for larger n / wider windows.
  int tune_g1_glv = 0;
  // batched-affine tree levels in front of the bucket kernel (msm_affine.cuh): -1 auto by bucket population, 0..3.
  // Implemented and parity-tested; OFF by default: at 2^20 (G1) measured 11.6 ms (2 levels) / 13.1 ms (3 levels) vs
  // 9.0 ms for the pure XYZZ bucket kernel — every warp of a wave reaches its binary-GCD inversion (ALU pipe) at the
  // same time instead of overlapping it with other warps' multiplications, and batches large enough to amortise it
  // (K >= 128 pairs per thread) leave too few threads at this size.  Groundwork for larger N.
  int tune_msm_affine_levels = 0;
  // bucket reduction: 0 = one thread per chunk of buckets (round 1), 1 = lane-cooperative (six lanes per chunk, round 2) for
  // the last — exposed — window group only (-1, the default, means this), 2 = lane-cooperative for every group.  The CPU test
  // harness defaults to 0 (every shuffle is two fiber barriers there: the emulated cooperative reduction costs minutes) and
  // switches it on for one small case (tests/test_msm_on_mock_cpu.py::test_cooperative_bucket_reduction).
#ifdef B200_HOST_EMUL
  int tune_msm_reduce = 0;
#else
  int tune_msm_reduce = -1;
#endif
  // 2-3 local windows (window shards of a multi-GPU MSM): one window per group (1) or one group (0).  Measured per shard of an
  // 8-way sharded 2^20 G1 MSM (profiles/records/r02_msm_shard_times.txt): one group 2.12 .. 3.05 ms, one window per group
  // 2.46 .. 3.22 ms — the overlapped reduce / Horner kernels of the upper window slow the lower window's bucket kernel by more
  // than they hide
  int tune_msm_tail_groups = 0;
  int tune_msm_reduce_min_chunk = 8;   // buckets per thread of the thread-per-chunk reduction, at least (capi_msm.cu)
  int tune_g1_prefetch = 1;    // G1 bucket kernel: cp.async double-buffered prefetch of the next point (1) or plain loads (0)
  int tune_pairing_chunks = 4; // independent Miller+final-exp chunks of a pairing batch kept in flight on 2 streams
  // G2 bucket kernel: 2 = accumulator in registers (255 regs, 2 blocks/SM); 3 = accumulator in shared memory, built
  // for 3 blocks/SM (168 regs); 4 = shared-memory accumulator, 2 blocks/SM (measured best: 31.2 vs 33.1 ms at 2^20)
  int tune_g2_acc_blocks = 4;
  int sm_count = 148;
  // scratch arena (device memory), bump-allocated per API call
  char *arena = nullptr;
  size_t arena_size = 0, arena_off = 0;
  // staging arena for the host-pointer entry points
  char *stage = nullptr;
  size_t stage_size = 0, stage_off = 0;
  uint32_t *inv_pow2 = nullptr;  // table of fp_inv.cuh (769 x 12 words), filled at ctx creation
  // optional per-kernel timing (CUDA events on the ctx stream around every launch)
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;       // 2 per record
  std::vector<const char *> ev_names;     // one per record
  // scalar-field NTT tables of capi_fr.cu (twiddles + coset powers for the last log_n used); freed by ctx_destroy
  void *fr_state = nullptr;
  void (*fr_state_free)(void *) = nullptr;
  // Miller loop / final exponentiation kernels: 4 = pairing_v4.cu, one thread per pairing (round 1).  The dual- / triple-
  // stream Fp2-multiply builds (v5 / v6) measured slower on B200 (63.5 / 67.4 vs 59.3 ms at 2^16 pairs) and were removed.
  // 7 = pairing_coop.cu: six lanes per pairing, Fp12 distributed over the lanes (round 2).
  // 0 = automatic (default) = the six-lane kernels at every batch size: one wave needs 8 880 pairs instead of 37 888 (9.1 vs
  // 22.0 ms at 8 192 pairs, 16.0 vs 22.7 ms at 16 384), and since the Montgomery reduction of the lazy dot products lost its
  // shift chains they also win on large batches (58.0 vs 60.8 ms at 65 536; before: 61.8 vs 59.3).  4 stays selectable.
  int tune_pairing_variant = 0;
  bool coop_for(size_t) const { return tune_pairing_variant != 4; }
  bool coop_products() const { return tune_pairing_variant != 4; }
  // scalar-multiplication batches (config 1): -1 thread per item, 0 auto (group kernel up to tune_mul_groups_max_n items, items
  // per warp = ceil(n / (4 * SMs)) clipped to 1..5), 1..5 forced items per warp, 6 = round 1's one warp per item
  int tune_mul_groups = 0;
  int tune_mul_groups_max_n = 9000;   // measured crossover with the thread-per-item kernel (G1): 8192 -> 5.05 vs 5.64 ms, 16384 -> 7.67 vs 5.63 ms
  // pairing batches of more than two waves: independent chunks on two streams (capi_pairing.cu).  2^16 pairs, whole-wave chunks:
  // 1 chunk 54.6 ms, 2: 52.5, 3: 51.1, 4: 53.8
  int tune_coop_chunks = 3;
  // G2 line coefficients for the six-lane kernels: batches of at most this many pairs use the six-lanes-per-Q kernel
  // (k_coop_g2_prepare: latency), larger ones one thread per Q (k_g2_prepare: throughput)
  int tune_coop_prepare_max = 5000;   // measured: 1024 pairs 1.22 vs 2.28 ms, 4096: 1.63 vs 2.29, 8192: 3.67 vs 2.32
  int tune_coop_warps = 12;          // warps per block (one block per SM) of the lane-cooperative pairing kernels, 1..16
  int tune_coop_split = 1;           // 1: Miller loop and final exponentiation of a pairing batch as two launches of the kernel
  bool coop_attr_done[8] = {};       // cudaFuncSetAttribute(max dynamic shared memory) done on this device, per kernel build
  // multi-GPU (capi_multi.cu): NCCL communicator of this rank, and a (world + 1) x 576-byte exchange buffer on the device
  void *nccl_comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  char *comm_buf = nullptr;       // cudaFuncSetAttribute(max dynamic shared memory) done on this device
};

namespace b200 {

inline int set_err(b200_ctx *ctx, cudaError_t e, const char *what) {
  snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, cudaGetErrorString(e));
  return e == cudaErrorMemoryAllocation ? B200_ENOMEM : B200_ECUDA;
}
#define B200_CUDA(ctx, call)                                         \
  do {                                                               \
    cudaError_t e__ = (call);                                        \
    if (e__ != cudaSuccess) return b200::set_err(ctx, e__, #call);   \
  } while (0)

// make sure the scratch arena holds `bytes` and reset the bump pointer
inline int arena_reserve(b200_ctx *ctx, size_t bytes) {
  ctx->arena_off = 0;
  if (bytes <= ctx->arena_size) return B200_OK;
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->arena) cudaFree(ctx->arena);
  ctx->arena = nullptr;
  ctx->arena_size = 0;
  size_t want = bytes + (bytes >> 3) + (1 << 20);
  B200_CUDA(ctx, cudaMalloc((void **)&ctx->arena, want));
  ctx->arena_size = want;
  return B200_OK;
}
template <class T>
inline T *arena_take(b200_ctx *ctx, size_t count) {
  size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
  T *p = reinterpret_cast<T *>(ctx->arena + ctx->arena_off);
  ctx->arena_off += bytes;
  return p;  // caller guarantees arena_reserve covered the total
}
inline size_t arena_pad(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

inline int stage_reserve(b200_ctx *ctx, size_t bytes) {
  ctx->stage_off = 0;
  if (bytes <= ctx->stage_size) return B200_OK;
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->stage) cudaFree(ctx->stage);
  ctx->stage = nullptr;
  ctx->stage_size = 0;
  size_t want = bytes + (1 << 16);
  B200_CUDA(ctx, cudaMalloc((void **)&ctx->stage, want));
  ctx->stage_size = want;
  return B200_OK;
}
inline void *stage_take(b200_ctx *ctx, size_t bytes) {
  void *p = ctx->stage + ctx->stage_off;
  ctx->stage_off += (bytes + 255) & ~(size_t)255;
  return p;
}

constexpr size_t MAX_TIMING_RECORDS = 8192;
// returns the index of the record opened (or -1): records an event BEFORE the launch
inline int timing_begin(b200_ctx *ctx, const char *name, cudaStream_t strm = nullptr) {
  if (!ctx->timing || ctx->ev_names.size() >= MAX_TIMING_RECORDS) return -1;
  if (strm == nullptr) strm = ctx->stream;
  size_t r = ctx->ev_names.size();
  while (ctx->ev_pool.size() < 2 * (r + 1)) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return -1;
    ctx->ev_pool.push_back(e);
  }
  ctx->ev_names.push_back(name);
  cudaEventRecord(ctx->ev_pool[2 * r], strm);
  return (int)r;
}
inline void timing_end(b200_ctx *ctx, int r, cudaStream_t strm = nullptr) {
  if (r >= 0) cudaEventRecord(ctx->ev_pool[2 * r + 1], strm ? strm : ctx->stream);
}

// dynamic shared memory of a kernel: `extern __shared__ T name[];` — in the CPU test harness a pointer into one static
// buffer (one block runs at a time there)
#ifdef B200_HOST_EMUL
#define B200_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emul_dyn_smem)
#else
#define B200_DYN_SMEM(type, name) extern __shared__ type name[]
#endif

// the one place a kernel is launched from.  The CPU test harness (tests/emul/, B200_HOST_EMUL) compiles the host side of
// the units with non-cooperative kernels against a mock runtime and turns a launch into a (block, thread) loop.
#ifdef B200_HOST_EMUL
#define B200_KERNEL_LAUNCH(kernel, grid, block, smem, strm, ...) \
  (emul_trace_launch(#kernel), emul_kernel_launch(kernel, grid, block, __VA_ARGS__))
#else
#define B200_KERNEL_LAUNCH(kernel, grid, block, smem, strm, ...) kernel<<<(grid), (block), (smem), (strm)>>>(__VA_ARGS__)
#endif

#define B200_LAUNCH(ctx, kernel, grid, block, smem, ...)                                 \
  do {                                                                                   \
    int tr__ = b200::timing_begin(ctx, #kernel);                                         \
    B200_KERNEL_LAUNCH(kernel, grid, block, smem, (ctx)->stream, __VA_ARGS__);           \
    b200::timing_end(ctx, tr__);                                                         \
    (ctx)->launches++;                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) return b200::set_err(ctx, e__, "launch " #kernel);           \
  } while (0)

// same, on an explicit stream of the ctx
#define B200_LAUNCH_ON(ctx, strm, kernel, grid, block, smem, ...)                        \
  do {                                                                                   \
    int tr__ = b200::timing_begin(ctx, #kernel, strm);                                   \
    B200_KERNEL_LAUNCH(kernel, grid, block, smem, strm, __VA_ARGS__);                    \
    b200::timing_end(ctx, tr__, strm);                                                   \
    (ctx)->launches++;                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) return b200::set_err(ctx, e__, "launch " #kernel);           \
  } while (0)

struct ctx_guard {
  b200_ctx *c;
  int prev = -1;
  bool ok = true;
  explicit ctx_guard(b200_ctx *ctx) : c(ctx) {
    c->mu.lock();
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != c->device && cudaSetDevice(c->device) != cudaSuccess) ok = false;
  }
  ~ctx_guard() {
    if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
    c->mu.unlock();
  }
};

}  // namespace b200
