// C ABI, part 7: multi-GPU behind the library (SURVEY §8b "the ctx owns ... NCCL communicators", §8e).
//
// Two ways in, the same device path underneath:
//   * one process (or thread) per GPU:  b200_comm_unique_id on rank 0, the caller ships the 128-byte id to the other
//     ranks by whatever it has (MPI, a socket, torch.distributed in bench.py), then b200_ctx_comm_init(ctx, id, rank, world)
//     on every rank; b200_g{1,2}_msm_sharded_dev is then a collective call: shard -> ncclAllGather of the 144 / 288-byte
//     partials -> complete-add combine, all enqueued on the ctx stream with NO host synchronisation in between.
//   * one process, all GPUs:  b200_multi_create(n_gpus) owns one ctx per device (ncclCommInitAll) and one host thread
//     per device; b200_multi_g{1,2}_msm takes HOST pointers, sends every device only its slice of the points (point-range
//     sharding) or everything (window sharding), and returns the combined point.
// NCCL has no elliptic-curve reduction: "allreduce of partial sums" = all-gather + local complete adds.
// libnccl is loaded at run time (dlopen "libnccl.so.2"): a single-GPU user needs no NCCL at all.
#include <dlfcn.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <thread>

#include "ctx.cuh"

using namespace b200;

// internal entry points of the other units (enqueue only, no synchronisation, ctx mutex NOT taken)
int b200i_msm_enqueue(b200_ctx *ctx, int k, const void *points, const void *inf, const void *scalars, size_t n, int shard,
                      int n_shards, void *out);
int b200i_sum_enqueue(b200_ctx *ctx, int k, const void *parts, size_t n, void *out);
int b200i_msm_check(b200_ctx *ctx);   // after a stream synchronisation: B200_EINVAL when the last MSM saw a scalar >= q

namespace {

// ---- the few NCCL entry points used, resolved with dlsym (signatures as in nccl.h 2.x)
typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
struct nccl_api {
  void *h = nullptr;
  int (*GetUniqueId)(nccl_uid *) = nullptr;
  int (*CommInitRank)(nccl_comm *, int, nccl_uid, int) = nullptr;
  int (*CommInitAll)(nccl_comm *, int, const int *) = nullptr;
  int (*CommDestroy)(nccl_comm) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int /*ncclDataType_t*/, nccl_comm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};
nccl_api &nccl() {
  static nccl_api a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      a.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.h) break;
    }
    if (!a.h) return;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
    a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.h, "ncclCommInitAll");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(a.h, "ncclAllGather");
    a.GroupStart = (decltype(a.GroupStart))dlsym(a.h, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.h, "ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllGather && a.GroupStart && a.GroupEnd;
  });
  return a;
}
constexpr int NCCL_UINT8 = 1;  // ncclUint8 == ncclChar + 1 (nccl.h: ncclInt8 = 0, ncclUint8 = 1)

int nccl_fail(b200_ctx *ctx, int rc, const char *what) {
  nccl_api &a = nccl();
  snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, a.GetErrorString ? a.GetErrorString(rc) : "NCCL error");
  return B200_ENCCL;
}

// contiguous share [lo, hi) of n items for `rank` (pairs / points shard by index)
void index_range(size_t n, int rank, int world, size_t *lo, size_t *hi) {
  size_t base = n / world, rem = n % world;
  *lo = (size_t)rank * base + ((size_t)rank < rem ? rank : rem);
  *hi = *lo + base + ((size_t)rank < rem ? 1 : 0);
}

// shard -> all-gather -> combine on ctx->stream; every rank ends with the full sum in `out`
int msm_sharded_enqueue(b200_ctx *ctx, int k, const void *points, const void *inf, const void *scalars, size_t n, int mode,
                        void *out) {
  const size_t PB = (size_t)144 * k, AB = (size_t)96 * k;
  if (ctx->comm_world <= 1 || ctx->nccl_comm == nullptr) return b200i_msm_enqueue(ctx, k, points, inf, scalars, n, 0, 1, out);
  const int r = ctx->comm_rank, w = ctx->comm_world;
  char *parts = ctx->comm_buf;  // w partials + my own at the end
  char *mine = parts + PB * w;
  int rc;
  if (mode == B200_SHARD_POINTS) {
    size_t lo, hi;
    index_range(n, r, w, &lo, &hi);
    rc = b200i_msm_enqueue(ctx, k, (const char *)points + AB * lo, inf ? (const uint8_t *)inf + lo : nullptr,
                           (const char *)scalars + 32 * lo, hi - lo, 0, 1, mine);
  } else {
    rc = b200i_msm_enqueue(ctx, k, points, inf, scalars, n, r, w, mine);
  }
  if (rc != B200_OK) return rc;
  int nrc = nccl().AllGather(mine, parts, PB, NCCL_UINT8, (nccl_comm)ctx->nccl_comm, ctx->stream);
  if (nrc != 0) return nccl_fail(ctx, nrc, "ncclAllGather");
  return b200i_sum_enqueue(ctx, k, parts, (size_t)w, out);
}

}  // namespace

#define CHECK_CTX(ctx)                      \
  if ((ctx) == nullptr) return B200_EINVAL; \
  ctx_guard guard__(ctx);                   \
  if (!guard__.ok) return B200_ENODEV

// ---- single process, all GPUs
struct b200_multi {
  int n = 0;
  std::vector<b200_ctx *> ctx;
  // one worker thread per device: enqueueing ~100 launches per MSM from one host thread would serialise the devices
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv, cv_done;
  std::function<int(int)> job;
  uint64_t gen = 0;
  int pending = 0;
  std::vector<int> rc;
  bool stop = false;
  std::mutex api_mu;  // one call at a time
  int mode = B200_SHARD_POINTS;
};

namespace {
void multi_worker(b200_multi *m, int d) {
  uint64_t seen = 0;
  cudaSetDevice(m->ctx[d]->device);
  for (;;) {
    std::function<int(int)> job;
    {
      std::unique_lock<std::mutex> lk(m->mu);
      m->cv.wait(lk, [&] { return m->stop || m->gen != seen; });
      if (m->stop) return;
      seen = m->gen;
      job = m->job;
    }
    int r = job(d);
    {
      std::lock_guard<std::mutex> lk(m->mu);
      m->rc[d] = r;
      if (--m->pending == 0) m->cv_done.notify_all();
    }
  }
}
// run job(d) on every device's worker thread, wait for all; returns the first error
int multi_run(b200_multi *m, std::function<int(int)> job) {
  std::unique_lock<std::mutex> lk(m->mu);
  m->job = std::move(job);
  m->pending = m->n;
  m->gen++;
  m->cv.notify_all();
  m->cv_done.wait(lk, [&] { return m->pending == 0; });
  for (int d = 0; d < m->n; d++)
    if (m->rc[d] != B200_OK) return m->rc[d];
  return B200_OK;
}

template <int K>
int multi_msm(b200_multi *m, const void *points, const uint8_t *inf, const void *scalars, size_t n, void *out) {
  constexpr size_t AB = 96 * K, PB = 144 * K;
  std::lock_guard<std::mutex> api(m->api_mu);
  const int w = m->n, mode = m->mode;
  int rc = multi_run(m, [&](int d) -> int {
    b200_ctx *c = m->ctx[d];
    std::lock_guard<std::mutex> g(c->mu);
    size_t lo = 0, hi = n;
    if (mode == B200_SHARD_POINTS) index_range(n, d, w, &lo, &hi);
    size_t cnt = hi - lo;
    int r = stage_reserve(c, AB * cnt + 33 * cnt + PB + 8 * 256);
    if (r != B200_OK) return r;
    void *dp = stage_take(c, AB * cnt), *ds = stage_take(c, 32 * cnt), *di = inf ? stage_take(c, cnt) : nullptr;
    void *dout = stage_take(c, PB);
    if (cnt) {
      // scalars first (main stream), points on stream3 under the counting sort — same schedule as the one-GPU b200_g1_msm
      B200_CUDA(c, cudaEventRecord(c->ev_sync[31], c->stream));
      B200_CUDA(c, cudaStreamWaitEvent(c->stream3, c->ev_sync[31], 0));
      B200_CUDA(c, cudaMemcpyAsync(ds, (const char *)scalars + 32 * lo, 32 * cnt, cudaMemcpyHostToDevice, c->stream));
      if (inf) B200_CUDA(c, cudaMemcpyAsync(di, inf + lo, cnt, cudaMemcpyHostToDevice, c->stream));
      B200_CUDA(c, cudaMemcpyAsync(dp, (const char *)points + AB * lo, AB * cnt, cudaMemcpyHostToDevice, c->stream3));
      B200_CUDA(c, cudaEventRecord(c->ev_sync[30], c->stream3));
      c->msm_points_event_pending = true;
    }
    char *parts = c->comm_buf, *mine = parts + PB * w;
    if (mode == B200_SHARD_POINTS)
      r = b200i_msm_enqueue(c, K, dp, di, ds, cnt, 0, 1, w > 1 ? (void *)mine : dout);
    else
      r = b200i_msm_enqueue(c, K, dp, di, ds, cnt, d, w, w > 1 ? (void *)mine : dout);
    if (c->msm_points_event_pending) {
      c->msm_points_event_pending = false;
      cudaStreamWaitEvent(c->stream, c->ev_sync[30], 0);
    }
    if (r != B200_OK) return r;
    if (w > 1) {
      int nrc = nccl().AllGather(mine, parts, PB, NCCL_UINT8, (nccl_comm)c->nccl_comm, c->stream);
      if (nrc != 0) return nccl_fail(c, nrc, "ncclAllGather");
      if (d == 0) {
        r = b200i_sum_enqueue(c, K, parts, (size_t)w, dout);
        if (r != B200_OK) return r;
      }
    }
    if (d == 0) B200_CUDA(c, cudaMemcpyAsync(out, dout, PB, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(c, cudaStreamSynchronize(c->stream));
    return b200i_msm_check(c);
  });
  return rc;
}
}  // namespace

extern "C" {

int b200_comm_unique_id(uint8_t *id) {
  if (!id) return B200_EINVAL;
  nccl_api &a = nccl();
  if (!a.ok) return B200_ENCCL;
  nccl_uid u;
  if (a.GetUniqueId(&u) != 0) return B200_ENCCL;
  memcpy(id, u.internal, B200_COMM_ID_BYTES);
  return B200_OK;
}
int b200_ctx_comm_init(b200_ctx *ctx, const uint8_t *id, int rank, int world) {
  CHECK_CTX(ctx);
  if (!id || world < 1 || rank < 0 || rank >= world || ctx->nccl_comm) return B200_EINVAL;
  nccl_api &a = nccl();
  if (!a.ok) {
    snprintf(ctx->err, sizeof(ctx->err), "libnccl.so.2 not found or incomplete");
    return B200_ENCCL;
  }
  nccl_uid u;
  memcpy(u.internal, id, B200_COMM_ID_BYTES);
  nccl_comm comm = nullptr;
  int nrc = a.CommInitRank(&comm, world, u, rank);
  if (nrc != 0) return nccl_fail(ctx, nrc, "ncclCommInitRank");
  ctx->nccl_comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  B200_CUDA(ctx, cudaMalloc((void **)&ctx->comm_buf, (size_t)(world + 1) * 576));
  return B200_OK;
}
int b200_ctx_comm_destroy(b200_ctx *ctx) {
  CHECK_CTX(ctx);
  if (ctx->nccl_comm) {
    cudaStreamSynchronize(ctx->stream);
    nccl().CommDestroy((nccl_comm)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  if (ctx->comm_buf) cudaFree(ctx->comm_buf);
  ctx->comm_buf = nullptr;
  ctx->comm_rank = 0;
  ctx->comm_world = 1;
  return B200_OK;
}
int b200_ctx_comm_rank(const b200_ctx *ctx) { return ctx ? ctx->comm_rank : -1; }
int b200_ctx_comm_world(const b200_ctx *ctx) { return ctx ? ctx->comm_world : -1; }

int b200_g1_msm_sharded_dev(b200_ctx *ctx, const void *points, const void *inf, const void *scalars, size_t n, int mode,
                            void *out) {
  CHECK_CTX(ctx);
  if (!out || (n && (!points || !scalars)) || (mode != B200_SHARD_WINDOWS && mode != B200_SHARD_POINTS)) return B200_EINVAL;
  int rc = msm_sharded_enqueue(ctx, 1, points, inf, scalars, n, mode, out);
  if (rc != B200_OK) return rc;
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return b200i_msm_check(ctx);
}
int b200_g2_msm_sharded_dev(b200_ctx *ctx, const void *points, const void *inf, const void *scalars, size_t n, int mode,
                            void *out) {
  CHECK_CTX(ctx);
  if (!out || (n && (!points || !scalars)) || (mode != B200_SHARD_WINDOWS && mode != B200_SHARD_POINTS)) return B200_EINVAL;
  int rc = msm_sharded_enqueue(ctx, 2, points, inf, scalars, n, mode, out);
  if (rc != B200_OK) return rc;
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return b200i_msm_check(ctx);
}

int b200_multi_create(int n_gpus, b200_multi **out) {
  if (!out) return B200_EINVAL;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return B200_ENODEV;
  if (n_gpus <= 0) n_gpus = count;
  if (n_gpus > count) return B200_EINVAL;
  b200_multi *m = new (std::nothrow) b200_multi();
  if (!m) return B200_ENOMEM;
  m->n = n_gpus;
  m->rc.assign(n_gpus, B200_OK);
  int rc = B200_OK;
  for (int d = 0; d < n_gpus && rc == B200_OK; d++) {
    b200_ctx *c = nullptr;
    rc = b200_ctx_create(d, &c);
    if (rc == B200_OK) m->ctx.push_back(c);
  }
  if (rc == B200_OK && n_gpus > 1) {
    nccl_api &a = nccl();
    if (!a.ok) rc = B200_ENCCL;
    if (rc == B200_OK) {
      std::vector<nccl_comm> comms(n_gpus, nullptr);
      std::vector<int> devs(n_gpus);
      for (int d = 0; d < n_gpus; d++) devs[d] = d;
      if (a.CommInitAll(comms.data(), n_gpus, devs.data()) != 0) rc = B200_ENCCL;
      for (int d = 0; d < n_gpus && rc == B200_OK; d++) {
        m->ctx[d]->nccl_comm = comms[d];
        m->ctx[d]->comm_rank = d;
        m->ctx[d]->comm_world = n_gpus;
      }
    }
  }
  for (int d = 0; d < (int)m->ctx.size() && rc == B200_OK; d++) {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(d);
    if (cudaMalloc((void **)&m->ctx[d]->comm_buf, (size_t)(n_gpus + 1) * 576) != cudaSuccess) rc = B200_ENOMEM;
    cudaSetDevice(prev);
  }
  if (rc != B200_OK) {
    for (b200_ctx *c : m->ctx) {
      b200_ctx_comm_destroy(c);
      b200_ctx_destroy(c);
    }
    delete m;
    return rc;
  }
  for (int d = 0; d < n_gpus; d++) m->workers.emplace_back(multi_worker, m, d);
  *out = m;
  return B200_OK;
}
void b200_multi_destroy(b200_multi *m) {
  if (!m) return;
  {
    std::lock_guard<std::mutex> lk(m->mu);
    m->stop = true;
    m->cv.notify_all();
  }
  for (std::thread &t : m->workers) t.join();
  for (b200_ctx *c : m->ctx) {
    b200_ctx_comm_destroy(c);
    b200_ctx_destroy(c);
  }
  delete m;
}
int b200_multi_gpus(const b200_multi *m) { return m ? m->n : 0; }
b200_ctx *b200_multi_ctx(b200_multi *m, int i) { return (m && i >= 0 && i < m->n) ? m->ctx[i] : nullptr; }
int b200_multi_set_sharding(b200_multi *m, int mode) {
  if (!m || (mode != B200_SHARD_WINDOWS && mode != B200_SHARD_POINTS)) return B200_EINVAL;
  std::lock_guard<std::mutex> api(m->api_mu);
  m->mode = mode;
  return B200_OK;
}
int b200_multi_g1_msm(b200_multi *m, const b200_g1_affine *points, const uint8_t *inf, const b200_scalar *scalars, size_t n,
                      b200_g1_projective *out) {
  if (!m || !out || (n && (!points || !scalars))) return B200_EINVAL;
  return multi_msm<1>(m, points, inf, scalars, n, out);
}
int b200_multi_g2_msm(b200_multi *m, const b200_g2_affine *points, const uint8_t *inf, const b200_scalar *scalars, size_t n,
                      b200_g2_projective *out) {
  if (!m || !out || (n && (!points || !scalars))) return B200_EINVAL;
  return multi_msm<2>(m, points, inf, scalars, n, out);
}

}  // extern "C"
