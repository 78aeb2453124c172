// C ABI, part 1: context, batched field-tower ops (parity surface), element-wise group ops,
// batched scalar multiplication (BASELINE config 1), batch_normalize, partial-sum combine,
// and the IMAD-peak microbenchmark.  See include/bls12381_b200.h for the contract.
#include <new>

#include "ctx.cuh"
#include "curve.cuh"
#include "curve_warp.cuh"
#include "fp_inv.cuh"
#include "pairing.cuh"

using namespace b200;

// ================================================================ kernels
namespace {

template <int LEVEL>
__global__ void k_tower_op(int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, const uint32_t *pow2) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int W = 6 * LEVEL;
  const uint64_t *pa = a + W * i, *pb = b ? b + W * i : nullptr;
  uint64_t *po = out + W * i;
  if constexpr (LEVEL == 1) {
    fp x = fp_load(pa), y = pb ? fp_load(pb) : fp_zero(), r;
    switch (op) {
      case B200_OP_MUL: r = fp_mul(x, y); break;
      case B200_OP_ADD: r = fp_add(x, y); break;
      case B200_OP_SUB: r = fp_sub(x, y); break;
      case B200_OP_SQUARE: r = fp_sqr_c(x); break;   // the dedicated squaring the curve code uses
      case B200_OP_NEG: r = fp_neg(x); break;
      case B200_OP_INVERT_FAST: r = fp_inv_fast(x, pow2); break;
      default: r = fp_inv(x); break;
    }
    fp_store(po, r);
  } else if constexpr (LEVEL == 2) {
    fp2 x = fp2_load(pa), y = pb ? fp2_load(pb) : fp2_zero(), r;
    switch (op) {
      case B200_OP_MUL: r = M2(x, y); break;  // the called, lazily reduced multiply every G2/pairing kernel uses
      case B200_OP_ADD: r = fp2_add(x, y); break;
      case B200_OP_SUB: r = fp2_sub(x, y); break;
      case B200_OP_SQUARE: r = S2(x); break;
      case B200_OP_NEG: r = fp2_neg(x); break;
      case B200_OP_INVERT: r = fp2_inv(x); break;
      case B200_OP_MUL_BY_NONRESIDUE: r = fp2_mul_by_nonresidue(x); break;
      default: r = fp2_conj(x); break;  // frobenius == conjugate (src/fp2.rs:141)
    }
    fp2_store(po, r);
  } else if constexpr (LEVEL == 6) {
    fp6 x, y, r;
    fp6_load(&x, pa);
    if (pb) fp6_load(&y, pb);
    switch (op) {
      case B200_OP_MUL: fp6_mul(&r, &x, &y); break;
      case B200_OP_ADD: fp6_add(&r, &x, &y); break;
      case B200_OP_SUB: fp6_sub(&r, &x, &y); break;
      case B200_OP_SQUARE: fp6_sqr(&r, &x); break;
      case B200_OP_NEG: fp6_neg(&r, &x); break;
      case B200_OP_INVERT: fp6_inv(&r, &x); break;
      case B200_OP_FROBENIUS: fp6_frobenius(&r, &x); break;
      default: fp6_mul_by_nonresidue(&r, &x); break;
    }
    fp6_store(po, &r);
  } else {
    fp12 x, y, r;
    fp12_load(&x, pa);
    if (pb) fp12_load(&y, pb);
    switch (op) {
      case B200_OP_MUL: fp12_mul(&r, &x, &y); break;
      case B200_OP_SQUARE: fp12_sqr(&r, &x); break;
      case B200_OP_INVERT: fp12_inv(&r, &x); break;
      case B200_OP_FROBENIUS: fp12_frobenius(&r, &x); break;
      case B200_OP_CONJUGATE: fp12_conj(&r, &x); break;
      default: cyclotomic_square(&r, &x); break;
    }
    fp12_store(po, &r);
  }
}

template <class F>
__global__ void k_double(const char *p, char *out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  proj_store<F>(out + PB * i, proj_double(proj_load<F>(p + PB * i)));
}
template <class F>
__global__ void k_add(const char *p, const char *q, char *out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  proj_store<F>(out + PB * i, proj_add(proj_load<F>(p + PB * i), proj_load<F>(q + PB * i)));
}
template <class F>
__global__ void k_add_mixed(const char *p, const char *qxy, const uint8_t *qinf, char *out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  proj_store<F>(out + PB * i, proj_add_mixed(proj_load<F>(p + PB * i), affine_load<F>(qxy, qinf, i)));
}
// BASELINE config 1: one thread per (point, scalar); the reference's 255-step double-and-add
template <class F>
__global__ void __launch_bounds__(128) k_mul_batch(const char *p, const uint32_t *s, char *out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  uint32_t by[8];
  const uint4 *sp = reinterpret_cast<const uint4 *>(s + 8 * i);
  uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
  by[0] = lo.x; by[1] = lo.y; by[2] = lo.z; by[3] = lo.w;
  by[4] = hi.x; by[5] = hi.y; by[6] = hi.z; by[7] = hi.w;
  proj_store<F>(out + PB * i, proj_multiply(proj_load<F>(p + PB * i), by));
}

// Small batches (config 1 has 1024 items = 8 warps' worth of threads on 148 SMs): one WARP per item with the
// lane-parallel group operations of curve_warp.cuh — 3 dependent multiplication levels per doubling and 2 per
// addition instead of 8 + 12 sequential ones.  Same formulas and values as multiply(); the addition is
// skipped (not computed-and-discarded) when the scalar bit is 0 — the GPU path is variable-time by design.
template <class F>
__global__ void __launch_bounds__(128) k_mul_batch_warp(const char *p, const uint32_t *s, char *out, size_t n) {
  size_t item = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (item >= n) return;
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  uint32_t by[8];
  const uint4 *sp = reinterpret_cast<const uint4 *>(s + 8 * item);
  uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
  by[0] = lo.x; by[1] = lo.y; by[2] = lo.z; by[3] = lo.w;
  by[4] = hi.x; by[5] = hi.y; by[6] = hi.z; by[7] = hi.w;
  proj<F> base = proj_load<F>(p + PB * item), acc = proj_identity<F>();
#pragma unroll 1
  for (int bit = 254; bit >= 0; bit--) {
    acc = warp_double(acc, lane);
    if ((by[bit >> 5] >> (bit & 31)) & 1) acc = warp_add(acc, base, lane);
  }
  if (lane == 0) proj_store<F>(out + PB * item, acc);
}

// The same schedule with SEVERAL items per warp: item g of a warp is worked on by the six lanes 6g .. 6g+5 (grp_double / grp_add
// of curve_warp.cuh), `gpw` = 1..5 items per warp.  A warp-wide field multiplication costs the same issue slots whether 6 or 30
// lanes are active, and one warp per scheduler (SMSP) already runs at the latency of its dependent chain, so the best shape for
// a small batch is the SMALLEST gpw that leaves at most one warp per SMSP (config 1: 1024 items -> 2 per warp = 512 warps on
// 592 SMSPs instead of 1024 warps, two per SMSP).  The addition is executed when ANY item of the warp has its bit set and
// selected per item.
template <class F>
__global__ void __launch_bounds__(32) k_mul_batch_grp(const char *p, const uint32_t *s, char *out, size_t n, int gpw) {
  const int lane = threadIdx.x & 31;
  int grp = lane / 6;
  const int sub = lane - 6 * grp;
  const size_t item = (size_t)blockIdx.x * gpw + grp;
  const bool live = grp < gpw && item < n;
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  uint32_t by[8];
#pragma unroll
  for (int k = 0; k < 8; k++) by[k] = 0;
  proj<F> base = proj_identity<F>(), acc = proj_identity<F>();
  if (live) {
    const uint4 *sp = reinterpret_cast<const uint4 *>(s + 8 * item);
    uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
    by[0] = lo.x; by[1] = lo.y; by[2] = lo.z; by[3] = lo.w;
    by[4] = hi.x; by[5] = hi.y; by[6] = hi.z; by[7] = hi.w;
    base = proj_load<F>(p + PB * item);
  }
  const int gbase = 6 * grp;
#pragma unroll 1
  for (int bit = 254; bit >= 0; bit--) {
    acc = grp_double(acc, sub, gbase);
    const bool mine = (by[bit >> 5] >> (bit & 31)) & 1;
    if (__any_sync(0xffffffffu, mine)) {
      proj<F> t = grp_add(acc, base, sub, gbase);
      acc = proj_select(acc, t, mine);
    }
  }
  if (live && sub == 0) proj_store<F>(out + PB * item, acc);
}

// batch_normalize (src/g1.rs:806-839): Montgomery's trick per thread over a strided subsequence
// i = t, t+T, t+2T, ... (coalesced across the warp).  One inversion per thread.  The prefix products
// are parked in out.x exactly like the reference parks them in q.x.
B200_DEV fp f_inv_gcd(const fp &a, const uint32_t *pow2) { return fp_inv_fast(a, pow2); }
B200_DEV fp2 f_inv_gcd(const fp2 &a, const uint32_t *pow2) {  // src/fp2.rs:300-320 over the binary-GCD Fp inverse
  fp t = fp_inv_fast(fp_add(fp_mul_c(a.c0, a.c0), fp_mul_c(a.c1, a.c1)), pow2);
  return fp2{fp_mul_c(a.c0, t), fp_mul_c(a.c1, fp_neg(t))};
}
template <class F>
__global__ void __launch_bounds__(128) k_batch_normalize(const char *p, size_t n, char *oxy, uint8_t *oinf,
                                                       const uint32_t *pow2) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x, T = (size_t)gridDim.x * blockDim.x;
  if (t >= n) return;
  constexpr size_t FB = field_traits<F>::bytes, PB = 3 * FB, AB = 2 * FB;
  F acc = field_traits<F>::one();
  for (size_t i = t; i < n; i += T) {
    f_store(oxy + AB * i, acc);
    F z = field_traits<F>::load(p + PB * i + 2 * FB);
    if (!f_is_zero(z)) acc = f_mul(acc, z);
  }
  acc = f_inv_gcd(acc, pow2);  // one inversion per thread: binary GCD (fp_inv.cuh), same value as the Fermat inverse
  size_t last = t + ((n - 1 - t) / T) * T;
  for (size_t i = last;; i -= T) {
    F z = field_traits<F>::load(p + PB * i + 2 * FB);
    bool skip = f_is_zero(z);
    F tmp = f_mul(field_traits<F>::load(oxy + AB * i), acc);
    if (!skip) acc = f_mul(acc, z);
    affine<F> a{f_mul(field_traits<F>::load(p + PB * i), tmp), f_mul(field_traits<F>::load(p + PB * i + FB), tmp), false};
    if (skip) a = affine_identity<F>();
    affine_store<F>(oxy, oinf, i, a);
    if (i == t) break;
  }
}

// out = sum of n projective points: one block, tree in shared memory with complete adds
template <class F>
__global__ void __launch_bounds__(128) k_sum(const char *parts, size_t n, char *out) {
  B200_DYN_SMEM(char, smem);
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  proj<F> acc = proj_identity<F>();
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) acc = proj_add(acc, proj_load<F>(parts + PB * i));
  proj_store<F>(smem + PB * threadIdx.x, acc);
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      proj<F> a = proj_load<F>(smem + PB * threadIdx.x), b = proj_load<F>(smem + PB * (threadIdx.x + s));
      proj_store<F>(smem + PB * threadIdx.x, proj_add(a, b));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) proj_store<F>(out, proj_load<F>(smem));
}

// dependent-free IMAD.WIDE.U32 streams: 8 independent accumulators per thread
__global__ void __launch_bounds__(256) k_imad_peak(int iters, uint64_t *sink) {
  uint64_t acc[8];
  uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 977u;
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = (uint64_t)k * 0x9e3779b97f4a7c15ull + threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
#pragma unroll
      for (int k = 0; k < 8; k++)
#ifdef B200_HOST_EMUL  // CPU test harness: the same multiply-add in C
        acc[k] += (uint64_t)(a + k) * (b + r);
#else
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a + k), "r"(b + r));
#endif
    }
  }
  uint64_t x = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) x ^= acc[k];
  if (x == 0x1234567ull) sink[0] = x;  // never true in practice; keeps the chain alive
}

// the two halves of the product as separate instructions: mad.lo + mad.hi per product (two multiplier instructions per
// 32x32 -> 64 product; without the carry link between the halves it is not even a full 64-bit accumulate).  Audit of the roofline denominator: shows whether the wide form is the cheapest way
// to a 64-bit product on sm_100a.  Counted in the same unit (one 32x32+64 multiply-add per lo/hi PAIR).
__global__ void __launch_bounds__(256) k_imad_peak_lohi(int iters, uint64_t *sink) {
  uint32_t lo[8], hi[8];
  uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 977u;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    lo[k] = k * 0x9e3779b9u + threadIdx.x;
    hi[k] = k * 0x7f4a7c15u + blockIdx.x;
  }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
#ifdef B200_HOST_EMUL
        uint64_t t = (uint64_t)(a + k) * (b + r);
        lo[k] += (uint32_t)t;
        hi[k] += (uint32_t)(t >> 32);
#else
        // (a carry-linked mad.lo.cc / madc.hi pair is fused by ptxas into ONE IMAD.WIDE.U32 — which is how fp_mul gets its wide
        // multiply-adds; the unlinked pair below stays two instructions, IMAD + IMAD.HI.U32)
        asm volatile("mad.lo.u32 %0, %2, %3, %0;\n\tmad.hi.u32 %1, %2, %3, %1;" : "+r"(lo[k]), "+r"(hi[k]) : "r"(a + k), "r"(b + r));
#endif
      }
    }
  }
  uint32_t x = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) x ^= lo[k] ^ hi[k];
  if (x == 0x1234567u) sink[0] = x;
}

inline unsigned nblk(size_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

template <class F>
int mul_batch_dev(b200_ctx *ctx, const void *p, const void *s, size_t n, void *out) {
  if (n == 0) return B200_OK;
  // latency regime: several lanes per item.  tune_mul_groups: -1 = thread per item always, 0 = auto, 1..5 = items per warp of the
  // group kernel, 6 = round 1's one-warp-per-item kernel
  const int mg = ctx->tune_mul_groups;
  if (mg == 6) {
    B200_LAUNCH(ctx, k_mul_batch_warp<F>, nblk(n * 32, 128), 128, 0, (const char *)p, (const uint32_t *)s, (char *)out, n);
    return B200_OK;
  }
  if (mg >= 1 || (mg == 0 && n <= (size_t)ctx->tune_mul_groups_max_n)) {
    int gpw = mg;
    if (gpw < 1) {  // the smallest count that leaves at most one warp per scheduler; beyond that, full warps
      const size_t slots = 4u * (size_t)ctx->sm_count;
      gpw = (int)((n + slots - 1) / slots);
      if (gpw < 1) gpw = 1;
    }
    if (gpw > 5) gpw = 5;
    B200_LAUNCH(ctx, k_mul_batch_grp<F>, nblk(n, gpw), 32, 0, (const char *)p, (const uint32_t *)s, (char *)out, n, gpw);
    return B200_OK;
  }
  // small batches: 32-thread blocks so the work spreads over more SMs
  unsigned block = n <= 32u * 1024u ? 32u : 128u;
  B200_LAUNCH(ctx, k_mul_batch<F>, nblk(n, block), block, 0, (const char *)p, (const uint32_t *)s, (char *)out, n);
  return B200_OK;
}
template <class F>
int batch_normalize_dev(b200_ctx *ctx, const void *p, size_t n, void *oxy, void *oinf) {
  if (n == 0) return B200_OK;
  // ~64 points per thread amortises the 613-FpM inversion; never more threads than points
  size_t threads = (n + 63) / 64;
  unsigned block = 128;
  unsigned grid = nblk(threads, block);
  B200_LAUNCH(ctx, k_batch_normalize<F>, grid, block, 0, (const char *)p, n, (char *)oxy, (uint8_t *)oinf, ctx->inv_pow2);
  return B200_OK;
}
template <class F>
int sum_dev(b200_ctx *ctx, const void *parts, size_t n, void *out) {
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  unsigned block = 32;
  while (block < 128 && block < n) block <<= 1;
  B200_LAUNCH(ctx, k_sum<F>, 1, block, PB * block, (const char *)parts, n, (char *)out);
  return B200_OK;
}

// host-pointer wrapper: stage inputs, run, copy back
struct stager {
  b200_ctx *ctx;
  int rc = B200_OK;
  explicit stager(b200_ctx *c, size_t total) : ctx(c) { rc = stage_reserve(c, total + 8 * 256); }
  void *in(const void *host, size_t bytes) {
    if (rc != B200_OK || host == nullptr) return nullptr;
    void *d = stage_take(ctx, bytes);
    if (bytes) {
      cudaError_t e = cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, ctx->stream);
      if (e != cudaSuccess) rc = set_err(ctx, e, "H2D");
    }
    return d;
  }
  void *out(size_t bytes) { return rc == B200_OK ? stage_take(ctx, bytes) : nullptr; }
  int back(void *host, const void *dev, size_t bytes) {
    if (rc != B200_OK) return rc;
    if (bytes && host) {
      cudaError_t e = cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream);
      if (e != cudaSuccess) return rc = set_err(ctx, e, "D2H");
    }
    return rc;
  }
  int sync() {
    if (rc != B200_OK) return rc;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = set_err(ctx, e, "sync");
    return rc;
  }
};

}  // namespace

#define CHECK_CTX(ctx)                 \
  if ((ctx) == nullptr) return B200_EINVAL; \
  ctx_guard guard__(ctx);              \
  if (!guard__.ok) return B200_ENODEV

// ================================================================ context
// enqueue-only combine for capi_multi.cu (no lock, no synchronisation)
int b200i_sum_enqueue(b200_ctx *ctx, int k, const void *parts, size_t n, void *out) {
  return k == 1 ? sum_dev<fp>(ctx, parts, n, out) : sum_dev<fp2>(ctx, parts, n, out);
}

extern "C" {

int b200_ctx_create(int device, b200_ctx **out) { return b200_ctx_create_on_stream(device, nullptr, out); }
int b200_ctx_create_on_stream(int device, void *stream, b200_ctx **out) {
  if (out == nullptr) return B200_EINVAL;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return B200_ENODEV;
  int prev = 0;
  if (cudaGetDevice(&prev) != cudaSuccess) return B200_ENODEV;
  if (device < 0) device = prev;
  if (device >= count) return B200_EINVAL;
  b200_ctx *c = new (std::nothrow) b200_ctx();
  if (!c) return B200_ENOMEM;
  c->device = device;
  c->own_stream = stream == nullptr;
  if (!c->own_stream) c->stream = (cudaStream_t)stream;
  bool ok = cudaSetDevice(device) == cudaSuccess &&
            (!c->own_stream || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess) &&
            cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; ok && i < b200_ctx::N_SYNC_EVENTS; i++) ok = cudaEventCreateWithFlags(&c->ev_sync[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    delete c;
    cudaSetDevice(prev);
    return B200_ENODEV;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
  if (cudaMalloc((void **)&c->inv_pow2, FP_INV_TABLE_WORDS * sizeof(uint32_t)) != cudaSuccess) {
    b200_ctx_destroy(c);
    cudaSetDevice(prev);
    return B200_ENOMEM;
  }
  B200_KERNEL_LAUNCH(k_fp_inv_table_init, 1, 32, 0, c->stream, c->inv_pow2);
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) {
    b200_ctx_destroy(c);
    cudaSetDevice(prev);
    return B200_ECUDA;
  }
  cudaSetDevice(prev);
  *out = c;
  return B200_OK;
}
void b200_ctx_destroy(b200_ctx *ctx) {
  if (!ctx) return;
  if (ctx->nccl_comm || ctx->comm_buf) b200_ctx_comm_destroy(ctx);
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(ctx->device);
  if (ctx->stream) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);  // a caller's stream (b200_ctx_create_on_stream) stays the caller's
  }
  if (ctx->stream2) {
    cudaStreamSynchronize(ctx->stream2);
    cudaStreamDestroy(ctx->stream2);
  }
  if (ctx->stream3) {
    cudaStreamSynchronize(ctx->stream3);
    cudaStreamDestroy(ctx->stream3);
  }
  for (int i = 0; i < b200_ctx::N_SYNC_EVENTS; i++)
    if (ctx->ev_sync[i]) cudaEventDestroy(ctx->ev_sync[i]);
  if (ctx->inv_pow2) cudaFree(ctx->inv_pow2);
  if (ctx->fr_state && ctx->fr_state_free) ctx->fr_state_free(ctx->fr_state);
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->stage) cudaFree(ctx->stage);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  cudaSetDevice(prev);
  delete ctx;
}
const char *b200_strerror(int code) {
  switch (code) {
    case B200_OK: return "ok";
    case B200_EINVAL: return "invalid argument";
    case B200_ENODEV: return "no usable CUDA device";
    case B200_ECUDA: return "CUDA runtime error";
    case B200_ENOMEM: return "out of memory";
    case B200_ENCCL: return "NCCL unavailable or failed";
    default: return "unknown error";
  }
}
const char *b200_last_error(const b200_ctx *ctx) { return ctx ? ctx->err : ""; }
int b200_ctx_device(const b200_ctx *ctx) { return ctx ? ctx->device : -1; }
void *b200_ctx_stream(const b200_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t b200_ctx_launch_count(const b200_ctx *ctx) {
  if (!ctx) return 0;
  std::lock_guard<std::mutex> g(const_cast<b200_ctx *>(ctx)->mu);
  return ctx->launches;
}
int b200_ctx_set_msm_window(b200_ctx *ctx, int c) {
  if (!ctx) return B200_EINVAL;
  if (c != 0 && (c < 2 || c > 24)) return B200_EINVAL;
  std::lock_guard<std::mutex> g(ctx->mu);   // serialised with running calls, like b200_ctx_set_tuning
  ctx->msm_c = c;
  return B200_OK;
}

int b200_ctx_set_tuning(b200_ctx *ctx, const char *key, int value) {
  if (!ctx || !key) return B200_EINVAL;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!strcmp(key, "msm_window")) {
    if (value != 0 && (value < 2 || value > 24)) return B200_EINVAL;
    ctx->msm_c = value;
  } else if (!strcmp(key, "g2_acc_blocks")) {
    if (value < 2 || value > 4) return B200_EINVAL;  // 2: registers; 3: shared memory, 3 blocks/SM; 4: shared, 2 blocks
    ctx->tune_g2_acc_blocks = value;
  } else if (!strcmp(key, "msm_affine_levels")) {
    if (value < -1 || value > 3) return B200_EINVAL;
    ctx->tune_msm_affine_levels = value;
  } else if (!strcmp(key, "g1_glv")) {
    if (value < 0 || value > 2) return B200_EINVAL;
    ctx->tune_g1_glv = value;
  } else if (!strcmp(key, "msm_tail_groups")) {
    ctx->tune_msm_tail_groups = value != 0;
  } else if (!strcmp(key, "msm_reduce_min_chunk")) {
    if (value < 1 || value > 64) return B200_EINVAL;
    ctx->tune_msm_reduce_min_chunk = value;
  } else if (!strcmp(key, "msm_reduce")) {
    if (value < -1 || value > 2) return B200_EINVAL;
    ctx->tune_msm_reduce = value;
  } else if (!strcmp(key, "g1_prefetch")) {
    ctx->tune_g1_prefetch = value != 0;
  } else if (!strcmp(key, "pairing_variant")) {
    if (value != 0 && value != 4 && value != 7) return B200_EINVAL;  // 0 = by batch size, 4 = one thread per pairing, 7 = six lanes
    ctx->tune_pairing_variant = value;
  } else if (!strcmp(key, "coop_split")) {
    ctx->tune_coop_split = value != 0;
  } else if (!strcmp(key, "coop_prepare_max")) {
    if (value < 0) return B200_EINVAL;
    ctx->tune_coop_prepare_max = value;
  } else if (!strcmp(key, "coop_chunks")) {
    if (value < 1 || value > 64) return B200_EINVAL;
    ctx->tune_coop_chunks = value;
  } else if (!strcmp(key, "coop_warps")) {
    if (value < 1 || value > 12) return B200_EINVAL;
    ctx->tune_coop_warps = value;
  } else if (!strcmp(key, "mul_groups")) {
    if (value < -1 || value > 6) return B200_EINVAL;
    ctx->tune_mul_groups = value;
  } else if (!strcmp(key, "mul_groups_max_n")) {
    if (value < 0) return B200_EINVAL;
    ctx->tune_mul_groups_max_n = value;
  } else if (!strcmp(key, "pairing_chunks")) {
    if (value < 1 || value > 64) return B200_EINVAL;
    ctx->tune_pairing_chunks = value;
  } else {
    return B200_EINVAL;
  }
  return B200_OK;
}
int b200_ctx_set_timing(b200_ctx *ctx, int on) {
  if (!ctx) return B200_EINVAL;
  std::lock_guard<std::mutex> g(ctx->mu);
  ctx->timing = on != 0;
  ctx->ev_names.clear();
  return B200_OK;
}
int b200_ctx_get_timing(b200_ctx *ctx, char *names, size_t names_len, float *ms, int max) {
  CHECK_CTX(ctx);
  if (max < 0 || (max && !ms)) return B200_EINVAL;
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int nrec = (int)ctx->ev_names.size();
  size_t off = 0;
  if (names && names_len) names[0] = 0;
  for (int r = 0; r < nrec && r < max; r++) {
    float t = 0;
    if (cudaEventElapsedTime(&t, ctx->ev_pool[2 * r], ctx->ev_pool[2 * r + 1]) != cudaSuccess) t = -1.f;
    ms[r] = t;
    if (names) {
      size_t l = strlen(ctx->ev_names[r]);
      if (off + l + 2 <= names_len) {
        memcpy(names + off, ctx->ev_names[r], l);
        off += l;
        names[off++] = '\n';
        names[off] = 0;
      }
    }
  }
  return nrec;
}

// ================================================================ field tower
int b200_tower_op(b200_ctx *ctx, int level, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
  CHECK_CTX(ctx);
  if (!a || !out) return B200_EINVAL;
  if (level != 1 && level != 2 && level != 6 && level != 12) return B200_EINVAL;
  bool binary = op == B200_OP_MUL || op == B200_OP_ADD || op == B200_OP_SUB;
  if (binary && !b) return B200_EINVAL;
  static const int ok1[] = {0, 1, 2, 3, 4, 5, 10}, ok2[] = {0, 1, 2, 3, 4, 5, 6, 7, 8}, ok6[] = {0, 1, 2, 3, 4, 5, 6, 8},
                   ok12[] = {0, 3, 5, 6, 7, 9};
  const int *okl = level == 1 ? ok1 : level == 2 ? ok2 : level == 6 ? ok6 : ok12;
  int nok = level == 1 ? 7 : level == 2 ? 9 : level == 6 ? 8 : 6;
  bool found = false;
  for (int i = 0; i < nok; i++) found |= okl[i] == op;
  if (!found) return B200_EINVAL;
  if (n == 0) return B200_OK;
  size_t bytes = (size_t)48 * level * n;
  stager st(ctx, 3 * bytes);
  const uint64_t *da = (const uint64_t *)st.in(a, bytes), *db = binary ? (const uint64_t *)st.in(b, bytes) : nullptr;
  uint64_t *dout = (uint64_t *)st.out(bytes);
  if (st.rc != B200_OK) return st.rc;
  unsigned block = level >= 6 ? 64 : 128;
  switch (level) {
    case 1: B200_LAUNCH(ctx, k_tower_op<1>, nblk(n, block), block, 0, op, da, db, dout, n, ctx->inv_pow2); break;
    case 2: B200_LAUNCH(ctx, k_tower_op<2>, nblk(n, block), block, 0, op, da, db, dout, n, ctx->inv_pow2); break;
    case 6: B200_LAUNCH(ctx, k_tower_op<6>, nblk(n, block), block, 0, op, da, db, dout, n, ctx->inv_pow2); break;
    default: B200_LAUNCH(ctx, k_tower_op<12>, nblk(n, block), block, 0, op, da, db, dout, n, ctx->inv_pow2); break;
  }
  st.back(out, dout, bytes);
  return st.sync();
}

// ================================================================ group ops (element-wise)
#define GROUP_API(G, F, AFF, PROJ)                                                                                      \
  int b200_##G##_double_batch(b200_ctx *ctx, const PROJ *p, size_t n, PROJ *out) {                                      \
    CHECK_CTX(ctx);                                                                                                     \
    if (!p || !out) return B200_EINVAL;                                                                                 \
    if (n == 0) return B200_OK;                                                                                         \
    size_t pb = sizeof(PROJ) * n;                                                                                       \
    stager st(ctx, 2 * pb);                                                                                             \
    const char *dp = (const char *)st.in(p, pb);                                                                        \
    char *dout = (char *)st.out(pb);                                                                                    \
    if (st.rc != B200_OK) return st.rc;                                                                                 \
    B200_LAUNCH(ctx, k_double<F>, nblk(n, 128), 128, 0, dp, dout, n);                                                   \
    st.back(out, dout, pb);                                                                                             \
    return st.sync();                                                                                                   \
  }                                                                                                                     \
  int b200_##G##_add_batch(b200_ctx *ctx, const PROJ *p, const PROJ *q, size_t n, PROJ *out) {                          \
    CHECK_CTX(ctx);                                                                                                     \
    if (!p || !q || !out) return B200_EINVAL;                                                                           \
    if (n == 0) return B200_OK;                                                                                         \
    size_t pb = sizeof(PROJ) * n;                                                                                       \
    stager st(ctx, 3 * pb);                                                                                             \
    const char *dp = (const char *)st.in(p, pb), *dq = (const char *)st.in(q, pb);                                      \
    char *dout = (char *)st.out(pb);                                                                                    \
    if (st.rc != B200_OK) return st.rc;                                                                                 \
    B200_LAUNCH(ctx, k_add<F>, nblk(n, 128), 128, 0, dp, dq, dout, n);                                                  \
    st.back(out, dout, pb);                                                                                             \
    return st.sync();                                                                                                   \
  }                                                                                                                     \
  int b200_##G##_add_mixed_batch(b200_ctx *ctx, const PROJ *p, const AFF *q, const uint8_t *q_inf, size_t n,            \
                                 PROJ *out) {                                                                           \
    CHECK_CTX(ctx);                                                                                                     \
    if (!p || !q || !out) return B200_EINVAL;                                                                           \
    if (n == 0) return B200_OK;                                                                                         \
    size_t pb = sizeof(PROJ) * n, ab = sizeof(AFF) * n;                                                                 \
    stager st(ctx, 2 * pb + ab + n);                                                                                    \
    const char *dp = (const char *)st.in(p, pb), *dq = (const char *)st.in(q, ab);                                      \
    const uint8_t *di = (const uint8_t *)st.in(q_inf, n);                                                               \
    char *dout = (char *)st.out(pb);                                                                                    \
    if (st.rc != B200_OK) return st.rc;                                                                                 \
    B200_LAUNCH(ctx, k_add_mixed<F>, nblk(n, 128), 128, 0, dp, dq, di, dout, n);                                        \
    st.back(out, dout, pb);                                                                                             \
    return st.sync();                                                                                                   \
  }                                                                                                                     \
  int b200_##G##_mul_batch_dev(b200_ctx *ctx, const void *p, const void *s, size_t n, void *out) {                      \
    CHECK_CTX(ctx);                                                                                                     \
    if (n && (!p || !s || !out)) return B200_EINVAL;                                                                    \
    int rc = mul_batch_dev<F>(ctx, p, s, n, out);                                                                       \
    if (rc != B200_OK) return rc;                                                                                       \
    B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));                                                                 \
    return B200_OK;                                                                                                     \
  }                                                                                                                     \
  int b200_##G##_mul_batch(b200_ctx *ctx, const PROJ *p, const b200_scalar *s, size_t n, PROJ *out) {                   \
    CHECK_CTX(ctx);                                                                                                     \
    if (n && (!p || !s || !out)) return B200_EINVAL;                                                                    \
    if (n == 0) return B200_OK;                                                                                         \
    size_t pb = sizeof(PROJ) * n;                                                                                       \
    stager st(ctx, 2 * pb + 32 * n);                                                                                    \
    const void *dp = st.in(p, pb), *ds = st.in(s, 32 * n);                                                              \
    void *dout = st.out(pb);                                                                                            \
    if (st.rc != B200_OK) return st.rc;                                                                                 \
    int rc = mul_batch_dev<F>(ctx, dp, ds, n, dout);                                                                    \
    if (rc != B200_OK) return rc;                                                                                       \
    st.back(out, dout, pb);                                                                                             \
    return st.sync();                                                                                                   \
  }                                                                                                                     \
  int b200_##G##_batch_normalize_dev(b200_ctx *ctx, const void *p, size_t n, void *out_xy, void *out_inf) {             \
    CHECK_CTX(ctx);                                                                                                     \
    if (n && (!p || !out_xy || !out_inf)) return B200_EINVAL;                                                           \
    int rc = batch_normalize_dev<F>(ctx, p, n, out_xy, out_inf);                                                        \
    if (rc != B200_OK) return rc;                                                                                       \
    B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));                                                                 \
    return B200_OK;                                                                                                     \
  }                                                                                                                     \
  int b200_##G##_batch_normalize(b200_ctx *ctx, const PROJ *p, size_t n, AFF *out, uint8_t *out_inf) {                  \
    CHECK_CTX(ctx);                                                                                                     \
    if (n && (!p || !out || !out_inf)) return B200_EINVAL;                                                              \
    if (n == 0) return B200_OK;                                                                                         \
    size_t pb = sizeof(PROJ) * n, ab = sizeof(AFF) * n;                                                                 \
    stager st(ctx, pb + ab + n);                                                                                        \
    const void *dp = st.in(p, pb);                                                                                      \
    void *dxy = st.out(ab), *di = st.out(n);                                                                            \
    if (st.rc != B200_OK) return st.rc;                                                                                 \
    int rc = batch_normalize_dev<F>(ctx, dp, n, dxy, di);                                                               \
    if (rc != B200_OK) return rc;                                                                                       \
    st.back(out, dxy, ab);                                                                                              \
    st.back(out_inf, di, n);                                                                                            \
    return st.sync();                                                                                                   \
  }                                                                                                                     \
  int b200_##G##_sum_dev(b200_ctx *ctx, const void *parts, size_t n, void *out) {                                       \
    CHECK_CTX(ctx);                                                                                                     \
    if (!out || (n && !parts)) return B200_EINVAL;                                                                      \
    int rc = sum_dev<F>(ctx, parts, n, out);                                                                            \
    if (rc != B200_OK) return rc;                                                                                       \
    B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));                                                                 \
    return B200_OK;                                                                                                     \
  }

GROUP_API(g1, fp, b200_g1_affine, b200_g1_projective)
GROUP_API(g2, fp2, b200_g2_affine, b200_g2_projective)

// ================================================================ measurement helper
int b200_imad_peak(b200_ctx *ctx, int iters, double *imad_per_sec, double *ms_out) {
  return b200_imad_peak_mode(ctx, iters, 0, imad_per_sec, ms_out);
}
int b200_imad_peak_mode(b200_ctx *ctx, int iters, int mode, double *imad_per_sec, double *ms_out) {
  CHECK_CTX(ctx);
  if (iters <= 0 || !imad_per_sec || mode < 0 || mode > 1) return B200_EINVAL;
  int rc = arena_reserve(ctx, 256);
  if (rc != B200_OK) return rc;
  uint64_t *sink = arena_take<uint64_t>(ctx, 1);
  cudaEvent_t e0, e1;
  B200_CUDA(ctx, cudaEventCreate(&e0));
  B200_CUDA(ctx, cudaEventCreate(&e1));
  unsigned grid = (unsigned)ctx->sm_count * 8, block = 256;
  if (mode == 0) {
    B200_LAUNCH(ctx, k_imad_peak, grid, block, 0, iters / 8 + 1, sink);  // warm-up
  } else {
    B200_LAUNCH(ctx, k_imad_peak_lohi, grid, block, 0, iters / 8 + 1, sink);
  }
  B200_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  if (mode == 0) {
    B200_LAUNCH(ctx, k_imad_peak, grid, block, 0, iters, sink);
  } else {
    B200_LAUNCH(ctx, k_imad_peak_lohi, grid, block, 0, iters, sink);
  }
  B200_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  B200_CUDA(ctx, cudaEventSynchronize(e1));
  float ms = 0;
  B200_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  double total = (double)grid * block * (double)iters * 16.0 * 8.0;
  *imad_per_sec = total / (ms * 1e-3);
  if (ms_out) *ms_out = ms;
  return B200_OK;
}

}  // extern "C"
