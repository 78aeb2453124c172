// C ABI, part 3: Pippenger bucket multi-scalar multiplication for G1 and G2 (BASELINE configs 2, 3, 5).
//
// The reference has no MSM: it can only express sum_i p_i * s_i as N constant-time 255-step
// double-and-adds folded with `+` (src/g1.rs:573-579, :754-774, :161-171).  This file
// computes the same group element with the bucket method, reorganised for a B200:
//
//   1. k_msm_count    signed c-bit window digits of every scalar -> per-(window,bucket) histogram
//   2. k_msm_scan     exclusive scan of the histogram (one block per window)
//   3. k_msm_scatter  counting-sort scatter: point index (+ sign bit) grouped by bucket
//   4. k_msm_accumulate  one thread per bucket: XYZZ mixed additions (8M+2S) of its points with explicit
//                     P+P / P+(-P) / empty-accumulator handling (the 126 MB L2 holds the whole 96/192 MB
//                     point set, so the gathers are L2 hits); the bucket is stored in the reference's
//                     projective form
//   5. k_msm_reduce   sum_b (b+1) * B_b per window: per-thread running sums over a chunk of buckets,
//                     chunk offset by a short double-and-add, shared-memory tree per block
//   6. k_msm_horner   sums the per-block partials and runs Horner over the windows, group by group from
//                     the top on a side stream, overlapped with steps 4-5 of the lower windows
//
// Steps 5-6 use the reference's COMPLETE formulas (curve.cuh); step 4's exceptional cases (duplicate points,
// P + (-P)) are branched on explicitly; identity inputs and zero digits never reach a bucket.  Window sharding (shard, n_shards) restricts
// steps 1-5 to windows w = shard (mod n_shards); step 6 then yields sum_{w in shard} 2^(cw) S_w.
// Fp2 multiply of this unit: lazy reduction with row-alternated products / reductions (fp2.cuh; measured best for the
// G2 bucket kernel in round 2: 30.1 vs 31.3 ms at 2^20)
#define B200_FP2_LAZY3 1
#include <cmath>
#include "ctx.cuh"
#include "curve.cuh"
#include "curve_warp.cuh"
#include "glv.cuh"
#include "msm_affine.cuh"

using namespace b200;

int b200i_msm_check(b200_ctx *ctx);

namespace {

constexpr int MAX_WINDOWS = 128;

struct msm_plan {
  int c;          // window bits
  int nwin;       // total windows = ceil(256 / c)
  int nloc;       // windows handled by this shard
  int nbuckets;   // 2^(c-1) per window
  int glv;        // G1 only: scalars are split k = k1 + k2*lambda, entries refer to P (half 0) or phi(P) (half 1)
  int win[MAX_WINDOWS];  // global index of local window j
};
// entry of a bucket's point list: bits 0..29 point index, bit 30 = phi(P) instead of P (GLV), bit 31 = negate
constexpr uint32_t ENT_NEG = 0x80000000u, ENT_PHI = 0x40000000u, ENT_IDX = 0x3fffffffu;

// signed digit of window w for a canonical 256-bit little-endian scalar held in 8 words.
// digits d_w in [-2^(c-1), 2^(c-1)], sum_w d_w 2^(cw) == s.  Returns magnitude and sign.
__device__ __forceinline__ uint32_t window_bits(const uint32_t s[8], int lo, int c) {
  // bits [lo, lo+c) of s, zero beyond bit 255
  int word = lo >> 5, sh = lo & 31;
  uint64_t v = 0;
  if (word < 8) v = s[word];
  if (word + 1 < 8) v |= (uint64_t)s[word + 1] << 32;
  return (uint32_t)(v >> sh) & ((1u << c) - 1u);
}

__device__ __forceinline__ void load_scalar(uint32_t s[8], const uint32_t *scalars, size_t i) {
  const uint4 *sp = reinterpret_cast<const uint4 *>(scalars + 8 * i);
  uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
  s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w;
  s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
}

// Walks the windows of one scalar; calls f(local_window_index, bucket (0-based magnitude-1), flags) for every
// non-zero signed digit whose window belongs to this shard.  flags = ENT_NEG and/or ENT_PHI.
template <class Fn>
__device__ __forceinline__ void walk_digits(const msm_plan &pl, const uint32_t v[8], uint32_t base_flags, Fn f) {
  uint32_t carry = 0;
  int j = 0;
  const uint32_t half = 1u << (pl.c - 1);
  for (int w = 0; w < pl.nwin; w++) {
    uint32_t d = window_bits(v, w * pl.c, pl.c) + carry;
    bool neg = d > half;
    carry = neg ? 1u : 0u;
    uint32_t mag = neg ? (1u << pl.c) - d : d;
    if (j < pl.nloc && pl.win[j] == w) {
      if (mag != 0) f(j, mag - 1u, base_flags ^ (neg ? ENT_NEG : 0u));
      j++;
    }
  }
}
template <class Fn>
__device__ __forceinline__ void for_each_digit(const msm_plan &pl, const uint32_t s[8], Fn f) {
  if (!pl.glv) {
    walk_digits(pl, s, 0u, f);
    return;
  }
  glv_parts g = glv_decompose(s);
  uint32_t v[8] = {g.k1[0], g.k1[1], g.k1[2], g.k1[3], 0u, 0u, 0u, 0u};
  walk_digits(pl, v, g.neg1 ? ENT_NEG : 0u, f);
  v[0] = g.k2[0]; v[1] = g.k2[1]; v[2] = g.k2[2]; v[3] = g.k2[3];
  walk_digits(pl, v, ENT_PHI | (g.neg2 ? ENT_NEG : 0u), f);
}

// q (src/scalar.rs:76-81), 32-bit little-endian words: the ABI takes Scalar::to_bytes(), i.e. canonical scalars < q
__device__ __constant__ const uint32_t FR_Q[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                                  0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
__device__ __forceinline__ bool scalar_ge_q(const uint32_t s[8]) {
  for (int k = 7; k >= 0; k--) {
    if (s[k] != FR_Q[k]) return s[k] > FR_Q[k];
  }
  return true;
}

// `bad` (first window group only): set when a scalar is >= q.  The signed-window recoding and the GLV split assume
// canonical input (the carry out of the top window is dropped when c divides 256); instead of returning a wrong point for raw
// 32-byte strings the call fails with B200_EINVAL.
__global__ void __launch_bounds__(256) k_msm_count(msm_plan pl, const uint32_t *scalars, const uint8_t *inf, size_t n,
                                                 uint32_t *hist, uint32_t *bad) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (inf && inf[i]) return;
  uint32_t s[8];
  load_scalar(s, scalars, i);
  if (bad != nullptr && scalar_ge_q(s)) *bad = 1u;
  for_each_digit(pl, s, [&](int j, uint32_t b, uint32_t) { atomicAdd(&hist[(size_t)j * pl.nbuckets + b], 1u); });
}

// one block per local window: offsets[b] = exclusive prefix of hist over buckets; cursor := 0
__global__ void __launch_bounds__(1024) k_msm_scan(int nbuckets, const uint32_t *hist, uint32_t *offsets) {
  __shared__ uint32_t part[1024];
  const uint32_t *h = hist + (size_t)blockIdx.x * nbuckets;
  uint32_t *o = offsets + (size_t)blockIdx.x * nbuckets;
  int per = (nbuckets + blockDim.x - 1) / blockDim.x;
  int lo = threadIdx.x * per, hi = min(lo + per, nbuckets);
  uint32_t sum = 0;
  for (int b = lo; b < hi; b++) sum += h[b];
  part[threadIdx.x] = sum;
  __syncthreads();
  // Hillis–Steele inclusive scan over the 1024 partials
  for (int off = 1; off < (int)blockDim.x; off <<= 1) {
    uint32_t v = threadIdx.x >= (unsigned)off ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - sum;
  for (int b = lo; b < hi; b++) {
    uint32_t c = h[b];
    o[b] = run;
    run += c;
  }
}

__global__ void __launch_bounds__(256) k_msm_scatter(msm_plan pl, const uint32_t *scalars, const uint8_t *inf, size_t n,
                                                   size_t sstride, const uint32_t *offsets, uint32_t *cursor,
                                                   uint32_t *sorted) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (inf && inf[i]) return;
  uint32_t s[8];
  load_scalar(s, scalars, i);
  for_each_digit(pl, s, [&](int j, uint32_t b, uint32_t flags) {
    size_t k = (size_t)j * pl.nbuckets + b;
    uint32_t pos = offsets[k] + atomicAdd(&cursor[k], 1u);
    sorted[(size_t)j * sstride + pos] = (uint32_t)i | flags;
  });
}

// (Round 2 tried the north-star's staging of this sort: per-block histograms and cursors in SHARED memory — one block per
// (window, slice of the scalars), 2^15 counters = 128 KB, slice rows + column sums instead of global atomics (commit 78b2ab7 "MSM:
// counting sort through per-block shared-memory histograms").  Measured on B200, isolated kernel times (ncu, 2^20 scalars, 16
// windows): global atomics count 0.21 + scan 0.25 + scatter 0.29 = 0.76 ms per MSM — the L2 retires ~110 G atomics/s — against
// 0.54 + 0.61 (column sums) + 0.25 + 0.60 = 2.0 ms; whole step 9.37 vs 8.50 ms.  Removed; records in profiles/.)
// ---- bucket scheduling: order (window,bucket) slots by population, largest first, so the 32 lanes of
// a warp walk buckets of (nearly) equal length (round-1 ncu: 23/32 active lanes with natural order) and
// the last, partially filled wave holds only the short buckets.  Counting sort on min(count, 1023).
constexpr int SIZE_BINS = 1024;
// Buckets with >= GIANT_BUCKET points (skewed or adversarial scalars: e.g. all-equal scalars put all n points of a
// window into ONE bucket) would serialise on one thread.  They land in the last size bin, i.e. at the FRONT of
// `order`; each is split into GIANT_PARTS ranges summed by whole blocks and then combined.
constexpr uint32_t GIANT_BUCKET = SIZE_BINS - 1;
constexpr int GIANT_PARTS = 32;
// scheduling key of a slot: giants (by their level-0 population) in the last bin, everything else by the number
// of entries the bucket kernel will walk (after the affine levels, if any)
__device__ __forceinline__ uint32_t size_key(const uint32_t *counts, const uint32_t *hist0, size_t k) {
  return hist0[k] >= (uint32_t)SIZE_BINS - 1 ? (uint32_t)SIZE_BINS - 1 : min(counts[k], (uint32_t)SIZE_BINS - 2);
}
__global__ void __launch_bounds__(256) k_msm_size_hist(size_t total, const uint32_t *counts, const uint32_t *hist0,
                                                     uint32_t *size_hist) {
  __shared__ uint32_t sh[SIZE_BINS];
  for (int i = threadIdx.x; i < SIZE_BINS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k < total) atomicAdd(&sh[size_key(counts, hist0, k)], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < SIZE_BINS; i += blockDim.x)
    if (sh[i]) atomicAdd(&size_hist[i], sh[i]);
}
// single block: base[bin] = number of slots with a LARGER bin (descending order)
__global__ void __launch_bounds__(SIZE_BINS) k_msm_size_scan(const uint32_t *size_hist, uint32_t *size_base) {
  __shared__ uint32_t part[SIZE_BINS];
  int r = SIZE_BINS - 1 - threadIdx.x;  // thread 0 owns the largest bin
  uint32_t v = size_hist[r];
  part[threadIdx.x] = v;
  __syncthreads();
  for (int off = 1; off < SIZE_BINS; off <<= 1) {
    uint32_t a = threadIdx.x >= (unsigned)off ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += a;
    __syncthreads();
  }
  size_base[r] = part[threadIdx.x] - v;
}
// warp-aggregated scatter: lanes that hit the same bin are found with match.any, one atomic per group
__global__ void __launch_bounds__(256) k_msm_size_scatter(size_t total, const uint32_t *counts, const uint32_t *hist0,
                                                        const uint32_t *size_base, uint32_t *size_cursor, uint32_t *order) {
  size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k >= total) return;
  uint32_t bin = size_key(counts, hist0, k);
  unsigned lane = threadIdx.x & 31u;
  unsigned peers = __match_any_sync(__activemask(), bin);
  int leader = __ffs(peers) - 1;
  uint32_t rank = __popc(peers & ((1u << lane) - 1u));
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(&size_cursor[bin], (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  order[size_base[bin] + base + rank] = (uint32_t)k;
}

// one thread per (local window, bucket), visited in `order`
template <class F, int MINB, bool PREFETCH>
__global__ void __launch_bounds__(128, MINB) k_msm_accumulate(int nbuckets, size_t total, size_t slot0, const char *points,
                                                      const char *bx, size_t sstride, const uint32_t *offsets,
                                                      const uint32_t *hist, const uint32_t *sorted,
                                                      const uint32_t *order, char *buckets) {
  size_t t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t0 >= total) return;
  size_t k = slot0 + order[t0];   // `order` holds slot indices relative to the group's first slot
  constexpr size_t FB = field_traits<F>::bytes, AB = 2 * FB, PB = 3 * FB;
  size_t j = k / nbuckets;
  const uint32_t *idx = sorted + j * sstride + offsets[k];
  uint32_t cnt = hist[k];
  if (cnt >= GIANT_BUCKET) return;  // split across blocks by k_msm_giant_parts / k_msm_giant_final
  xyzz<F> acc = xyzz_identity<F>();
  if constexpr (PREFETCH) {
    // software pipeline: the NEXT point of the bucket is copied global -> shared with cp.async (LDGSTS, no
    // registers held across the ~3000-instruction addition) while the current addition runs.  Double buffer,
    // 16-byte chunks interleaved across the block's threads (conflict-free LDS.128).
    B200_DYN_SMEM(uint4, pf);
    constexpr int NCH = (int)(AB / 16);
    auto issue = [&](int buf, uint32_t e) {
      const char *src = points + AB * (size_t)(e & ENT_IDX);
      // GLV: phi(P) = (beta x, y): x comes from the precomputed beta*x array, y from the point
      const char *srcx = (e & ENT_PHI) ? bx + FB * (size_t)(e & ENT_IDX) : src;
#pragma unroll
      for (int c = 0; c < NCH; c++) {
        const char *g = c < NCH / 2 ? srcx + 16 * c : src + 16 * c;
#ifdef B200_HOST_EMUL  // CPU test harness: the copy is synchronous
        pf[(buf * NCH + c) * 128 + threadIdx.x] = *reinterpret_cast<const uint4 *>(g);
#else
        unsigned dst = (unsigned)__cvta_generic_to_shared(&pf[(buf * NCH + c) * 128 + threadIdx.x]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g) : "memory");
#endif
      }
#ifndef B200_HOST_EMUL
      asm volatile("cp.async.commit_group;" ::: "memory");
#endif
    };
    uint32_t e = cnt ? __ldg(idx) : 0u;
    if (cnt) issue(0, e);
    for (uint32_t t = 0; t < cnt; t++) {
      uint32_t e_next = 0;
      if (t + 1 < cnt) {
        e_next = __ldg(idx + t + 1);
        issue((t + 1) & 1, e_next);
#ifndef B200_HOST_EMUL
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
      }
      const uint4 *b = &pf[((t & 1) * NCH) * 128 + threadIdx.x];
      uint32_t w[AB / 4];
#pragma unroll
      for (int c = 0; c < NCH; c++) {
        uint4 v = b[c * 128];
        w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
      }
      F x, y;
      memcpy(&x, w, FB);
      memcpy(&y, w + FB / 4, FB);
      if (e >> 31) y = f_neg(y);
      acc = xyzz_add_mixed(acc, x, y);
      e = e_next;
    }
  } else {
    for (uint32_t t = 0; t < cnt; t++) {
      uint32_t e = __ldg(idx + t);
      const char *pp = points + AB * (size_t)(e & ENT_IDX);
      F x = field_traits<F>::load_ro((e & ENT_PHI) ? bx + FB * (size_t)(e & ENT_IDX) : pp), y = field_traits<F>::load_ro(pp + FB);
      if (e >> 31) y = f_neg(y);
      acc = xyzz_add_mixed(acc, x, y);
    }
  }
  proj_store<F>(buckets + PB * k, xyzz_to_proj(acc));
}

// ---- G2 variant of the bucket kernel: the XYZZ accumulator (4 x Fp2 = 96 words) lives in SHARED memory,
// word-interleaved across the 128 threads of the block (word w of thread t at sm[w*128 + t]: conflict-free),
// and the madd is ordered so that at most ~4 Fp2 values are live at once.  With the accumulator in registers
// the kernel needs 255 registers (2 blocks/SM, ~48 % of the IMAD pipe); this form is built for 3 blocks/SM.
struct sm_fp2 {
  uint32_t *p;  // &sm[component * 24 * 128 + tid]
  B200_DEV fp2 get() const {
    fp2 r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.c0.v[i] = p[i * 128];
#pragma unroll
    for (int i = 0; i < 12; i++) r.c1.v[i] = p[(12 + i) * 128];
    return r;
  }
  B200_DEV void set(const fp2 &a) const {
#pragma unroll
    for (int i = 0; i < 12; i++) p[i * 128] = a.c0.v[i];
#pragma unroll
    for (int i = 0; i < 12; i++) p[(12 + i) * 128] = a.c1.v[i];
  }
};
// rare path (same x): kept out of line so it does not inflate the hot loop's register allocation
static __device__ __noinline__ void g2sm_same_x(uint32_t *smt, const char *pt, bool negate, bool r_is_zero, bool *empty) {
  const sm_fp2 X{smt}, Y{smt + 24 * 128}, ZZ{smt + 48 * 128}, ZZZ{smt + 72 * 128};
  fp2 px = fp2_load_ro(pt), py = fp2_load_ro(pt + 96);
  if (negate) py = fp2_neg(py);
  if (!r_is_zero || fp2_is_zero(py)) {  // P + (-P)
    *empty = true;
    return;
  }
  xyzz<fp2> d = xyzz_add_mixed(xyzz<fp2>{X.get(), Y.get(), ZZ.get(), ZZZ.get()}, px, py);  // takes its P + P branch
  X.set(d.x);
  Y.set(d.y);
  ZZ.set(d.zz);
  ZZZ.set(d.zzz);
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_msm_accumulate_g2sm(int nbuckets, size_t total, size_t slot0, const char *points,
                                                                 size_t sstride, const uint32_t *offsets, const uint32_t *hist,
                                                                 const uint32_t *sorted, const uint32_t *order, char *buckets) {
  B200_DYN_SMEM(uint32_t, sm);
  size_t t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t0 >= total) return;
  size_t k = slot0 + order[t0];
  size_t j = k / nbuckets;
  const uint32_t *idx = sorted + j * sstride + offsets[k];
  uint32_t cnt = hist[k];
  if (cnt >= GIANT_BUCKET) return;  // handled by k_msm_giant_parts / k_msm_giant_final
  const sm_fp2 X{sm + threadIdx.x}, Y{sm + 24 * 128 + threadIdx.x}, ZZ{sm + 48 * 128 + threadIdx.x},
      ZZZ{sm + 72 * 128 + threadIdx.x};
  bool empty = true;
  for (uint32_t t = 0; t < cnt; t++) {
    uint32_t e = __ldg(idx + t);
    const char *pp_ = points + 192 * (size_t)(e & ENT_IDX);
    fp2 px = fp2_load_ro(pp_), py = fp2_load_ro(pp_ + 96);
    if (e >> 31) py = fp2_neg(py);
    if (empty) {
      X.set(px);
      Y.set(py);
      ZZ.set(fp2_one());
      ZZZ.set(fp2_one());
      empty = false;
      continue;
    }
    fp2 p = fp2_sub(M2(px, ZZ.get()), X.get());
    fp2 r = fp2_sub(M2(py, ZZZ.get()), Y.get());
    if (fp2_is_zero(p)) {  // same x: P + P or P + (-P)   (rare; see xyzz_add_mixed)
      g2sm_same_x(sm + threadIdx.x, pp_, (e >> 31) != 0, fp2_is_zero(r), &empty);
      continue;
    }
    fp2 pp = S2(p);
    ZZ.set(M2(ZZ.get(), pp));
    fp2 ppp = M2(p, pp);
    ZZZ.set(M2(ZZZ.get(), ppp));
    fp2 q = M2(X.get(), pp);
    fp2 x3 = fp2_sub(fp2_sub(S2(r), ppp), fp2_dbl(q));
    fp2 tt = M2(Y.get(), ppp);
    X.set(x3);
    Y.set(fp2_sub(M2(r, fp2_sub(q, x3)), tt));
  }
  proj<fp2> out = proj_identity<fp2>();
  if (!empty) {
    fp2 zz = ZZ.get(), zzz = ZZZ.get();
    out = proj<fp2>{M2(X.get(), zzz), M2(Y.get(), zz), M2(zz, zzz)};
  }
  proj_store<fp2>(buckets + (size_t)288 * k, out);
}

// block (b) works on giant g = b / GIANT_PARTS, part b % GIANT_PARTS; grid-stride over (giant, part) pairs.
// n_giant is read from the size histogram on the device (no host round trip); usually 0 -> immediate exit.
template <class F>
__global__ void __launch_bounds__(128) k_msm_giant_parts(int nbuckets, size_t slot0, const uint32_t *size_hist,
                                                       const uint32_t *order, const char *points, const char *bx,
                                                       size_t sstride, const uint32_t *offsets, const uint32_t *hist,
                                                       const uint32_t *sorted, char *gparts, uint32_t max_giants) {
  B200_DYN_SMEM(char, smem);
  constexpr size_t FB = field_traits<F>::bytes, AB = 2 * FB, PB = 3 * FB;
  uint32_t ng = min(size_hist[SIZE_BINS - 1], max_giants);
  for (uint32_t item = blockIdx.x; item < ng * GIANT_PARTS; item += gridDim.x) {
    uint32_t g = item / GIANT_PARTS, part = item % GIANT_PARTS;
    size_t k = slot0 + order[g];
    size_t j = k / nbuckets;
    const uint32_t *idx = sorted + j * sstride + offsets[k];
    uint32_t cnt = hist[k];
    uint32_t per = (cnt + GIANT_PARTS - 1) / GIANT_PARTS;
    uint32_t lo = part * per, hi = min(lo + per, cnt);
    xyzz<F> acc = xyzz_identity<F>();
    for (uint32_t t = lo + threadIdx.x; t < hi; t += blockDim.x) {
      uint32_t e = __ldg(idx + t);
      const char *pp = points + AB * (size_t)(e & ENT_IDX);
      F x = field_traits<F>::load_ro((e & ENT_PHI) ? bx + FB * (size_t)(e & ENT_IDX) : pp), y = field_traits<F>::load_ro(pp + FB);
      if (e >> 31) y = f_neg(y);
      acc = xyzz_add_mixed(acc, x, y);
    }
    proj_store<F>(smem + PB * threadIdx.x, xyzz_to_proj(acc));
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) {
        proj<F> a = proj_load<F>(smem + PB * threadIdx.x), b2 = proj_load<F>(smem + PB * (threadIdx.x + s));
        proj_store<F>(smem + PB * threadIdx.x, proj_add(a, b2));
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) proj_store<F>(gparts + PB * item, proj_load<F>(smem));
    __syncthreads();
  }
}
// one warp per giant: sum its GIANT_PARTS partial sums (lane-parallel complete additions), write the bucket
template <class F>
__global__ void __launch_bounds__(32) k_msm_giant_final(size_t slot0, const uint32_t *size_hist, const uint32_t *order,
                                                      const char *gparts, char *buckets, uint32_t max_giants) {
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  uint32_t ng = min(size_hist[SIZE_BINS - 1], max_giants);
  const int lane = threadIdx.x;
  for (uint32_t g = blockIdx.x; g < ng; g += gridDim.x) {
    proj<F> acc = proj_identity<F>();
#pragma unroll 1
    for (int p = 0; p < GIANT_PARTS; p++) acc = warp_add(acc, proj_load<F>(gparts + PB * ((size_t)g * GIANT_PARTS + p)), lane);
    if (lane == 0) proj_store<F>(buckets + PB * (slot0 + order[g]), acc);
  }
}

// GLV prologue: bx[i] = beta * x_i  (x-coordinate of phi(P_i))
__global__ void __launch_bounds__(256) k_msm_glv_bx(const char *points, size_t n, char *bx) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  fp_store(bx + 48 * i, fp_mul(fp_load_ro(points + 96 * i), fp_const(K_GLV_BETA)));
}
// test surface for the decomposition: out[i] = k1 (16 B LE) | k2 (16 B LE), sign[i] = neg1 | neg2 << 1
__global__ void k_glv_decompose(const uint32_t *scalars, size_t n, uint32_t *out, uint8_t *sign) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[8];
  load_scalar(s, scalars, i);
  glv_parts g = glv_decompose(s);
  for (int k = 0; k < 4; k++) {
    out[8 * i + k] = g.k1[k];
    out[8 * i + 4 + k] = g.k2[k];
  }
  sign[i] = (g.neg1 ? 1 : 0) | (g.neg2 ? 2 : 0);
}

// small multiple k * P, k < 2^24, by double-and-add (MSB first)
template <class F>
__device__ proj<F> proj_mul_small(const proj<F> &p, uint32_t k) {
  proj<F> acc = proj_identity<F>();
  if (k == 0) return acc;
  int top = 31 - __clz(k);
#pragma unroll 1
  for (int b = top; b >= 0; b--) {
    acc = proj_double(acc);
    if ((k >> b) & 1) acc = proj_add(acc, p);
  }
  return acc;
}

// grid = (blocks_per_window, nloc).  Thread handles `chunk` consecutive buckets.
template <class F, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_msm_reduce(int nbuckets, int chunk, const char *buckets, char *partials) {
  B200_DYN_SMEM(char, smem);
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  int j = blockIdx.y;
  int t = blockIdx.x * BLOCK + threadIdx.x;       // chunk index within the window
  int lo = t * chunk, hi = min(lo + chunk, nbuckets);
  const char *wb = buckets + PB * (size_t)j * nbuckets;
  proj<F> run = proj_identity<F>(), acc = proj_identity<F>();
  for (int b = hi - 1; b >= lo; b--) {
    run = proj_add(run, proj_load<F>(wb + PB * b));
    acc = proj_add(acc, run);
  }
  // acc = sum (b - lo + 1) B_b ; bucket b is worth (b + 1):  add lo * run
  if (lo > 0 && lo < nbuckets) acc = proj_add(acc, proj_mul_small(run, (uint32_t)lo));
  proj_store<F>(smem + PB * threadIdx.x, acc);
  __syncthreads();
  for (int s = BLOCK / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      proj<F> a = proj_load<F>(smem + PB * threadIdx.x), b2 = proj_load<F>(smem + PB * (threadIdx.x + s));
      proj_store<F>(smem + PB * threadIdx.x, proj_add(a, b2));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) proj_store<F>(partials + PB * ((size_t)j * gridDim.x + blockIdx.x), proj_load<F>(smem));
}

// ---- lane-cooperative bucket reduction (round 2).  k_msm_reduce above gives every THREAD a chunk of buckets: ~60 dependent
// complete additions of ~13 us each (a single thread cannot keep more than one multiplier busy) — 0.8 ms per window group,
// the exposed tail of every MSM and the limiter of multi-GPU strong scaling.  Here a GROUP of six lanes (curve_warp.cuh:
// the independent products of a formula level in different lanes, ~1.7 us per addition) owns a chunk, five groups per warp.
//   k_msm_reduce_coop   group g of window j: run = sum B_b, acc = sum (b - lo + 1) B_b over its chunk [lo, lo + chunk), then
//                       acc += lo * run by a fixed-length double-and-add (uniform control flow: full-mask shuffles)
//   k_msm_fold_coop     sums `per` consecutive partials per group, until few enough are left for the Horner kernel
template <class F, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_msm_reduce_coop(int nbuckets, int chunk, int groups_per_window, int lo_bits,
                                                               const char *buckets, char *partials) {
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int grp = lane / 6;
  const bool live = grp < 5;
  if (!live) grp = 0;
  const int sub = live ? lane - 6 * grp : lane - 30, base = live ? 6 * grp : 30;
  const int j = blockIdx.y;
  int gi = (blockIdx.x * WARPS + warp) * 5 + grp;
  const bool active = live && gi < groups_per_window;
  if (gi >= groups_per_window) gi = groups_per_window - 1;
  const int lo = gi * chunk;
  const char *wb = buckets + PB * (size_t)j * nbuckets;
  proj<F> run = proj_identity<F>(), acc = proj_identity<F>();
#pragma unroll 1
  for (int t = chunk - 1; t >= 0; t--) {
    int b = lo + t;
    proj<F> bk = b < nbuckets ? proj_load<F>(wb + PB * b) : proj_identity<F>();
    run = grp_add(run, bk, sub, base);
    acc = grp_add(acc, run, sub, base);
  }
  // bucket b is worth (b + 1): acc so far counts (b - lo + 1); add lo * run (MSB first, every group the same number of steps)
  proj<F> t = proj_identity<F>();
#pragma unroll 1
  for (int bit = lo_bits - 1; bit >= 0; bit--) {
    t = grp_double(t, sub, base);
    proj<F> t2 = grp_add(t, run, sub, base);
    t = proj_select(t, t2, ((lo >> bit) & 1) != 0);
  }
  acc = grp_add(acc, t, sub, base);
  if (active && sub == 0) proj_store<F>(partials + PB * ((size_t)j * groups_per_window + gi), acc);
}
template <class F, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_msm_fold_coop(int n_in, int per, int n_out, const char *in, char *out) {
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int grp = lane / 6;
  const bool live = grp < 5;
  if (!live) grp = 0;
  const int sub = live ? lane - 6 * grp : lane - 30, base = live ? 6 * grp : 30;
  const int j = blockIdx.y;
  int gi = (blockIdx.x * WARPS + warp) * 5 + grp;
  const bool active = live && gi < n_out;
  if (gi >= n_out) gi = n_out - 1;
  proj<F> acc = proj_identity<F>();
#pragma unroll 1
  for (int t = 0; t < per; t++) {
    int i = gi * per + t;
    proj<F> v = i < n_in ? proj_load<F>(in + PB * ((size_t)j * n_in + i)) : proj_identity<F>();
    acc = grp_add(acc, v, sub, base);
  }
  if (active && sub == 0) proj_store<F>(out + PB * ((size_t)j * n_out + gi), acc);
}

// Horner over the windows, one GROUP of windows per launch, top window first, on the ctx's side stream:
// while the (one-thread, latency-bound) chain  acc <- 2^(c*gap) * acc + S_w  of the upper windows runs,
// the bucket kernels of the lower windows keep all SMs busy on the main stream; only the last group's
// step is exposed.  Lanes j < cnt first sum the per-block partials of local window (j_top - j).
// `prev_w` = global index of the window processed last (-1: none yet); `final_shift` = doublings to apply after the group's
// lowest window: down to the next group's top window, or (last group) down to bit 0 of the shard's lowest window.
template <class F>
__global__ void __launch_bounds__(32) k_msm_horner(msm_plan pl, int j_top, int cnt, int prev_w, int final_shift,
                                                 int parts_per_window, const char *partials, char *hacc) {
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  const int lane = threadIdx.x;
  // ONE warp, lane-parallel group operations (curve_warp.cuh): every lane carries the same accumulator
  proj<F> acc = prev_w < 0 ? proj_identity<F>() : proj_load<F>(hacc);
#pragma unroll 1
  for (int i = 0; i < cnt; i++) {
    int j = j_top - i, w = pl.win[j];
    proj<F> sw = proj_identity<F>();
#pragma unroll 1
    for (int k = 0; k < parts_per_window; k++)
      sw = warp_add(sw, proj_load<F>(partials + PB * ((size_t)j * parts_per_window + k)), lane);
    // the doublings from the previous GROUP's lowest window down to this group's top window were already applied at the end
    // of the previous launch (tail_shift), while this group's buckets were still being accumulated and reduced
    if (prev_w >= 0 && i > 0) {
#pragma unroll 1
      for (int k = (prev_w - w) * pl.c; k > 0; k--) acc = warp_double(acc, lane);
    }
    acc = warp_add(acc, sw, lane);
    prev_w = w;
  }
#pragma unroll 1
  for (int k = final_shift; k > 0; k--) acc = warp_double(acc, lane);
  if (lane == 0) proj_store<F>(hacc, acc);
}

template <class F>
__global__ void k_store_identity(char *out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) proj_store<F>(out, proj_identity<F>());
}

inline unsigned nblk(size_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

// is window width c a good choice for n points?  One thread walks one bucket, so what matters is the LONGEST bucket that is
// not handed to the block-parallel giant path (>= GIANT_BUCKET points).  The scalars are uniform below q (255 bits), so only the
// TOP window is special: it holds the top tb = 255 - c*floor(255/c) bits, i.e. qt = q / 2^shift possible values, and its last two
// buckets are lighter than the rest (b = floor(qt) covers only the fractional part f of a value range, and a carry out of the
// window below is rarer there).  Expected populations, n scalars:   ordinary top bucket n / qt;   digit B = floor(qt):
// n (f (1 - pc) + 1/2) / qt;   digit B + 1: n f pc / qt   with pc = max(0, f - 1/2) / f.   Each of them must be either short
// (<= 4 x the average bucket of the other windows, or <= 96 points: ~1 ms on one thread, the size of the fixed costs) or safely
// above the giant threshold.  Measured before this rule (round 2, profiles/records/r02_g1_msm_small_windows.txt): 2^14 points,
// c = 11: top buckets of 4580, 4556, 4461 and 547 points — the 547-point one alone kept a thread busy for 5.5 ms (MSM 7.9 ms;
// c = 12: 2.7 ms); 2^13 points, c = 9: 11.3 ms (c = 13: 2.9 ms).
// tb = 0 (c divides 255 = 3 * 5 * 17) is the worst case: the signed-digit carry of the full top window opens one more
// window whose single bucket collects half of all points.
bool window_ok(int c, size_t n) {
  int tb = 255 - c * (255 / c);
  if (tb == 0) return false;
  const int shift = c * (255 / c);
  const double qt = ldexp((double)0x73eda753299d7d48ull, 192 - shift);   // q / 2^shift (src/scalar.rs:76-81, top 64 bits)
  const double B = floor(qt), f = qt - B, pc = f > 0.5 ? (f - 0.5) / f : 0.0;
  const double avg = (double)n / (double)((size_t)1 << (c - 1));
  const double lim = 4.0 * avg + 8.0 > 96.0 ? 4.0 * avg + 8.0 : 96.0;
  auto fine = [&](double p) { return p <= lim || p >= 1.1 * (double)GIANT_BUCKET; };
  const double generic = (double)n / qt, pB = (double)n * (f * (1.0 - pc) + 0.5) / qt, pB1 = (double)n * f * pc / qt;
  return fine(generic) && fine(B >= 1.0 ? pB : 0.0) && fine(pB1);
}
int auto_window(size_t n) {
  // minimise  W * n (bucket adds) + W * 2^(c-1) * ~3 (reduction) ; measured sweet spots on B200.  Wider windows first: for small
  // n the step is latency (longest bucket, Horner chain), and more, shorter buckets win (2^14: c = 12 2.7 ms, c = 9 4.8 ms)
  int lg = 0;
  while (((size_t)1 << (lg + 1)) <= n) lg++;
  int c = lg - 4;
  if (c < 4) c = 4;
  if (c > 16) c = 16;
  const int order[6] = {0, 1, 2, 3, 4, -1};
  for (int k = 0; k < 6; k++) {
    int cc = c + order[k];
    if (cc >= 4 && cc <= 18 && window_ok(cc, n)) return cc;
  }
  return c;
}

template <class F>
int msm_dev(b200_ctx *ctx, const void *points, const void *inf, const void *scalars, size_t n, int shard, int n_shards,
            void *out) {
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  if (n_shards < 1 || shard < 0 || shard >= n_shards) return B200_EINVAL;
  if (n >= ((size_t)1 << 30)) return B200_EINVAL;
  msm_plan pl;
  // GLV (G1 only): 2n entries with 127-bit sub-scalars -> ceil(128/c) windows instead of ceil(256/c)
  pl.glv = (sizeof(F) == sizeof(fp) && (ctx->tune_g1_glv == 1 || (ctx->tune_g1_glv == 2 && n_shards > 1))) ? 1 : 0;
  pl.c = ctx->msm_c ? ctx->msm_c : auto_window(pl.glv ? 2 * n : n);
  pl.nwin = ((pl.glv ? 128 : 256) + pl.c - 1) / pl.c;
  const size_t sstride = pl.glv ? 2 * n : n;  // entries per window
  if (pl.nwin > MAX_WINDOWS) return B200_EINVAL;
  pl.nbuckets = 1 << (pl.c - 1);
  pl.nloc = 0;
  for (int w = shard; w < pl.nwin; w += n_shards) pl.win[pl.nloc++] = w;
  ctx->msm_bad_flag = nullptr;
  if (n == 0 || pl.nloc == 0) {
    B200_LAUNCH(ctx, k_store_identity<F>, 1, 32, 0, (char *)out);
    return B200_OK;
  }
  size_t total = (size_t)pl.nloc * pl.nbuckets;
  // reduction geometry
  // block size of the thread-per-chunk reduction (smem: RB * PB).  128 threads = the register footprint of ONE bucket-kernel block
  // (G2: 128 x 255): a reduction block running under the next group's bucket kernel displaces exactly one of its blocks; with 64
  // threads (round 1) twice as many blocks each displaced one — measured 28.3 vs 23.7 ms for the G2 2^20 step (msm_reduce = 1)
  constexpr int RB = 128;
  // threads per window (power of two): at most 2048, at least `tune_msm_reduce_min_chunk` buckets each — every 128 threads are one
  // partial sum that k_msm_horner adds up serially (3 us each), so narrow windows (2^11 buckets: small MSMs) must not be cut into
  // 16 one-bucket-per-thread blocks
  int chunks = pl.nbuckets >= 2048 ? 2048 : pl.nbuckets;
  {
    int min_chunk = ctx->tune_msm_reduce_min_chunk < 1 ? 1 : ctx->tune_msm_reduce_min_chunk;
    while (chunks > RB && pl.nbuckets / chunks < min_chunk) chunks >>= 1;
  }
  if (chunks < RB) chunks = RB;
  int chunk = (pl.nbuckets + chunks - 1) / chunks;
  if (chunk < 1) chunk = 1;
  int blocks_per_window = chunks / RB;
  // lane-cooperative reduction (default): groups of six lanes own RCHUNK buckets; partial counts per level: G, G/16, ... <= 8
  // tune_msm_reduce: 0 never, 1 only where the reduction is EXPOSED (the last window group of a call), 2 every group,
  // -1 (default) = 1.  Measured at 2^20 on one GPU: G1 8.85 (0) / 8.51 (1) / 9.50 (2) ms; G2 29.0 (0) / 23.7 (1) / 24.5 (2) ms with
  // 128-thread blocks in the thread-per-chunk kernel (64-thread blocks: 28.3 ms in mode 1 — see RB above).  The cooperative
  // kernels finish sooner but do more total work and take more of the GPU away from the pipe-bound bucket kernel of the next group.
  const int reduce_mode = ctx->tune_msm_reduce >= 0 ? ctx->tune_msm_reduce : 1;
  const bool coop_reduce = reduce_mode != 0;
  constexpr int RCHUNK = 16, RFOLD = 16, RWARPS = 4;
  int r_chunk = pl.nbuckets < RCHUNK ? pl.nbuckets : RCHUNK, r_lo_bits = 0;
  while ((1 << r_lo_bits) < pl.nbuckets) r_lo_bits++;
  int r_n[6], r_levels = 0;
  r_n[r_levels++] = (pl.nbuckets + r_chunk - 1) / r_chunk;
  while (r_n[r_levels - 1] > 8 && r_levels < 6) {
    r_n[r_levels] = (r_n[r_levels - 1] + RFOLD - 1) / RFOLD;
    r_levels++;
  }
  size_t r_bytes = 0;
  for (int l = 0; l < r_levels; l++) r_bytes += arena_pad((size_t)pl.nloc * r_n[l] * PB);
  // at most (entries of the largest group) / GIANT_BUCKET giants can exist at once
  const uint32_t max_giants = (uint32_t)((size_t)pl.nloc * sstride / GIANT_BUCKET + pl.nloc);
  size_t need = 4 * arena_pad(total * 4) + 3 * arena_pad(4 * SIZE_BINS * 4) + arena_pad((size_t)pl.nloc * sstride * 4) +
                arena_pad(total * PB) + arena_pad((size_t)pl.nloc * blocks_per_window * PB) + arena_pad(PB) +
                (pl.glv ? arena_pad(48 * n) : 0) + arena_pad((size_t)max_giants * GIANT_PARTS * PB) + 4096 + 256 +
                (coop_reduce ? r_bytes : 0);
  if (ctx->tune_msm_affine_levels != 0) {  // upper bound of the affine-level scratch (largest group <= all local windows)
    need += 8 * arena_pad(total * 4) + 3 * 256;
    for (int l = 1; l <= 3; l++) need += arena_pad((size_t)pl.nloc * ((sstride >> l) + pl.nbuckets + 8) * 2 * field_traits<F>::bytes);
  }
  int rc = arena_reserve(ctx, need);
  if (rc != B200_OK) return rc;
  uint32_t *hist = arena_take<uint32_t>(ctx, total);
  uint32_t *cursor = arena_take<uint32_t>(ctx, total);
  uint32_t *offsets = arena_take<uint32_t>(ctx, total);
  uint32_t *size_hist = arena_take<uint32_t>(ctx, 4 * SIZE_BINS);   // [parity][hist | cursor]
  uint32_t *size_base = arena_take<uint32_t>(ctx, 2 * SIZE_BINS);   // [parity]
  uint32_t *order = arena_take<uint32_t>(ctx, total);
  uint32_t *sorted = arena_take<uint32_t>(ctx, (size_t)pl.nloc * sstride);
  char *buckets = arena_take<char>(ctx, total * PB);
  char *partials = arena_take<char>(ctx, (size_t)pl.nloc * blocks_per_window * PB);
  char *hacc = arena_take<char>(ctx, PB);
  char *r_buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (coop_reduce)
    for (int l = 0; l < r_levels; l++) r_buf[l] = arena_take<char>(ctx, (size_t)pl.nloc * r_n[l] * PB);
  char *bx = pl.glv ? arena_take<char>(ctx, 48 * n) : nullptr;
  char *gparts = arena_take<char>(ctx, (size_t)max_giants * GIANT_PARTS * PB);
  uint32_t *bad = arena_take<uint32_t>(ctx, 1);
  ctx->msm_bad_flag = bad;
  B200_CUDA(ctx, cudaMemsetAsync(bad, 0, sizeof(uint32_t), ctx->stream));
  auto wait_points = [&]() -> int {   // msm_host: the points arrive on stream3 (ev_sync[30]); nothing before this reads them
    if (ctx->msm_points_event_pending) {
      ctx->msm_points_event_pending = false;
      B200_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_sync[30], 0));
    }
    return B200_OK;
  };
  if (pl.glv) {
    int rcw = wait_points();
    if (rcw != B200_OK) return rcw;
    B200_LAUNCH(ctx, k_msm_glv_bx, nblk(n, 256), 256, 0, (const char *)points, n, bx);
  }
  // hist and cursor are adjacent: one memset
  B200_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)((char *)offsets - (char *)hist), ctx->stream));
  // Window groups, top-down.  A group = consecutive local windows = a contiguous range of slots.
  // Three streams: accumulate(g) on `stream`; reduce(g) on stream2 (after accumulate(g)); horner(g) on
  // stream3 (after reduce(g) and horner(g-1)).  reduce/horner are latency-bound, few-thread kernels: they
  // run under the next group's accumulate; only the last group's reduce + Horner step is exposed.
  // affine tree levels before the bucket kernel (msm_affine.cuh): by average bucket population
  int LV = ctx->tune_msm_affine_levels;
  if (LV < 0) {
    size_t avg = sstride / (size_t)pl.nbuckets;
    LV = avg >= 24 ? 3 : avg >= 12 ? 2 : avg >= 6 ? 1 : 0;
  }
  if (LV > 3) LV = 3;
  int gsz[16], ng = 0;
  if (LV > 0 && pl.nloc >= 6) {
    // fewer, larger groups so that the level kernels fill the GPU (16 -> 7 + 6 + 3)
    gsz[ng++] = (pl.nloc * 7 + 8) / 16;
    gsz[ng++] = ((pl.nloc - gsz[0]) * 2 + 2) / 3;
    gsz[ng++] = pl.nloc - gsz[0] - gsz[1];
  } else {
    int left = pl.nloc;
    if (left >= 12) {
      gsz[ng++] = (left * 3) / 8; left -= gsz[ng - 1];     // 16 -> 6
      gsz[ng++] = (left + 1) / 2; left -= gsz[ng - 1];     //       5
      gsz[ng++] = (left * 3 + 4) / 5; left -= gsz[ng - 1]; //       3
      gsz[ng++] = left;                                    //       2
    } else if (left >= 4) {
      gsz[ng++] = left - left / 2 - (left >= 6 ? 1 : 0); left -= gsz[ng - 1];
      if (left >= 3) { gsz[ng++] = left - 1; left = 1; }
      gsz[ng++] = left;
    } else if (ctx->tune_msm_tail_groups && left >= 2) {
      // one window per group: the top window's reduction and its long run of Horner doublings (pre-applied at the end of its
      // launch) overlap the bucket accumulation of the windows below — what a window shard of an 8-GPU MSM needs
      while (left > 0) { gsz[ng++] = 1; left--; }
    } else {
      gsz[ng++] = left;   // (round 1, slow reduction: splitting 2-3 windows into [n-1, 1] was slower: 3.2-4.0 vs 2.5-3.5 ms per shard)
    }
  }
  // scratch of the affine levels, sized for the largest group and reused by every group (main stream only)
  constexpr size_t ABY = 2 * field_traits<F>::bytes;
  uint32_t *lv_np = nullptr, *lv_po = nullptr, *lv_cnt[4] = {nullptr, nullptr, nullptr, nullptr},
           *lv_off[4] = {nullptr, nullptr, nullptr, nullptr};
  char *lv_buf[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t lv_cap[4] = {0, 0, 0, 0};
  if (LV > 0) {
    int maxg = 0;
    for (int g = 0; g < ng; g++) maxg = gsz[g] > maxg ? gsz[g] : maxg;
    size_t gt = (size_t)maxg * pl.nbuckets;
    size_t extra = 2 * arena_pad(gt * 4);
    for (int l = 1; l <= LV; l++) {
      lv_cap[l] = (sstride >> l) + pl.nbuckets + 8;
      extra += 2 * arena_pad(gt * 4) + arena_pad((size_t)maxg * lv_cap[l] * ABY);
    }
    // second, separately grown arena region: taken from the same arena (reserved above via `need`)
    if (ctx->arena_off + extra > ctx->arena_size) return B200_ENOMEM;
    lv_np = arena_take<uint32_t>(ctx, gt);
    lv_po = arena_take<uint32_t>(ctx, gt);
    for (int l = 1; l <= LV; l++) {
      lv_cnt[l] = arena_take<uint32_t>(ctx, gt);
      lv_off[l] = arena_take<uint32_t>(ctx, gt);
      lv_buf[l] = arena_take<char>(ctx, (size_t)maxg * lv_cap[l] * ABY);
    }
  }
  // whatever path leaves this function (including the error returns below), the main stream is re-joined with both side
  // streams, so the next call — ordered only against ctx->stream — can never overlap kernels still using the arena
  struct join_guard {
    b200_ctx *c;
    ~join_guard() {
      cudaEventRecord(c->ev_sync[b200_ctx::N_SYNC_EVENTS - 2], c->stream2);
      cudaStreamWaitEvent(c->stream, c->ev_sync[b200_ctx::N_SYNC_EVENTS - 2], 0);
      cudaEventRecord(c->ev_sync[b200_ctx::N_SYNC_EVENTS - 1], c->stream3);
      cudaStreamWaitEvent(c->stream, c->ev_sync[b200_ctx::N_SYNC_EVENTS - 1], 0);
    }
  } join_on_exit{ctx};
  // order the side streams after whatever is still queued on the main stream (previous calls)
  B200_CUDA(ctx, cudaEventRecord(ctx->ev_sync[0], ctx->stream));
  B200_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_sync[0], 0));
  B200_CUDA(ctx, cudaStreamWaitEvent(ctx->stream3, ctx->ev_sync[0], 0));
  // counting sort of (window, bucket) -> point-index lists, per group: group 0 on the main stream, group g+1 on
  // stream2 while accumulate(g) runs
  auto sort_group = [&](cudaStream_t st, int j_lo, int cnt) -> int {
    msm_plan pg = pl;
    pg.nloc = cnt;
    for (int i = 0; i < cnt; i++) pg.win[i] = pl.win[j_lo + i];
    size_t s0 = (size_t)j_lo * pl.nbuckets;
    B200_LAUNCH_ON(ctx, st, k_msm_count, nblk(n, 256), 256, 0, pg, (const uint32_t *)scalars, (const uint8_t *)inf, n, hist + s0,
                   st == ctx->stream ? bad : (uint32_t *)nullptr);   // checked once, by the first group (main stream)
    B200_LAUNCH_ON(ctx, st, k_msm_scan, cnt, 1024, 0, pl.nbuckets, hist + s0, offsets + s0);
    B200_LAUNCH_ON(ctx, st, k_msm_scatter, nblk(n, 256), 256, 0, pg, (const uint32_t *)scalars, (const uint8_t *)inf, n,
                   sstride, offsets + s0, cursor + s0, sorted + (size_t)j_lo * sstride);
    return B200_OK;
  };
  {
    int rc0 = sort_group(ctx->stream, pl.nloc - gsz[0], gsz[0]);
    if (rc0 != B200_OK) return rc0;
  }
  int j_top = pl.nloc - 1, prev_w = -1;
  for (int g = 0; g < ng; g++) {
    int cnt = gsz[g], j_lo = j_top - cnt + 1;
    size_t s0 = (size_t)j_lo * pl.nbuckets, gtotal = (size_t)cnt * pl.nbuckets;
    if (g > 0) B200_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_sync[20 + g], 0));  // sort(g) done on stream2
    if (g == 0) {
      int rcw = wait_points();   // the bucket kernels are the first to read the points
      if (rcw != B200_OK) return rcw;
    }
    // per-group scheduling scratch: SIZE_BINS-sized arrays are double-buffered by group parity
    uint32_t *sh = size_hist + (size_t)(g & 1) * 2 * SIZE_BINS, *sc = sh + SIZE_BINS;
    uint32_t *sb = size_base + (size_t)(g & 1) * SIZE_BINS;
    // affine tree levels of this group
    const uint32_t *cprev = hist + s0, *in_off = offsets + s0;
    for (int l = 1; l <= LV; l++) {
      int K = 64 >> (l - 1);
      if (K < 8) K = 8;
      B200_LAUNCH(ctx, k_aff_counts, nblk(gtotal, 256), 256, 0, gtotal, cprev, hist + s0, GIANT_BUCKET, lv_np, lv_cnt[l]);
      B200_LAUNCH(ctx, k_msm_scan, cnt, 1024, 0, pl.nbuckets, lv_np, lv_po);
      B200_LAUNCH(ctx, k_msm_scan, cnt, 1024, 0, pl.nbuckets, lv_cnt[l], lv_off[l]);
      size_t maxpairs = (sstride >> l) + 1;
      dim3 lgrid(nblk(maxpairs, (unsigned)K * 128u), cnt);
      if (l == 1) {
        B200_LAUNCH(ctx, (k_aff_level<F, true>), lgrid, 128, 0, pl.nbuckets, K, (const char *)points, (const char *)bx, sstride,
                    in_off, sorted + (size_t)j_lo * sstride, (const char *)nullptr, (size_t)0, lv_np, lv_po, lv_off[l], lv_buf[l],
                    lv_cap[l], ctx->inv_pow2);
        B200_LAUNCH(ctx, (k_aff_leftover<F, true>), nblk(gtotal, 256), 256, 0, pl.nbuckets, gtotal, (const char *)points,
                    (const char *)bx, sstride, cprev, hist + s0, GIANT_BUCKET, in_off, sorted + (size_t)j_lo * sstride,
                    (const char *)nullptr, (size_t)0, lv_np, lv_off[l], lv_buf[l], lv_cap[l]);
      } else {
        B200_LAUNCH(ctx, (k_aff_level<F, false>), lgrid, 128, 0, pl.nbuckets, K, (const char *)points, (const char *)bx, sstride,
                    in_off, (const uint32_t *)nullptr, (const char *)lv_buf[l - 1], lv_cap[l - 1], lv_np, lv_po, lv_off[l],
                    lv_buf[l], lv_cap[l], ctx->inv_pow2);
        B200_LAUNCH(ctx, (k_aff_leftover<F, false>), nblk(gtotal, 256), 256, 0, pl.nbuckets, gtotal, (const char *)points,
                    (const char *)bx, sstride, cprev, hist + s0, GIANT_BUCKET, in_off, (const uint32_t *)nullptr,
                    (const char *)lv_buf[l - 1], lv_cap[l - 1], lv_np, lv_off[l], lv_buf[l], lv_cap[l]);
      }
      cprev = lv_cnt[l];
      in_off = lv_off[l];
    }
    B200_CUDA(ctx, cudaMemsetAsync(sh, 0, 2 * SIZE_BINS * sizeof(uint32_t), ctx->stream));
    B200_LAUNCH(ctx, k_msm_size_hist, nblk(gtotal, 256), 256, 0, gtotal, cprev, hist + s0, sh);
    B200_LAUNCH(ctx, k_msm_size_scan, 1, SIZE_BINS, 0, sh, sb);
    B200_LAUNCH(ctx, k_msm_size_scatter, nblk(gtotal, 256), 256, 0, gtotal, cprev, hist + s0, sb, sc, order + s0);
    if (LV > 0) {
      B200_LAUNCH(ctx, (k_msm_accumulate_buf<F, (sizeof(F) == sizeof(fp) ? 3 : 2)>), nblk(gtotal, 128), 128, 0, pl.nbuckets, gtotal,
                  s0, (const char *)lv_buf[LV], lv_cap[LV], lv_cnt[LV], lv_off[LV], hist, GIANT_BUCKET, order + s0, buckets);
    } else if constexpr (sizeof(F) == sizeof(fp)) {
      if (ctx->tune_g1_prefetch) {
      B200_LAUNCH(ctx, (k_msm_accumulate<F, 3, true>), nblk(gtotal, 128), 128, 2 * 2 * field_traits<F>::bytes * 128, pl.nbuckets,
                  gtotal, s0, (const char *)points, bx, sstride, offsets, hist, sorted, order + s0, buckets);
      } else {
      B200_LAUNCH(ctx, (k_msm_accumulate<F, 3, false>), nblk(gtotal, 128), 128, 0, pl.nbuckets, gtotal, s0, (const char *)points,
                  bx, sstride, offsets, hist, sorted, order + s0, buckets);
      }
    } else if (ctx->tune_g2_acc_blocks == 2) {   // accumulator in registers: 255 regs, 2 blocks/SM
      B200_LAUNCH(ctx, (k_msm_accumulate<F, 2, false>), nblk(gtotal, 128), 128, 0, pl.nbuckets, gtotal, s0, (const char *)points,
                  (const char *)nullptr, sstride, offsets, hist, sorted, order + s0, buckets);
    } else if (ctx->tune_g2_acc_blocks == 3) {   // accumulator in shared memory, built for 3 blocks/SM
      B200_LAUNCH(ctx, (k_msm_accumulate_g2sm<3>), nblk(gtotal, 128), 128, 96 * 128 * 4, pl.nbuckets, gtotal, s0,
                  (const char *)points, sstride, offsets, hist, sorted, order + s0, buckets);
    } else {                                     // accumulator in shared memory, 2 blocks/SM (255 regs)
      B200_LAUNCH(ctx, (k_msm_accumulate_g2sm<2>), nblk(gtotal, 128), 128, 96 * 128 * 4, pl.nbuckets, gtotal, s0,
                  (const char *)points, sstride, offsets, hist, sorted, order + s0, buckets);
    }
    // giant buckets of this group (normally none: both kernels read the count on the device and exit at once)
    B200_LAUNCH(ctx, k_msm_giant_parts<F>, 4 * ctx->sm_count, 128, 128 * PB, pl.nbuckets, s0, sh, order + s0, (const char *)points,
                (const char *)bx, sstride, offsets, hist, sorted, gparts, max_giants);
    B200_LAUNCH(ctx, k_msm_giant_final<F>, 64, 32, 0, s0, sh, order + s0, gparts, buckets, max_giants);
    B200_CUDA(ctx, cudaEventRecord(ctx->ev_sync[1 + 2 * g], ctx->stream));
    if (g + 1 < ng) {  // sort the next group under this group's accumulate
      int rc1 = sort_group(ctx->stream2, j_lo - gsz[g + 1], gsz[g + 1]);
      if (rc1 != B200_OK) return rc1;
      B200_CUDA(ctx, cudaEventRecord(ctx->ev_sync[20 + g + 1], ctx->stream2));
    }
    B200_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_sync[1 + 2 * g], 0));
    const bool coop_here = reduce_mode == 2 || (reduce_mode == 1 && g == ng - 1);
    if (coop_here) {
      dim3 g0((unsigned)((r_n[0] + 5 * RWARPS - 1) / (5 * RWARPS)), cnt);
      B200_LAUNCH_ON(ctx, ctx->stream2, (k_msm_reduce_coop<F, RWARPS>), g0, 32 * RWARPS, 0, pl.nbuckets, r_chunk, r_n[0], r_lo_bits,
                     buckets + PB * s0, r_buf[0] + PB * (size_t)j_lo * r_n[0]);
      for (int l = 1; l < r_levels; l++) {
        dim3 gl((unsigned)((r_n[l] + 5 * RWARPS - 1) / (5 * RWARPS)), cnt);
        B200_LAUNCH_ON(ctx, ctx->stream2, (k_msm_fold_coop<F, RWARPS>), gl, 32 * RWARPS, 0, r_n[l - 1], RFOLD, r_n[l],
                       r_buf[l - 1] + PB * (size_t)j_lo * r_n[l - 1], r_buf[l] + PB * (size_t)j_lo * r_n[l]);
      }
    } else {
      dim3 rgrid(blocks_per_window, cnt);
      B200_LAUNCH_ON(ctx, ctx->stream2, (k_msm_reduce<F, RB>), rgrid, RB, RB * PB, pl.nbuckets, chunk, buckets + PB * s0,
                     partials + PB * (size_t)j_lo * blocks_per_window);
    }
    B200_CUDA(ctx, cudaEventRecord(ctx->ev_sync[2 + 2 * g], ctx->stream2));
    B200_CUDA(ctx, cudaStreamWaitEvent(ctx->stream3, ctx->ev_sync[2 + 2 * g], 0));
    bool last = g == ng - 1;
    // doublings applied at the END of this launch: down to bit 0 after the last group, else down to the next group's top
    // window (so that the next launch starts with an addition as soon as its window sums exist)
    int final_shift = last ? pl.win[0] * pl.c : (pl.win[j_lo] - pl.win[j_lo - 1]) * pl.c;
    B200_LAUNCH_ON(ctx, ctx->stream3, k_msm_horner<F>, 1, 32, 0, pl, j_top, cnt, prev_w, final_shift,
                   coop_here ? r_n[r_levels - 1] : blocks_per_window,
                   coop_here ? (const char *)r_buf[r_levels - 1] : (const char *)partials, hacc);
    prev_w = pl.win[j_lo];
    j_top = j_lo - 1;
  }
  B200_CUDA(ctx, cudaEventRecord(ctx->ev_sync[b200_ctx::N_SYNC_EVENTS - 3], ctx->stream3));
  B200_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_sync[b200_ctx::N_SYNC_EVENTS - 3], 0));
  B200_CUDA(ctx, cudaMemcpyAsync(out, hacc, PB, cudaMemcpyDeviceToDevice, ctx->stream));
  return B200_OK;
}

template <class F>
int msm_host(b200_ctx *ctx, const void *points, const uint8_t *inf, const void *scalars, size_t n, void *out) {
  constexpr size_t FB = field_traits<F>::bytes, AB = 2 * FB, PB = 3 * FB;
  int rc = stage_reserve(ctx, AB * n + 32 * n + n + PB + 8 * 256);
  if (rc != B200_OK) return rc;
  void *dp = stage_take(ctx, AB * n), *ds = stage_take(ctx, 32 * n), *di = inf ? stage_take(ctx, n) : nullptr;
  void *dout = stage_take(ctx, PB);
  if (n) {
    // scalars (and flags) first, on the main stream: the window digits / counting sort need nothing else.  The points —
    // three (G1) or six (G2) times as many bytes — follow on stream3 and overlap the sort; the bucket kernel waits for them.
    B200_CUDA(ctx, cudaEventRecord(ctx->ev_sync[31], ctx->stream));               // earlier users of the staging buffer
    B200_CUDA(ctx, cudaStreamWaitEvent(ctx->stream3, ctx->ev_sync[31], 0));
    B200_CUDA(ctx, cudaMemcpyAsync(ds, scalars, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
    if (inf) B200_CUDA(ctx, cudaMemcpyAsync(di, inf, n, cudaMemcpyHostToDevice, ctx->stream));
    B200_CUDA(ctx, cudaMemcpyAsync(dp, points, AB * n, cudaMemcpyHostToDevice, ctx->stream3));
    B200_CUDA(ctx, cudaEventRecord(ctx->ev_sync[30], ctx->stream3));
    ctx->msm_points_event_pending = true;
  }
  rc = msm_dev<F>(ctx, dp, di, ds, n, 0, 1, dout);
  if (ctx->msm_points_event_pending) {   // msm_dev left before its first bucket kernel: keep the ordering anyway
    ctx->msm_points_event_pending = false;
    cudaStreamWaitEvent(ctx->stream, ctx->ev_sync[30], 0);
  }
  if (rc != B200_OK) return rc;
  B200_CUDA(ctx, cudaMemcpyAsync(out, dout, PB, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return b200i_msm_check(ctx);
}

}  // namespace

// after the stream has been synchronised: did the last MSM of this ctx see a non-canonical scalar?
int b200i_msm_check(b200_ctx *ctx) {
  if (ctx->msm_bad_flag == nullptr) return B200_OK;
  uint32_t bad = 0;
  B200_CUDA(ctx, cudaMemcpyAsync(&bad, ctx->msm_bad_flag, sizeof(bad), cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (bad) {
    snprintf(ctx->err, sizeof(ctx->err), "MSM: a scalar is not canonical (>= q); the ABI takes Scalar::to_bytes()");
    return B200_EINVAL;
  }
  return B200_OK;
}
// enqueue-only MSM (shard of the windows) for capi_multi.cu (no lock, no synchronisation)
int b200i_msm_enqueue(b200_ctx *ctx, int k, const void *points, const void *inf, const void *scalars, size_t n, int shard,
                      int n_shards, void *out) {
  return k == 1 ? msm_dev<fp>(ctx, points, inf, scalars, n, shard, n_shards, out)
                : msm_dev<fp2>(ctx, points, inf, scalars, n, shard, n_shards, out);
}

#define CHECK_CTX(ctx)                      \
  if ((ctx) == nullptr) return B200_EINVAL; \
  ctx_guard guard__(ctx);                   \
  if (!guard__.ok) return B200_ENODEV

extern "C" {

int b200_g1_msm_shard_dev(b200_ctx *ctx, const void *points, const void *inf, const void *scalars, size_t n, int shard,
                          int n_shards, void *out) {
  CHECK_CTX(ctx);
  if (!out || (n && (!points || !scalars))) return B200_EINVAL;
  int rc = msm_dev<fp>(ctx, points, inf, scalars, n, shard, n_shards, out);
  if (rc != B200_OK) return rc;
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return b200i_msm_check(ctx);
}
int b200_g2_msm_shard_dev(b200_ctx *ctx, const void *points, const void *inf, const void *scalars, size_t n, int shard,
                          int n_shards, void *out) {
  CHECK_CTX(ctx);
  if (!out || (n && (!points || !scalars))) return B200_EINVAL;
  int rc = msm_dev<fp2>(ctx, points, inf, scalars, n, shard, n_shards, out);
  if (rc != B200_OK) return rc;
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return b200i_msm_check(ctx);
}
int b200_glv_decompose(b200_ctx *ctx, const b200_scalar *scalars, size_t n, uint8_t *k1k2, uint8_t *signs) {
  CHECK_CTX(ctx);
  if (n && (!scalars || !k1k2 || !signs)) return B200_EINVAL;
  if (n == 0) return B200_OK;
  int rc = stage_reserve(ctx, 32 * n + 32 * n + n + 1024);
  if (rc != B200_OK) return rc;
  void *ds = stage_take(ctx, 32 * n), *dk = stage_take(ctx, 32 * n), *dg = stage_take(ctx, n);
  B200_CUDA(ctx, cudaMemcpyAsync(ds, scalars, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
  B200_LAUNCH(ctx, k_glv_decompose, nblk(n, 128), 128, 0, (const uint32_t *)ds, n, (uint32_t *)dk, (uint8_t *)dg);
  B200_CUDA(ctx, cudaMemcpyAsync(k1k2, dk, 32 * n, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaMemcpyAsync(signs, dg, n, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
int b200_g1_msm_dev(b200_ctx *ctx, const void *points, const void *inf, const void *scalars, size_t n, void *out) {
  return b200_g1_msm_shard_dev(ctx, points, inf, scalars, n, 0, 1, out);
}
int b200_g2_msm_dev(b200_ctx *ctx, const void *points, const void *inf, const void *scalars, size_t n, void *out) {
  return b200_g2_msm_shard_dev(ctx, points, inf, scalars, n, 0, 1, out);
}
int b200_g1_msm(b200_ctx *ctx, const b200_g1_affine *points, const uint8_t *inf, const b200_scalar *scalars, size_t n,
                b200_g1_projective *out) {
  CHECK_CTX(ctx);
  if (!out || (n && (!points || !scalars))) return B200_EINVAL;
  return msm_host<fp>(ctx, points, inf, scalars, n, out);
}
int b200_g2_msm(b200_ctx *ctx, const b200_g2_affine *points, const uint8_t *inf, const b200_scalar *scalars, size_t n,
                b200_g2_projective *out) {
  CHECK_CTX(ctx);
  if (!out || (n && (!points || !scalars))) return B200_EINVAL;
  return msm_host<fp2>(ctx, points, inf, scalars, n, out);
}

}  // extern "C"
