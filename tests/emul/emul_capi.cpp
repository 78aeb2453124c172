// TEST INFRASTRUCTURE.  The device-side arithmetic of bls12_381_b200/csrc/ compiled for the HOST (see cuda_host_shim.h)
// behind a flat C API that mirrors include/bls12381_b200.h's array conventions, so tests/test_device_source_cpu.py can
// compare it with the oracle without a GPU.  Built twice: default (lazy-reduction Fp2, the G2/MSM translation units)
// and -DB200_FP2_KCALL (the pairing translation units).
#include "cuda_host_shim.h"
#include "fiber_warp.h"

#include "curve.cuh"
#include "curve_warp.cuh"
#include "fp_inv.cuh"
#include "glv.cuh"
#include "pairing.cuh"
#include "gt.cuh"
#ifdef EMUL_WITH_FR
#include "fr_ntt.cuh"
#include "h2c.cuh"
#endif

using namespace b200;

namespace {
std::vector<uint32_t> &pow2_table() {
  static std::vector<uint32_t> t;
  if (t.empty()) {
    t.resize(FP_INV_TABLE_WORDS);
    threadIdx = emul_dim3{0, 0, 0};
    blockIdx = emul_dim3{0, 0, 0};
    k_fp_inv_table_init(t.data());
  }
  return t;
}
template <class Fn>
void par_for(size_t n, int threads, Fn fn) {
  if (threads <= 1 || n < 2) {
    for (size_t i = 0; i < n; i++) fn(i);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++)
    th.emplace_back([=] {
      for (size_t i = t; i < n; i += threads) fn(i);
    });
  for (auto &x : th) x.join();
}
}  // namespace

extern "C" {

// op codes of include/bls12381_b200.h (B200_OP_*); extra codes >= 100 select implementation variants of Fp mul/sqr
int emul_tower_op(int level, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
  const uint32_t *pow2 = pow2_table().data();
  const int W = 6 * level;
  for (size_t i = 0; i < n; i++) {
    const uint64_t *pa = a + W * i, *pb = b ? b + W * i : nullptr;
    uint64_t *po = out + W * i;
    if (level == 1) {
      fp x = fp_load(pa), y = pb ? fp_load(pb) : fp_zero(), r;
      switch (op) {
        case 0: r = fp_mul(x, y); break;
        case 1: r = fp_add(x, y); break;
        case 2: r = fp_sub(x, y); break;
        case 3: r = fp_sqr_c(x); break;
        case 4: r = fp_neg(x); break;
        case 5: r = fp_inv(x); break;
        case 10: r = fp_inv_fast(x, pow2); break;
        case 100: r = fp_mul_c(x, y); break;
        case 101: r = fp_redc_wide(fp_mul_wide(x, y)); break;
        case 102: r = fp_sqr(x); break;
        case 103: r = fp_add_nr(x, y); break;  // caller keeps x + y < 2^384
        case 104: r = fp_mul_dual(x, y, y, x).r0; break;          // dual-stream product, first result
        case 105: r = fp_mul_dual(y, y, x, y).r1; break;          // ... second result of (y*y, x*y)
        default: return -1;
      }
      fp_store(po, r);
    } else if (level == 2) {
      fp2 x = fp2_load(pa), y = pb ? fp2_load(pb) : fp2_zero(), r;
      switch (op) {
        case 0: r = M2(x, y); break;
        case 1: r = fp2_add(x, y); break;
        case 2: r = fp2_sub(x, y); break;
        case 3: r = S2(x); break;
        case 4: r = fp2_neg(x); break;
        case 5: r = fp2_inv(x); break;
        case 6: case 7: r = fp2_conj(x); break;
        case 8: r = fp2_mul_by_nonresidue(x); break;
        case 100: r = fp2_mul(x, y); break;
        case 101: r = fp2_sqr(x); break;
        case 102: r = fp2_inv_ni(x); break;
        default: return -1;
      }
      fp2_store(po, r);
    } else if (level == 6) {
      fp6 x, y, r;
      fp6_load(&x, pa);
      if (pb) fp6_load(&y, pb);
      switch (op) {
        case 0: fp6_mul(&r, &x, &y); break;
        case 1: fp6_add(&r, &x, &y); break;
        case 2: fp6_sub(&r, &x, &y); break;
        case 3: fp6_sqr(&r, &x); break;
        case 4: fp6_neg(&r, &x); break;
        case 5: fp6_inv(&r, &x); break;
        case 6: fp6_frobenius(&r, &x); break;
        case 8: fp6_mul_by_nonresidue(&r, &x); break;
        default: return -1;
      }
      fp6_store(po, &r);
    } else if (level == 12) {
      fp12 x, y, r;
      fp12_load(&x, pa);
      if (pb) fp12_load(&y, pb);
      switch (op) {
        case 0: fp12_mul(&r, &x, &y); break;
        case 3: fp12_sqr(&r, &x); break;
        case 5: fp12_inv(&r, &x); break;
        case 6: fp12_frobenius(&r, &x); break;
        case 7: fp12_conj(&r, &x); break;
        case 9: cyclotomic_square(&r, &x); break;
        default: return -1;
      }
      fp12_store(po, &r);
    } else {
      return -1;
    }
  }
  return 0;
}

#define GROUP_API(NAME, F)                                                                                            \
  void emul_##NAME##_double(const char *p, char *out, size_t n) {                                                     \
    constexpr size_t PB = 3 * field_traits<F>::bytes;                                                                 \
    for (size_t i = 0; i < n; i++) proj_store<F>(out + PB * i, proj_double(proj_load<F>(p + PB * i)));                 \
  }                                                                                                                   \
  void emul_##NAME##_add(const char *p, const char *q, char *out, size_t n) {                                         \
    constexpr size_t PB = 3 * field_traits<F>::bytes;                                                                 \
    for (size_t i = 0; i < n; i++)                                                                                    \
      proj_store<F>(out + PB * i, proj_add(proj_load<F>(p + PB * i), proj_load<F>(q + PB * i)));                       \
  }                                                                                                                   \
  void emul_##NAME##_add_mixed(const char *p, const char *qxy, const uint8_t *qinf, char *out, size_t n) {            \
    constexpr size_t PB = 3 * field_traits<F>::bytes;                                                                 \
    for (size_t i = 0; i < n; i++)                                                                                    \
      proj_store<F>(out + PB * i, proj_add_mixed(proj_load<F>(p + PB * i), affine_load<F>(qxy, qinf, i)));             \
  }                                                                                                                   \
  void emul_##NAME##_mul(const char *p, const uint32_t *s, char *out, size_t n, int threads) {                        \
    constexpr size_t PB = 3 * field_traits<F>::bytes;                                                                 \
    par_for(n, threads, [=](size_t i) {                                                                               \
      uint32_t by[8];                                                                                                 \
      memcpy(by, s + 8 * i, 32);                                                                                      \
      proj_store<F>(out + PB * i, proj_multiply(proj_load<F>(p + PB * i), by));                                        \
    });                                                                                                               \
  }                                                                                                                   \
  void emul_##NAME##_to_affine(const char *p, char *xy, uint8_t *inf, size_t n) {                                     \
    constexpr size_t PB = 3 * field_traits<F>::bytes;                                                                 \
    for (size_t i = 0; i < n; i++) affine_store<F>(xy, inf, i, proj_to_affine(proj_load<F>(p + PB * i)));              \
  }                                                                                                                   \
  /* the bucket accumulator of the MSM: fold affine points into one XYZZ sum, in order */                             \
  void emul_##NAME##_xyzz_sum(const char *xy, const uint8_t *inf, size_t n, char *out) {                              \
    xyzz<F> acc = xyzz_identity<F>();                                                                                 \
    for (size_t i = 0; i < n; i++) {                                                                                  \
      affine<F> a = affine_load<F>(xy, inf, i);                                                                       \
      if (a.inf) continue;                                                                                            \
      acc = xyzz_add_mixed(acc, a.x, a.y);                                                                            \
    }                                                                                                                 \
    proj_store<F>(out, xyzz_to_proj(acc));                                                                            \
  }
GROUP_API(g1, fp)
GROUP_API(g2, fp2)

void emul_miller_loop(const char *pxy, const uint8_t *pinf, const char *qxy, const uint8_t *qinf, size_t n, uint64_t *out,
                      int threads) {
  par_for(n, threads, [=](size_t i) {
    fp12 f;
    miller_loop_pair(&f, affine_load<fp>(pxy, pinf, i), affine_load<fp2>(qxy, qinf, i));
    fp12_store(out + 72 * i, &f);
  });
}
void emul_final_exponentiation(const uint64_t *in, size_t n, uint64_t *out, int threads) {
  par_for(n, threads, [=](size_t i) {
    fp12 f;
    fp12_load(&f, in + 72 * i);
    final_exponentiation(&f);
    fp12_store(out + 72 * i, &f);
  });
}
void emul_g2_prepare(const char *qxy, int qinf, char *coeffs /* 68 * 288 B */) {
  uint8_t flag = qinf ? 1 : 0;
  g2_prepare(affine_load<fp2>(qxy, &flag, 0), coeffs);
}
void emul_miller_loop_prepared(const char *pxy, int pinf, const char *coeffs, int qinf, uint64_t *out) {
  uint8_t flag = pinf ? 1 : 0;
  fp12 f;
  miller_loop_prepared(&f, affine_load<fp>(pxy, &flag, 0), coeffs, qinf != 0);
  fp12_store(out, &f);
}
}  // extern "C"
namespace {
// the body of capi_basic.cu::k_mul_batch_warp: ONE WARP per item, lane-parallel group operations of curve_warp.cuh
// (shuffles) — run on the fiber scheduler of fiber_warp.h
template <class F>
void k_emul_mul_batch_warp(const char *p, const uint32_t *s, char *out, size_t n) {
  size_t item = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (item >= n) return;
  constexpr size_t PB = 3 * field_traits<F>::bytes;
  uint32_t by[8];
  memcpy(by, s + 8 * item, 32);
  proj<F> base = proj_load<F>(p + PB * item), acc = proj_identity<F>();
  for (int bit = 254; bit >= 0; bit--) {
    acc = warp_double(acc, lane);
    if ((by[bit >> 5] >> (bit & 31)) & 1) acc = warp_add(acc, base, lane);
  }
  if (lane == 0) proj_store<F>(out + PB * item, acc);
}
}  // namespace
extern "C" {
// 128-thread blocks = 4 items per block, like the GPU launch
void emul_g1_mul_warp(const char *p, const uint32_t *s, char *out, size_t n) {
  emul_cooperative_launch(k_emul_mul_batch_warp<fp>, (unsigned)((n * 32 + 127) / 128), 128u, p, s, out, n);
}
void emul_g2_mul_warp(const char *p, const uint32_t *s, char *out, size_t n) {
  emul_cooperative_launch(k_emul_mul_batch_warp<fp2>, (unsigned)((n * 32 + 127) / 128), 128u, p, s, out, n);
}
// self-test of the fiber scheduler: block reduction through shared memory + __syncthreads, warp sums through
// __shfl_down_sync, a ballot and an early-exiting warp
int emul_fiber_selftest(unsigned block, const int *in, int n, int *block_sum, int *warp_sums, unsigned *ballots) {
  static int sh[1024];
  auto kernel = [](const int *in, int n, int *block_sum, int *warp_sums, unsigned *ballots) {
    unsigned t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (w == 1 && blockDim.x > 64) return;          // a whole warp leaves early: barriers must not wait for it
    int v = (int)t < n ? in[t] : 0;
    sh[t] = v;
    __syncthreads();
    for (unsigned st = 1; st < blockDim.x; st <<= 1) {    // naive tree, every level separated by a barrier
      int add = (t % (2 * st) == 0 && t + st < blockDim.x) ? sh[t + st] : 0;
      __syncthreads();
      sh[t] += add;
      __syncthreads();
    }
    if (t == 0) *block_sum = sh[0];
    int x = v;
    for (int d = 16; d > 0; d >>= 1) x += __shfl_down_sync(0xffffffffu, x, d);
    if (lane == 0) warp_sums[w] = x;
    unsigned b = __ballot_sync(0xffffffffu, v & 1);
    if (lane == 0) ballots[w] = b;
  };
  for (unsigned i = 0; i < block; i++) sh[i] = 0;
  emul_cooperative_launch(kernel, 1u, block, in, n, block_sum, warp_sums, ballots);
  return 0;
}
void emul_gt_mul(const char *g, const uint32_t *s, size_t n, char *out, int threads) {
  par_for(n, threads, [=](size_t i) {
    fp12 x, acc;
    fp12_load(&x, g + 576 * i);
    gt_mul_scalar(&acc, &x, s + 8 * i);
    fp12_store(out + 576 * i, &acc);
  });
}
// out: k1[4] k2[4] neg1 neg2 (10 words per scalar)
void emul_glv_decompose(const uint32_t *s, size_t n, uint32_t *out) {
  for (size_t i = 0; i < n; i++) {
    glv_parts g = glv_decompose(s + 8 * i);
    memcpy(out + 10 * i, g.k1, 16);
    memcpy(out + 10 * i + 4, g.k2, 16);
    out[10 * i + 8] = g.neg1;
    out[10 * i + 9] = g.neg2;
  }
}

}  // extern "C"
#ifdef EMUL_WITH_FR
namespace {
// a kernel "launch" on the host: every (block, thread) in turn; blocks are spread over host threads.  Only valid for
// kernels without intra-block cooperation (no __syncthreads / shared memory / shuffles) — all of fr_ntt.cuh.
struct emul_launcher {
  int threads = 1;
  int launches = 0;
  template <class K, class... A>
  int operator()(K k, unsigned grid, unsigned block, A... a) {
    launches++;
    auto run = [=](unsigned b0, unsigned step) {
      blockDim = emul_dim3{block, 1, 1};
      gridDim = emul_dim3{grid, 1, 1};
      for (unsigned b = b0; b < grid; b += step) {
        blockIdx = emul_dim3{b, 0, 0};
        for (unsigned t = 0; t < block; t++) {
          threadIdx = emul_dim3{t, 0, 0};
          k(a...);
        }
      }
    };
    if (threads <= 1 || grid < 2) {
      run(0, 1);
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < threads; t++) th.emplace_back(run, (unsigned)t, (unsigned)threads);
      for (auto &x : th) x.join();
    }
    return 0;
  }
};
}  // namespace
extern "C" {
int emul_fr_op(int op, const char *a, const char *b, char *out, size_t n) {
  emul_launcher l;
  return l(k_fr_op, (unsigned)((n + 255) / 256), 256u, op, a, b, out, n);
}
int emul_fr_to_bytes(const char *a, char *out, size_t n) {
  emul_launcher l;
  return l(k_fr_to_bytes, (unsigned)((n + 255) / 256), 256u, a, out, n);
}
int emul_fr_from_bytes(const char *in, char *out, uint8_t *ok, size_t n) {
  emul_launcher l;
  return l(k_fr_from_bytes, (unsigned)((n + 255) / 256), 256u, in, out, ok, n);
}
// hash to curve: the launch sequence of b200_g{1,2}_hash_to_curve
int emul_h2c_expand(const uint8_t *msgs, const uint64_t *off, size_t n, const uint8_t *dst, size_t dst_len, uint32_t len_in_bytes,
                    uint8_t *out) {
  uint8_t dp[256];
  int dpl = h2c_dst_prime(dst, dst_len, dp);
  emul_launcher l;
  return l(k_h2c_expand, (unsigned)((n + 127) / 128), 128u, msgs, off, n, (const uint8_t *)dp, dpl, len_in_bytes, out);
}
int emul_h2c_hash(int group, const uint8_t *msgs, const uint64_t *off, size_t n, const uint8_t *dst, size_t dst_len, int encode,
                  char *out, int threads) {
  uint8_t dp[256];
  int dpl = h2c_dst_prime(dst, dst_len, dp);
  const int count = encode ? 1 : 2;
  std::vector<uint8_t> okm((size_t)h2c_okm_bytes(group, count) * n + 1);
  emul_launcher l;
  l.threads = threads;
  return h2c_hash_run(l, group, msgs, off, n, (const uint8_t *)dp, dpl, count, okm.data(), out);
}
int emul_h2c_stage(int group, int kind, const char *in, size_t n, char *out) {
  emul_launcher l;
  if (group == 1) return l(k_h2c_stage<fp>, (unsigned)((n + 127) / 128), 128u, kind, in, n, out);
  return l(k_h2c_stage<fp2>, (unsigned)((n + 127) / 128), 128u, kind, in, n, out);
}
// the full launch sequence of b200_fr_ntt (tables + passes); returns the number of pass launches
int emul_fr_ntt(const char *in, int log_n, int inverse, int coset, char *out, int threads) {
  emul_launcher l;
  l.threads = threads;
  std::vector<char> mem(fr_tables_bytes(log_n) + 256);
  char *base = mem.data() + ((256 - ((uintptr_t)mem.data() & 255)) & 255);
  fr_ntt_tables tb{};
  int rc = fr_ntt_build_tables(l, base, log_n, &tb);
  if (rc < 0) return rc;
  return fr_ntt_run(l, in, out, log_n, inverse != 0, coset != 0, tb);
}
}  // extern "C"
#endif
