// Batched scalar-field kernels and the number-theoretic transform over Fr.
//
// The reference exports the field (src/scalar.rs) and the 2^32-th root of unity the transform is built on
// (ROOT_OF_UNITY :200, S = 32 :191, MULTIPLICATIVE_GENERATOR = 7 :100); the transform itself is what the callers
// next to the MSM compute (bellman's EvaluationDomain::{fft, ifft, coset_fft, icoset_fft}):
//     forward   out[k] = sum_j a[j] w^(jk),          w = ROOT_OF_UNITY^(2^(32 - log_n))
//     inverse   out[j] = n^-1 sum_k a[k] w^(-jk)
//     coset     forward evaluates on g<w> (a[j] *= g^j first), inverse undoes it (out[j] *= g^-j), g = 7
// natural order in, natural order out.  All values canonical Montgomery limbs => bit-identical to any
// implementation over the reference's Scalar.
//
// Shape on the GPU: an element is 32 B = exactly one DRAM sector, so every access pattern is sector-efficient.
// Decimation in time: pass 0 gathers its inputs from the bit-reversed positions (out of place), then each pass does
// up to THREE butterfly stages on 8 elements held in registers by one thread (in place, no inter-thread traffic,
// no shared memory): ceil(log_n / 3) passes, each reading and writing the array once.  Stage t pairs (i, i + 2^t)
// with the twiddle w^((i mod 2^t) * n / 2^(t+1)) from an n/2-entry table (device memory, cached per log_n; the
// inverse transform reads the same table: w^-e = -w^(n/2 - e)).
// Work: (n/2) log_n butterflies x 1 multiplication (136 IMAD) — integer-pipe bound like everything else here:
// 2^24: 2.0e8 x 136 = 2.7e10 IMAD = 3.1 ms at the measured 8.77e12 IMAD/s; traffic 8 passes x 1 GiB = 1.3 ms.
#pragma once
#include "fr.cuh"

namespace b200 {

constexpr int FR_NTT_MAX_LOG_N = 28;
constexpr int FR_POW_LO_BITS = 12;  // power tables: LO[j] = base^j (j < 4096), HI[j] = scale * base^(4096 j)
constexpr uint32_t FR_POW_LO_MASK = (1u << FR_POW_LO_BITS) - 1;

enum { FR_BASE_OMEGA = 0, FR_BASE_GEN = 1, FR_BASE_GEN_INV = 2 };
enum { FR_SCALE_ONE = 0, FR_SCALE_NINV = 1 };

// out[j] = scale * base^(j << shift), j < count.  base: w_n = ROOT_OF_UNITY^(2^(32 - log_n)), g = 7 or g^-1;
// scale: 1 or n^-1 = TWO_INV^log_n.  One thread per entry (square-and-multiply over a <= 64-bit exponent).
static __global__ void __launch_bounds__(128) k_fr_pow_table(int base_kind, int scale_kind, int log_n, int shift, char *out,
                                                            uint32_t count) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  fr base;
  if (base_kind == FR_BASE_OMEGA) {
    base = fr_root_of_unity();
    for (int i = log_n; i < 32; i++) base = fr_mul_c(base, base);
  } else if (base_kind == FR_BASE_GEN) {
    base = fr_generator();
  } else {
    base = fr_inv(fr_generator());
  }
  fr r = fr_pow_u64(base, (unsigned long long)j << shift);
  if (scale_kind == FR_SCALE_NINV) {
    fr h = fr_two_inv();
    for (int i = 0; i < log_n; i++) r = fr_mul_c(r, h);
  }
  fr_store(out + 32 * (size_t)j, r);
}
// tw[e] = lo[e & 4095] * hi[e >> 12] = w^e,  e < half
static __global__ void __launch_bounds__(256) k_fr_twiddles(const char *lo, const char *hi, char *tw, size_t half) {
  size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (e >= half) return;
  fr r = fr_mul_c(fr_load_ro(lo + 32 * (e & FR_POW_LO_MASK)), fr_load_ro(hi + 32 * (e >> FR_POW_LO_BITS)));
  fr_store(tw + 32 * e, r);
}

B200_DEV fr fr_pow_lookup(const char *lo, const char *hi, size_t i) {
  return fr_mul_c(fr_load_ro(lo + 32 * (i & FR_POW_LO_MASK)), fr_load_ro(hi + 32 * (i >> FR_POW_LO_BITS)));
}
// w^e (forward) or w^-e = -w^(half - e) (inverse), 0 <= e < half
B200_DEV fr fr_twiddle(const char *tw, size_t half, size_t e, bool inverse) {
  if (!inverse || e == 0) return fr_load_ro(tw + 32 * e);
  return fr_neg(fr_load_ro(tw + 32 * (half - e)));
}
B200_DEV size_t fr_bitrev(size_t i, int log_n) {  // 1 <= log_n <= 32
  uint32_t v = (uint32_t)i, r = 0;
#pragma unroll 1
  for (int b = 0; b < log_n; b++) {
    r = (r << 1) | (v & 1u);
    v >>= 1;
  }
  return r;
}

// Stages s .. s+R-1 on 2^R elements per thread.  `first`: the pass gathers position i from in[bitrev(i)] (and applies
// the coset pre-scaling g^(bitrev(i)) when pre_lo != nullptr).  post_mode (last pass only): 0 none, 1 every output
// times the constant post_hi[0], 2 output i times post_lo[i & 4095] * post_hi[i >> 12].
// In place when in == out and !first (a thread reads and writes only its own 2^R positions).
template <int R>
__global__ void __launch_bounds__(256) k_fr_ntt_pass(const char *in, char *out, int log_n, int s, const char *tw, int inverse,
                                                   int first, const char *pre_lo, const char *pre_hi, int post_mode,
                                                   const char *post_lo, const char *post_hi) {
  constexpr int M = 1 << R;
  const size_t n = (size_t)1 << log_n, half = n >> 1;
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= (n >> R)) return;
  const size_t lo = t & (((size_t)1 << s) - 1), hi = t >> s;
  const size_t base = (hi << (s + R)) | lo;
  fr x[M];
#pragma unroll
  for (int j = 0; j < M; j++) {
    size_t idx = base + ((size_t)j << s);
    size_t src = first ? fr_bitrev(idx, log_n) : idx;
    x[j] = fr_load(in + 32 * src);
    if (first && pre_lo != nullptr) x[j] = fr_mul_c(x[j], fr_pow_lookup(pre_lo, pre_hi, src));
  }
#pragma unroll
  for (int q = 0; q < R; q++) {
    const int tt = s + q;  // stage: distance 2^tt, twiddle exponent (i mod 2^tt) << (log_n - 1 - tt)
#pragma unroll
    for (int j = 0; j < M; j++) {
      if (j & (1 << q)) continue;
      size_t imod = ((size_t)(j & ((1 << q) - 1)) << s) | lo;
      fr w = fr_twiddle(tw, half, imod << (log_n - 1 - tt), inverse != 0);
      fr v = fr_mul_c(x[j | (1 << q)], w);
      fr u = x[j];
      x[j] = fr_add(u, v);
      x[j | (1 << q)] = fr_sub(u, v);
    }
  }
#pragma unroll
  for (int j = 0; j < M; j++) {
    size_t idx = base + ((size_t)j << s);
    fr r = x[j];
    if (post_mode == 1) r = fr_mul_c(r, fr_load_ro(post_hi));
    if (post_mode == 2) r = fr_mul_c(r, fr_pow_lookup(post_lo, post_hi, idx));
    fr_store(out + 32 * idx, r);
  }
}

// n == 1: the transform is the identity (g^0 = 1, n^-1 = 1)
static __global__ void k_fr_copy1(const char *in, char *out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) fr_store(out, fr_load(in));
}

// element-wise field ops (op codes of include/bls12381_b200.h: B200_OP_MUL/ADD/SUB/SQUARE/NEG/INVERT, B200_OP_DOUBLE)
static __global__ void __launch_bounds__(256) k_fr_op(int op, const char *a, const char *b, char *out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  fr x = fr_load(a + 32 * i), y = b ? fr_load(b + 32 * i) : fr_zero(), r;
  switch (op) {
    case 0: r = fr_mul_c(x, y); break;
    case 1: r = fr_add(x, y); break;
    case 2: r = fr_sub(x, y); break;
    case 3: r = fr_mul_c(x, x); break;
    case 4: r = fr_neg(x); break;
    case 5: r = fr_inv(x); break;  // 0 -> 0 (the reference's CtOption is None there, src/scalar.rs:502)
    default: r = fr_dbl(x); break;
  }
  fr_store(out + 32 * i, r);
}
// Scalar::to_bytes (src/scalar.rs:284-296): Montgomery limbs -> canonical 32-byte little-endian integer
static __global__ void __launch_bounds__(256) k_fr_to_bytes(const char *a, char *out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  fr_store(out + 32 * i, fr_from_mont(fr_load(a + 32 * i)));
}
// Scalar::from_bytes (src/scalar.rs:256-281): ok[i] = 1 and Montgomery limbs when the integer is < q, else ok[i] = 0
// and zero limbs (the reference's CtOption is None; its inner value is unspecified)
static __global__ void __launch_bounds__(256) k_fr_from_bytes(const char *in, char *out, uint8_t *ok, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  fr x = fr_load(in + 32 * i);
  bool canon = fr_is_canonical(x);
  fr_store(out + 32 * i, canon ? fr_to_mont(x) : fr_zero());
  ok[i] = canon ? 1 : 0;
}

struct fr_ntt_tables {
  const char *tw;                  // w^e, e < n/2 (n >= 2)
  const char *g_lo, *g_hi;         // g^j / g^(4096 j)                 (forward coset)
  const char *gi_lo, *gi_hi;       // g^-j / n^-1 g^(-4096 j)          (inverse coset)
  const char *ninv;                // n^-1                             (inverse, no coset)
};

// The launch sequence of one transform.  `launch(kernel, grid, block, args...)` returns 0 on success; on the GPU it is
// a <<<>>> on the ctx stream (capi_fr.cu), in the CPU test harness a loop over (blockIdx, threadIdx) (tests/emul/).
// `in` != `out` (pass 0 is out of place).  Returns the number of launches or a negative error from `launch`.
template <class L>
int fr_ntt_run(L &&launch, const char *in, char *out, int log_n, bool inverse, bool coset, const fr_ntt_tables &tb) {
  if (log_n == 0) {
    int rc = launch(k_fr_copy1, 1u, 32u, in, out);
    return rc ? rc : 1;
  }
  const size_t n = (size_t)1 << log_n;
  int launches = 0;
  for (int s = 0; s < log_n; s += 3) {
    const int r = log_n - s >= 3 ? 3 : log_n - s;
    const bool first = s == 0, last = s + r == log_n;
    const char *src = first ? in : out;
    const char *pre_lo = (first && coset && !inverse) ? tb.g_lo : nullptr;
    const char *pre_hi = pre_lo ? tb.g_hi : nullptr;
    int post_mode = 0;
    const char *post_lo = nullptr, *post_hi = nullptr;
    if (last && inverse) {
      post_mode = coset ? 2 : 1;
      post_lo = coset ? tb.gi_lo : nullptr;
      post_hi = coset ? tb.gi_hi : tb.ninv;
    }
    const size_t threads = n >> r;
    const unsigned block = 128, grid = (unsigned)((threads + block - 1) / block);  // 138 regs (R = 3): 3 blocks = 12 warps per SM
    int rc;
    if (r == 3)
      rc = launch(k_fr_ntt_pass<3>, grid, block, src, out, log_n, s, tb.tw, (int)inverse, (int)first, pre_lo, pre_hi,
                  post_mode, post_lo, post_hi);
    else if (r == 2)
      rc = launch(k_fr_ntt_pass<2>, grid, block, src, out, log_n, s, tb.tw, (int)inverse, (int)first, pre_lo, pre_hi,
                  post_mode, post_lo, post_hi);
    else
      rc = launch(k_fr_ntt_pass<1>, grid, block, src, out, log_n, s, tb.tw, (int)inverse, (int)first, pre_lo, pre_hi,
                  post_mode, post_lo, post_hi);
    if (rc) return rc;
    launches++;
  }
  return launches;
}

// table sizes (entries) for a given log_n
inline size_t fr_tab_lo_count(int log_n) {
  size_t n = (size_t)1 << log_n;
  return n < ((size_t)1 << FR_POW_LO_BITS) ? n : ((size_t)1 << FR_POW_LO_BITS);
}
inline size_t fr_tab_hi_count(int log_n) {
  size_t h = ((size_t)1 << log_n) >> FR_POW_LO_BITS;
  return h ? h : 1;
}

// Builds every table of fr_ntt_tables into `mem` (caller-allocated, fr_tables_bytes(log_n) bytes, 256-B aligned);
// same launcher convention.  Returns the number of launches or a negative error.
inline size_t fr_tables_bytes(int log_n) {
  size_t n = (size_t)1 << log_n, half = n > 1 ? n / 2 : 1;
  size_t lo = fr_tab_lo_count(log_n), hi = fr_tab_hi_count(log_n);
  return 32 * (half + 3 * lo + 3 * hi + 1) + 8 * 256;
}
template <class L>
int fr_ntt_build_tables(L &&launch, char *mem, int log_n, fr_ntt_tables *tb) {
  const size_t n = (size_t)1 << log_n, half = n > 1 ? n / 2 : 1;
  const uint32_t lo = (uint32_t)fr_tab_lo_count(log_n), hi = (uint32_t)fr_tab_hi_count(log_n);
  size_t off = 0;
  auto take = [&](size_t count) {
    char *p = mem + off;
    off += (32 * count + 255) & ~(size_t)255;
    return p;
  };
  char *tw = take(half), *w_lo = take(lo), *w_hi = take(hi), *g_lo = take(lo), *g_hi = take(hi), *gi_lo = take(lo),
       *gi_hi = take(hi), *ninv = take(1);
  auto nb = [](uint32_t c) { return (c + 127u) / 128u; };
  int rc;
  if ((rc = launch(k_fr_pow_table, nb(lo), 128u, (int)FR_BASE_OMEGA, (int)FR_SCALE_ONE, log_n, 0, w_lo, lo))) return rc;
  if ((rc = launch(k_fr_pow_table, nb(hi), 128u, (int)FR_BASE_OMEGA, (int)FR_SCALE_ONE, log_n, FR_POW_LO_BITS, w_hi, hi))) return rc;
  if ((rc = launch(k_fr_twiddles, (unsigned)((half + 255) / 256), 256u, (const char *)w_lo, (const char *)w_hi, tw, half))) return rc;
  if ((rc = launch(k_fr_pow_table, nb(lo), 128u, (int)FR_BASE_GEN, (int)FR_SCALE_ONE, log_n, 0, g_lo, lo))) return rc;
  if ((rc = launch(k_fr_pow_table, nb(hi), 128u, (int)FR_BASE_GEN, (int)FR_SCALE_ONE, log_n, FR_POW_LO_BITS, g_hi, hi))) return rc;
  if ((rc = launch(k_fr_pow_table, nb(lo), 128u, (int)FR_BASE_GEN_INV, (int)FR_SCALE_ONE, log_n, 0, gi_lo, lo))) return rc;
  if ((rc = launch(k_fr_pow_table, nb(hi), 128u, (int)FR_BASE_GEN_INV, (int)FR_SCALE_NINV, log_n, FR_POW_LO_BITS, gi_hi, hi))) return rc;
  if ((rc = launch(k_fr_pow_table, 1u, 128u, (int)FR_BASE_GEN, (int)FR_SCALE_NINV, log_n, 0, ninv, 1u))) return rc;
  tb->tw = tw;
  tb->g_lo = g_lo;
  tb->g_hi = g_hi;
  tb->gi_lo = gi_lo;
  tb->gi_hi = gi_hi;
  tb->ninv = ninv;
  return 8;
}

}  // namespace b200
