// ORACLE — TEST INFRASTRUCTURE ONLY (see bls12_381_oracle.hpp header).  Flat C entry points over
// the CPU restatement so tests/, smoke() and bench.py's cpu_baseline / --impl reference legs can
// call it through ctypes.  Data layouts are the same as include/bls12381_b200.h.
#include <atomic>
#include <thread>

#include "bls12_381_oracle.hpp"
#include "h2c_oracle.hpp"

using namespace bls_oracle;

namespace {
template <class F>
void parallel_for(size_t n, int threads, F f) {
  if (threads <= 1 || n < 2) {
    for (size_t i = 0; i < n; i++) f(i);
    return;
  }
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([&] {
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= n) break;
        f(i);
      }
    });
  for (auto &th : pool) th.join();
}
inline G1Affine load_g1a(const uint64_t *xy, const uint8_t *inf, size_t i) {
  G1Affine p;
  std::memcpy(&p.x, xy + 12 * i, 96);
  p.infinity = inf ? (inf[i] != 0) : 0;
  // marshalling rule of the C ABI: a set flag means G1Affine::identity() (x=0,y=1; src/g1.rs:187-193),
  // whatever the coordinate bytes hold — the reference's invariant for identity-flagged points.
  if (p.infinity) p = g1a_identity();
  return p;
}
inline G2Affine load_g2a(const uint64_t *xy, const uint8_t *inf, size_t i) {
  G2Affine p;
  std::memcpy(&p.x, xy + 24 * i, 192);
  p.infinity = inf ? (inf[i] != 0) : 0;
  if (p.infinity) p = g2a_identity();
  return p;
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------- field tower, batched
// level: 1=Fp 2=Fp2 6=Fp6 12=Fp12 ; op: 0 mul 1 add 2 sub 3 square 4 neg 5 invert 6 frobenius 7 conjugate
// 8 mul_by_nonresidue 9 cyclotomic_square(Fp12)
int orc_tower_op(int level, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    if (level == 1) {
      Fp x{}, y{}, r{};
      std::memcpy(&x, a + 6 * i, 48);
      if (b) std::memcpy(&y, b + 6 * i, 48);
      switch (op) {
        case 0: r = fp_mul(x, y); break;
        case 1: r = fp_add(x, y); break;
        case 2: r = fp_sub(x, y); break;
        case 3: r = fp_square(x); break;
        case 4: r = fp_neg(x); break;
        case 5: r = fp_invert(x); break;
        default: return -1;
      }
      std::memcpy(out + 6 * i, &r, 48);
    } else if (level == 2) {
      Fp2 x{}, y{}, r{};
      std::memcpy(&x, a + 12 * i, 96);
      if (b) std::memcpy(&y, b + 12 * i, 96);
      switch (op) {
        case 0: r = fp2_mul(x, y); break;
        case 1: r = fp2_add(x, y); break;
        case 2: r = fp2_sub(x, y); break;
        case 3: r = fp2_square(x); break;
        case 4: r = fp2_neg(x); break;
        case 5: r = fp2_invert(x); break;
        case 6: r = fp2_frobenius_map(x); break;
        case 7: r = fp2_conjugate(x); break;
        case 8: r = fp2_mul_by_nonresidue(x); break;
        default: return -1;
      }
      std::memcpy(out + 12 * i, &r, 96);
    } else if (level == 6) {
      Fp6 x{}, y{}, r{};
      std::memcpy(&x, a + 36 * i, 288);
      if (b) std::memcpy(&y, b + 36 * i, 288);
      switch (op) {
        case 0: r = fp6_mul(x, y); break;
        case 1: r = fp6_add(x, y); break;
        case 2: r = fp6_sub(x, y); break;
        case 3: r = fp6_square(x); break;
        case 4: r = fp6_neg(x); break;
        case 5: r = fp6_invert(x); break;
        case 6: r = fp6_frobenius_map(x); break;
        case 8: r = fp6_mul_by_nonresidue(x); break;
        default: return -1;
      }
      std::memcpy(out + 36 * i, &r, 288);
    } else if (level == 12) {
      Fp12 x{}, y{}, r{};
      std::memcpy(&x, a + 72 * i, 576);
      if (b) std::memcpy(&y, b + 72 * i, 576);
      switch (op) {
        case 0: r = fp12_mul(x, y); break;
        case 3: r = fp12_square(x); break;
        case 5: r = fp12_invert(x); break;
        case 6: r = fp12_frobenius_map(x); break;
        case 7: r = fp12_conjugate(x); break;
        case 9: r = cyclotomic_square(x); break;
        default: return -1;
      }
      std::memcpy(out + 72 * i, &r, 576);
    } else
      return -1;
  }
  return 0;
}
// f.mul_by_014(c0,c1,c4)  (src/fp12.rs:116)
void orc_fp12_mul_by_014(const uint64_t *f, const uint64_t *c0, const uint64_t *c1, const uint64_t *c4,
                         uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Fp12 x;
    Fp2 a, b, c;
    std::memcpy(&x, f + 72 * i, 576);
    std::memcpy(&a, c0 + 12 * i, 96);
    std::memcpy(&b, c1 + 12 * i, 96);
    std::memcpy(&c, c4 + 12 * i, 96);
    Fp12 r = fp12_mul_by_014(x, a, b, c);
    std::memcpy(out + 72 * i, &r, 576);
  }
}
int orc_fp_sqrt(const uint64_t *a, uint64_t *out) {
  Fp x, r;
  std::memcpy(&x, a, 48);
  bool ok = fp_sqrt(x, r);
  std::memcpy(out, &r, 48);
  return ok;
}
int orc_fp2_sqrt(const uint64_t *a, uint64_t *out) {
  Fp2 x, r;
  std::memcpy(&x, a, 96);
  bool ok = fp2_sqrt(x, r);
  std::memcpy(out, &r, 96);
  return ok;
}
int orc_fp_from_bytes(const uint8_t *b, uint64_t *out) {
  Fp r;
  bool ok = fp_from_bytes(b, r);
  std::memcpy(out, &r, 48);
  return ok;
}
void orc_fp_to_bytes(const uint64_t *a, uint8_t *out) {
  Fp x;
  std::memcpy(&x, a, 48);
  fp_to_bytes(x, out);
}
int orc_fp_lex_largest(const uint64_t *a) {
  Fp x;
  std::memcpy(&x, a, 48);
  return fp_lexicographically_largest(x);
}

// ---------------------------------------------------------------- scalars
// Scalar::from_bytes_wide(64 B).to_bytes()  (src/scalar.rs:300, :284)
void orc_scalar_from_wide(const uint8_t *wide, uint8_t *out32, size_t n) {
  for (size_t i = 0; i < n; i++) fr_to_bytes(fr_from_bytes_wide(wide + 64 * i), out32 + 32 * i);
}
// Montgomery limbs -> canonical bytes (Scalar::to_bytes)
void orc_scalar_to_bytes(const uint64_t *mont, uint8_t *out32, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Scalar s;
    std::memcpy(&s, mont + 4 * i, 32);
    fr_to_bytes(s, out32 + 32 * i);
  }
}

// ---------------------------------------------------------------- scalar field
// op: 0 mul 1 add 2 sub 3 square 4 neg 5 invert (0 -> 0) 11 double; Montgomery limbs (4 x u64) in and out
int orc_fr_op(int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int threads) {
  if (op != 0 && op != 1 && op != 2 && op != 3 && op != 4 && op != 5 && op != 11) return -1;
  parallel_for(n, threads, [&](size_t i) {
    Scalar x{}, y{}, r{};
    std::memcpy(&x, a + 4 * i, 32);
    if (b) std::memcpy(&y, b + 4 * i, 32);
    switch (op) {
      case 0: r = fr_mul(x, y); break;
      case 1: r = fr_add(x, y); break;
      case 2: r = fr_sub(x, y); break;
      case 3: r = fr_square(x); break;
      case 4: r = fr_neg(x); break;
      case 5: r = fr_invert(x); break;
      default: r = fr_double(x); break;
    }
    std::memcpy(out + 4 * i, &r, 32);
  });
  return 0;
}
// Scalar::from_bytes: ok[i] = 1 when canonical; the limbs are written either way (like the CtOption's inner value)
void orc_fr_from_bytes(const uint8_t *in32, uint64_t *mont, uint8_t *ok, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Scalar s{};
    ok[i] = fr_from_bytes(in32 + 32 * i, s) ? 1 : 0;
    std::memcpy(mont + 4 * i, &s, 32);
  }
}
void orc_fr_pow(const uint64_t *a, const uint64_t *by, uint64_t *out) {
  Scalar x{};
  std::memcpy(&x, a, 32);
  Scalar r = fr_pow_vartime(x, by);
  std::memcpy(out, &r, 32);
}
void orc_fr_const(int which, uint64_t *out) {  // 0 one, 1 two_inv, 2 root_of_unity, 3 root_of_unity_inv, 4 generator
  const Scalar *c[] = {&FR_ONE, &FR_TWO_INV, &FR_ROOT_OF_UNITY, &FR_ROOT_OF_UNITY_INV, &FR_GENERATOR};
  std::memcpy(out, c[which], 32);
}
// out[k] = sum_j a[j] w^(jk) by the O(n^2) definition, w = ROOT_OF_UNITY^(2^(32 - log_n)) (its inverse when `inverse`;
// the inverse transform also divides by n).  coset: forward evaluates on g*<w> (a[j] *= g^j first), inverse undoes it.
void orc_fr_dft_naive(const uint64_t *a, int log_n, int inverse, int coset, uint64_t *out) {
  size_t n = (size_t)1 << log_n;
  std::vector<Scalar> x(n), y(n);
  std::memcpy(x.data(), a, 32 * n);
  Scalar g = FR_GENERATOR;
  if (coset && !inverse) {
    Scalar p = FR_ONE;
    for (size_t j = 0; j < n; j++) {
      x[j] = fr_mul(x[j], p);
      p = fr_mul(p, g);
    }
  }
  fr_dft_naive(x.data(), y.data(), n, fr_omega(log_n, inverse != 0));
  if (inverse) {
    Scalar ninv = FR_ONE;
    for (int i = 0; i < log_n; i++) ninv = fr_mul(ninv, FR_TWO_INV);
    Scalar ginv = fr_invert(g), p = FR_ONE;
    for (size_t j = 0; j < n; j++) {
      y[j] = fr_mul(y[j], ninv);
      if (coset) {
        y[j] = fr_mul(y[j], p);
        p = fr_mul(p, ginv);
      }
    }
  }
  std::memcpy(out, y.data(), 32 * n);
}
// the same transform in O(n log n): bit-reversal permutation, then log_n butterfly stages (decimation in time);
// stage t pairs (i, i + 2^t) with twiddle w^((i mod 2^t) * n / 2^(t+1)).  The CPU baseline of bench.py's fr_ntt
// workload and the oracle for sizes the definition above is too slow for.
void orc_fr_ntt(const uint64_t *a, int log_n, int inverse, int coset, uint64_t *out, int threads) {
  size_t n = (size_t)1 << log_n;
  std::vector<Scalar> x(n);
  Scalar g = FR_GENERATOR;
  {
    const Scalar *in = reinterpret_cast<const Scalar *>(a);
    std::vector<Scalar> gp;
    if (coset && !inverse) {
      gp.resize(n);
      Scalar p = FR_ONE;
      for (size_t j = 0; j < n; j++) {
        gp[j] = p;
        p = fr_mul(p, g);
      }
    }
    for (size_t i = 0; i < n; i++) {
      size_t r = 0;
      for (int b = 0; b < log_n; b++) r |= ((i >> b) & 1) << (log_n - 1 - b);
      Scalar v;
      std::memcpy(&v, &in[r], 32);
      x[i] = gp.empty() ? v : fr_mul(v, gp[r]);
    }
  }
  Scalar w = fr_omega(log_n, inverse != 0);
  std::vector<Scalar> tw(n > 1 ? n / 2 : 1);
  tw[0] = FR_ONE;
  for (size_t e = 1; e < n / 2; e++) tw[e] = fr_mul(tw[e - 1], w);
  for (int t = 0; t < log_n; t++) {
    size_t half = (size_t)1 << t, step = n >> (t + 1);
    parallel_for(n / 2, threads, [&](size_t k) {
      size_t lo = k & (half - 1), i = ((k >> t) << (t + 1)) | lo;
      Scalar u = x[i], v = fr_mul(x[i + half], tw[lo * step]);
      x[i] = fr_add(u, v);
      x[i + half] = fr_sub(u, v);
    });
  }
  if (inverse) {
    Scalar ninv = FR_ONE;
    for (int i = 0; i < log_n; i++) ninv = fr_mul(ninv, FR_TWO_INV);
    Scalar ginv = fr_invert(g), p = FR_ONE;
    for (size_t j = 0; j < n; j++) {
      x[j] = fr_mul(x[j], ninv);
      if (coset) {
        x[j] = fr_mul(x[j], p);
        p = fr_mul(p, ginv);
      }
    }
  }
  std::memcpy(out, x.data(), 32 * n);
}

// ---------------------------------------------------------------- Gt * Scalar (src/pairings.rs:296-323)
// double-and-add in Fp12, MSB first over the 32-byte LE scalar, leading bit skipped, acc + self computed at every bit and
// selected — as the reference does
void orc_gt_mul(const uint64_t *g, const uint8_t *scalars, uint64_t *out, size_t n, int threads) {
  parallel_for(n, threads, [&](size_t i) {
    Fp12 x{}, acc = fp12_one();
    std::memcpy(&x, g + 72 * i, 576);
    const uint8_t *by = scalars + 32 * i;
    bool first = true;
    for (int b = 31; b >= 0; b--)
      for (int k = 7; k >= 0; k--) {
        if (first) {
          first = false;
          continue;
        }
        acc = fp12_square(acc);
        Fp12 t = fp12_mul(acc, x);
        if ((by[b] >> k) & 1) acc = t;
      }
    std::memcpy(out + 72 * i, &acc, 576);
  });
}

// ---------------------------------------------------------------- hash to curve
int orc_expand_message_xmd_sha256(const uint8_t *msg, size_t msg_len, const uint8_t *dst, size_t dst_len, size_t len_in_bytes,
                                  uint8_t *out) {
  return expand_message_xmd_sha256(msg, msg_len, dst, dst_len, len_in_bytes, out) ? 0 : -1;
}
void orc_sha256(const uint8_t *msg, size_t n, uint8_t *out) {
  Sha256 h;
  h.update(msg, n);
  h.finish(out);
}
// msgs: concatenated messages, message i = msgs[off[i] .. off[i+1]); out: n projective points
int orc_g1_hash(const uint8_t *msgs, const uint64_t *off, size_t n, const uint8_t *dst, size_t dst_len, int encode, uint64_t *out,
                int threads) {
  std::atomic<int> bad{0};
  parallel_for(n, threads, [&](size_t i) {
    G1Projective r{};
    if (!g1_hash(msgs + off[i], (size_t)(off[i + 1] - off[i]), dst, dst_len, encode != 0, r)) bad = 1;
    std::memcpy(out + 18 * i, &r, 144);
  });
  return bad ? -1 : 0;
}
int orc_g2_hash(const uint8_t *msgs, const uint64_t *off, size_t n, const uint8_t *dst, size_t dst_len, int encode, uint64_t *out,
                int threads) {
  std::atomic<int> bad{0};
  parallel_for(n, threads, [&](size_t i) {
    G2Projective r{};
    if (!g2_hash(msgs + off[i], (size_t)(off[i + 1] - off[i]), dst, dst_len, encode != 0, r)) bad = 1;
    std::memcpy(out + 36 * i, &r, 288);
  });
  return bad ? -1 : 0;
}
// the stages on their own.  kind: 0 g1 sswu (Fp -> E' point), 1 g1 iso_map, 2 g1 map_to_curve (Fp -> E), 3 g1
// clear_cofactor, 4..7 the same for g2 (Fp2).  in: n field elements or n projective points; out: n projective points
int orc_h2c_stage(int kind, const uint64_t *in, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    if (kind < 4) {
      G1Projective r{}, p{};
      Fp u{};
      if (kind == 0 || kind == 2) std::memcpy(&u, in + 6 * i, 48); else std::memcpy(&p, in + 18 * i, 144);
      r = kind == 0 ? g1_map_to_curve_simple_swu(u) : kind == 1 ? g1_iso_map(p) : kind == 2 ? g1_map_to_curve(u) : g1p_clear_cofactor(p);
      std::memcpy(out + 18 * i, &r, 144);
    } else if (kind < 8) {
      G2Projective r{}, p{};
      Fp2 u{};
      if (kind == 4 || kind == 6) std::memcpy(&u, in + 12 * i, 96); else std::memcpy(&p, in + 36 * i, 288);
      r = kind == 4 ? g2_map_to_curve_simple_swu(u) : kind == 5 ? g2_iso_map(p) : kind == 6 ? g2_map_to_curve(u) : g2p_clear_cofactor(p);
      std::memcpy(out + 36 * i, &r, 288);
    } else {
      return -1;
    }
  }
  return 0;
}
// hash_to_field: 64 uniform bytes per Fp (src/hash_to_curve/map_g1.rs:513) ; sgn0 (:535, map_g2.rs:382)
void orc_fp_from_okm(const uint8_t *okm, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Fp r = fp_from_okm(okm + 64 * i);
    std::memcpy(out + 6 * i, &r, 48);
  }
}
void orc_sgn0(int level, const uint64_t *a, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    if (level == 1) {
      Fp x{};
      std::memcpy(&x, a + 6 * i, 48);
      out[i] = fp_sgn0(x);
    } else {
      Fp2 x{};
      std::memcpy(&x, a + 12 * i, 96);
      out[i] = fp2_sgn0(x);
    }
  }
}

// Scalar::from_okm (src/hash_to_curve/map_scalar.rs:17) on n x 48 bytes, and Scalar::hash_to_field: `count` scalars per
// message from expand_message_xmd(msg, dst, 48 * count)  (src/hash_to_curve/mod.rs:41-66)
void orc_fr_from_okm(const uint8_t *okm, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Scalar r = fr_from_okm(okm + 48 * i);
    std::memcpy(out + 4 * i, &r, 32);
  }
}
int orc_fr_hash_to_field(const uint8_t *msgs, const uint64_t *off, size_t n, const uint8_t *dst, size_t dst_len, int count,
                         uint64_t *out) {
  std::vector<uint8_t> okm((size_t)48 * count);
  for (size_t i = 0; i < n; i++) {
    if (!expand_message_xmd_sha256(msgs + off[i], (size_t)(off[i + 1] - off[i]), dst, dst_len, okm.size(), okm.data())) return -1;
    for (int c = 0; c < count; c++) {
      Scalar r = fr_from_okm(okm.data() + 48 * c);
      std::memcpy(out + 4 * (i * count + c), &r, 32);
    }
  }
  return 0;
}

// ---------------------------------------------------------------- G1
void orc_g1_generator(uint64_t *proj) {
  G1Projective g = g1p_generator();
  std::memcpy(proj, &g, 144);
}
void orc_g1_double(const uint64_t *p, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    G1Projective a;
    std::memcpy(&a, p + 18 * i, 144);
    a = g1p_double(a);
    std::memcpy(out + 18 * i, &a, 144);
  }
}
void orc_g1_add(const uint64_t *p, const uint64_t *q, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    G1Projective a, b;
    std::memcpy(&a, p + 18 * i, 144);
    std::memcpy(&b, q + 18 * i, 144);
    a = g1p_add(a, b);
    std::memcpy(out + 18 * i, &a, 144);
  }
}
void orc_g1_add_mixed(const uint64_t *p, const uint64_t *qxy, const uint8_t *qinf, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    G1Projective a;
    std::memcpy(&a, p + 18 * i, 144);
    a = g1p_add_mixed(a, load_g1a(qxy, qinf, i));
    std::memcpy(out + 18 * i, &a, 144);
  }
}
// out[i] = p[i] * s[i]   — G1Projective::multiply (src/g1.rs:754), raw limbs
void orc_g1_mul(const uint64_t *p, const uint8_t *s, uint64_t *out, size_t n, int threads) {
  parallel_for(n, threads, [&](size_t i) {
    G1Projective a;
    std::memcpy(&a, p + 18 * i, 144);
    a = g1p_multiply(a, s + 32 * i);
    std::memcpy(out + 18 * i, &a, 144);
  });
}
void orc_g1_to_affine(const uint64_t *p, uint64_t *oxy, uint8_t *oinf, size_t n) {
  for (size_t i = 0; i < n; i++) {
    G1Projective a;
    std::memcpy(&a, p + 18 * i, 144);
    G1Affine q = g1a_from_projective(a);
    std::memcpy(oxy + 12 * i, &q.x, 96);
    oinf[i] = q.infinity;
  }
}
void orc_g1_batch_normalize(const uint64_t *p, uint64_t *oxy, uint8_t *oinf, size_t n) {
  std::vector<G1Projective> in(n);
  std::vector<G1Affine> out(n);
  if (n) std::memcpy(in.data(), p, 144 * n);
  g1p_batch_normalize(in.data(), out.data(), n);
  for (size_t i = 0; i < n; i++) {
    std::memcpy(oxy + 12 * i, &out[i].x, 96);
    oinf[i] = out[i].infinity;
  }
}
// "MSM" as the reference API expresses it: sum_i (affine p_i lifted) * s_i, folded with add
void orc_g1_msm_naive(const uint64_t *xy, const uint8_t *inf, const uint8_t *s, size_t n, uint64_t *out, int threads) {
  if (threads < 1) threads = 1;
  std::vector<G1Projective> part((size_t)threads, g1p_identity());
  std::atomic<size_t> next{0};
  auto work = [&](int t) {
    G1Projective acc = g1p_identity();
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= n) break;
      acc = g1p_add(acc, g1p_multiply(g1p_from_affine(load_g1a(xy, inf, i)), s + 32 * i));
    }
    part[t] = acc;
  };
  if (threads == 1)
    work(0);
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(work, t);
    for (auto &th : pool) th.join();
  }
  G1Projective acc = g1p_identity();
  for (auto &p : part) acc = g1p_add(acc, p);
  std::memcpy(out, &acc, 144);
}
// CPU Pippenger built ONLY from the restated reference group ops; cross-validated against
// orc_g1_msm_naive in tests, used for full-size (2^20) parity checks.  c-bit unsigned windows.
void orc_g1_msm_pippenger(const uint64_t *xy, const uint8_t *inf, const uint8_t *s, size_t n, uint64_t *out, int c,
                          int threads) {
  int nwin = (255 + c - 1) / c;
  std::vector<G1Projective> wsum((size_t)nwin, g1p_identity());
  parallel_for((size_t)nwin, threads, [&](size_t w) {
    std::vector<G1Projective> bucket((size_t)1 << c, g1p_identity());
    for (size_t i = 0; i < n; i++) {
      unsigned d = 0;
      for (int b = 0; b < c; b++) {
        int bit = (int)w * c + b;
        if (bit < 256) d |= (unsigned)((s[32 * i + bit / 8] >> (bit % 8)) & 1) << b;
      }
      if (d) bucket[d] = g1p_add_mixed(bucket[d], load_g1a(xy, inf, i));
    }
    G1Projective run = g1p_identity(), acc = g1p_identity();
    for (size_t d = ((size_t)1 << c) - 1; d >= 1; d--) {
      run = g1p_add(run, bucket[d]);
      acc = g1p_add(acc, run);
    }
    wsum[w] = acc;
  });
  G1Projective acc = g1p_identity();
  for (int w = nwin - 1; w >= 0; w--) {
    for (int k = 0; k < c; k++) acc = g1p_double(acc);
    acc = g1p_add(acc, wsum[w]);
  }
  std::memcpy(out, &acc, 144);
}
int orc_g1_to_compressed(const uint64_t *xy, int inf, uint8_t *out) {
  g1a_to_compressed(inf ? g1a_identity() : load_g1a(xy, nullptr, 0), out);
  return 0;
}
int orc_g1_to_uncompressed(const uint64_t *xy, int inf, uint8_t *out) {
  g1a_to_uncompressed(inf ? g1a_identity() : load_g1a(xy, nullptr, 0), out);
  return 0;
}
int orc_g1_from_compressed(const uint8_t *in, uint64_t *xy, uint8_t *inf) {
  G1Affine p;
  bool ok = g1a_from_compressed(in, p);
  if (ok) {
    std::memcpy(xy, &p.x, 96);
    *inf = p.infinity;
  }
  return ok;
}
int orc_g1_from_uncompressed(const uint8_t *in, uint64_t *xy, uint8_t *inf) {
  G1Affine p;
  bool ok = g1a_from_uncompressed(in, p);
  if (ok) {
    std::memcpy(xy, &p.x, 96);
    *inf = p.infinity;
  }
  return ok;
}

// ---------------------------------------------------------------- G2
void orc_g2_generator(uint64_t *proj) {
  G2Projective g = g2p_generator();
  std::memcpy(proj, &g, 288);
}
void orc_g2_double(const uint64_t *p, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    G2Projective a;
    std::memcpy(&a, p + 36 * i, 288);
    a = g2p_double(a);
    std::memcpy(out + 36 * i, &a, 288);
  }
}
void orc_g2_add(const uint64_t *p, const uint64_t *q, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    G2Projective a, b;
    std::memcpy(&a, p + 36 * i, 288);
    std::memcpy(&b, q + 36 * i, 288);
    a = g2p_add(a, b);
    std::memcpy(out + 36 * i, &a, 288);
  }
}
void orc_g2_add_mixed(const uint64_t *p, const uint64_t *qxy, const uint8_t *qinf, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    G2Projective a;
    std::memcpy(&a, p + 36 * i, 288);
    a = g2p_add_mixed(a, load_g2a(qxy, qinf, i));
    std::memcpy(out + 36 * i, &a, 288);
  }
}
void orc_g2_mul(const uint64_t *p, const uint8_t *s, uint64_t *out, size_t n, int threads) {
  parallel_for(n, threads, [&](size_t i) {
    G2Projective a;
    std::memcpy(&a, p + 36 * i, 288);
    a = g2p_multiply(a, s + 32 * i);
    std::memcpy(out + 36 * i, &a, 288);
  });
}
void orc_g2_to_affine(const uint64_t *p, uint64_t *oxy, uint8_t *oinf, size_t n) {
  for (size_t i = 0; i < n; i++) {
    G2Projective a;
    std::memcpy(&a, p + 36 * i, 288);
    G2Affine q = g2a_from_projective(a);
    std::memcpy(oxy + 24 * i, &q.x, 192);
    oinf[i] = q.infinity;
  }
}
void orc_g2_batch_normalize(const uint64_t *p, uint64_t *oxy, uint8_t *oinf, size_t n) {
  std::vector<G2Projective> in(n);
  std::vector<G2Affine> out(n);
  if (n) std::memcpy(in.data(), p, 288 * n);
  g2p_batch_normalize(in.data(), out.data(), n);
  for (size_t i = 0; i < n; i++) {
    std::memcpy(oxy + 24 * i, &out[i].x, 192);
    oinf[i] = out[i].infinity;
  }
}
void orc_g2_msm_naive(const uint64_t *xy, const uint8_t *inf, const uint8_t *s, size_t n, uint64_t *out, int threads) {
  if (threads < 1) threads = 1;
  std::vector<G2Projective> part((size_t)threads, g2p_identity());
  std::atomic<size_t> next{0};
  auto work = [&](int t) {
    G2Projective acc = g2p_identity();
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= n) break;
      acc = g2p_add(acc, g2p_multiply(g2p_from_affine(load_g2a(xy, inf, i)), s + 32 * i));
    }
    part[t] = acc;
  };
  if (threads == 1)
    work(0);
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(work, t);
    for (auto &th : pool) th.join();
  }
  G2Projective acc = g2p_identity();
  for (auto &p : part) acc = g2p_add(acc, p);
  std::memcpy(out, &acc, 288);
}
void orc_g2_msm_pippenger(const uint64_t *xy, const uint8_t *inf, const uint8_t *s, size_t n, uint64_t *out, int c,
                          int threads) {
  int nwin = (255 + c - 1) / c;
  std::vector<G2Projective> wsum((size_t)nwin, g2p_identity());
  parallel_for((size_t)nwin, threads, [&](size_t w) {
    std::vector<G2Projective> bucket((size_t)1 << c, g2p_identity());
    for (size_t i = 0; i < n; i++) {
      unsigned d = 0;
      for (int b = 0; b < c; b++) {
        int bit = (int)w * c + b;
        if (bit < 256) d |= (unsigned)((s[32 * i + bit / 8] >> (bit % 8)) & 1) << b;
      }
      if (d) bucket[d] = g2p_add_mixed(bucket[d], load_g2a(xy, inf, i));
    }
    G2Projective run = g2p_identity(), acc = g2p_identity();
    for (size_t d = ((size_t)1 << c) - 1; d >= 1; d--) {
      run = g2p_add(run, bucket[d]);
      acc = g2p_add(acc, run);
    }
    wsum[w] = acc;
  });
  G2Projective acc = g2p_identity();
  for (int w = nwin - 1; w >= 0; w--) {
    for (int k = 0; k < c; k++) acc = g2p_double(acc);
    acc = g2p_add(acc, wsum[w]);
  }
  std::memcpy(out, &acc, 288);
}
int orc_g2_to_compressed(const uint64_t *xy, int inf, uint8_t *out) {
  g2a_to_compressed(inf ? g2a_identity() : load_g2a(xy, nullptr, 0), out);
  return 0;
}
int orc_g2_to_uncompressed(const uint64_t *xy, int inf, uint8_t *out) {
  g2a_to_uncompressed(inf ? g2a_identity() : load_g2a(xy, nullptr, 0), out);
  return 0;
}
int orc_g2_from_compressed(const uint8_t *in, uint64_t *xy, uint8_t *inf) {
  G2Affine p;
  bool ok = g2a_from_compressed(in, p);
  if (ok) {
    std::memcpy(xy, &p.x, 192);
    *inf = p.infinity;
  }
  return ok;
}
int orc_g2_from_uncompressed(const uint8_t *in, uint64_t *xy, uint8_t *inf) {
  G2Affine p;
  bool ok = g2a_from_uncompressed(in, p);
  if (ok) {
    std::memcpy(xy, &p.x, 192);
    *inf = p.infinity;
  }
  return ok;
}

// is_torsion_free / is_on_curve per point (src/g1.rs:401-418, src/g2.rs:475-491): out[i] bit0 on curve, bit1 torsion free
void orc_g1_checks(const uint64_t *xy, const uint8_t *inf, size_t n, uint8_t *out) {
  for (size_t i = 0; i < n; i++) {
    G1Affine p = load_g1a(xy, inf, i);
    out[i] = (g1a_is_on_curve(p) ? 1 : 0) | (g1a_is_torsion_free(p) ? 2 : 0);
  }
}
void orc_g2_checks(const uint64_t *xy, const uint8_t *inf, size_t n, uint8_t *out) {
  for (size_t i = 0; i < n; i++) {
    G2Affine p = load_g2a(xy, inf, i);
    out[i] = (g2a_is_on_curve(p) ? 1 : 0) | (g2a_is_torsion_free(p) ? 2 : 0);
  }
}

// ---------------------------------------------------------------- pairings
// out[i] = MillerLoopResult of (p_i, q_i), unprepared, identity handling as pairing() :636-651
void orc_miller_loop(const uint64_t *pxy, const uint8_t *pinf, const uint64_t *qxy, const uint8_t *qinf, size_t n,
                     uint64_t *out, int threads) {
  parallel_for(n, threads, [&](size_t i) {
    Fp12 f = miller_loop_pair(load_g1a(pxy, pinf, i), load_g2a(qxy, qinf, i));
    std::memcpy(out + 72 * i, &f, 576);
  });
}
void orc_final_exp(const uint64_t *in, size_t n, uint64_t *out, int threads) {
  parallel_for(n, threads, [&](size_t i) {
    Fp12 f;
    std::memcpy(&f, in + 72 * i, 576);
    f = final_exponentiation(f);
    std::memcpy(out + 72 * i, &f, 576);
  });
}
void orc_pairing(const uint64_t *pxy, const uint8_t *pinf, const uint64_t *qxy, const uint8_t *qinf, size_t n,
                 uint64_t *out, int threads) {
  parallel_for(n, threads, [&](size_t i) {
    Fp12 f = pairing(load_g1a(pxy, pinf, i), load_g2a(qxy, qinf, i));
    std::memcpy(out + 72 * i, &f, 576);
  });
}
// multi_miller_loop(&[(p_i, G2Prepared::from(q_i))]) -> one MillerLoopResult  (src/pairings.rs:554)
void orc_multi_miller_loop(const uint64_t *pxy, const uint8_t *pinf, const uint64_t *qxy, const uint8_t *qinf, size_t n,
                           uint64_t *out) {
  std::vector<G1Affine> ps(n);
  std::vector<G2Prepared> qs(n);
  for (size_t i = 0; i < n; i++) {
    ps[i] = load_g1a(pxy, pinf, i);
    qs[i] = g2_prepare(load_g2a(qxy, qinf, i));
  }
  Fp12 f = multi_miller_loop(ps.data(), qs.data(), n);
  std::memcpy(out, &f, 576);
}
// G2Prepared::from(q): 68 coefficient triples (3 x Fp2 each), returns count
int orc_g2_prepare(const uint64_t *qxy, int qinf, uint64_t *coeffs_out) {
  uint8_t inf = (uint8_t)qinf;
  G2Prepared pr = g2_prepare(load_g2a(qxy, &inf, 0));
  for (size_t k = 0; k < pr.coeffs.size(); k++) std::memcpy(coeffs_out + 36 * k, &pr.coeffs[k], 288);
  return (int)pr.coeffs.size();
}
int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }
}
