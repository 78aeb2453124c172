// C ABI, part 4 (the callers either side of the hot path): batched point
// (de)serialization on the device.
//   serialize   = G{1,2}Affine::to_compressed / to_uncompressed       src/g1.rs:221-260, src/g2.rs:254-299
//   deserialize = G{1,2}Affine::from_{un,}compressed_unchecked + is_on_curve
//                                                                     src/g1.rs:275-390 :414, src/g2.rs:313-464 :487
// Wire format: src/notes/serialization.rs:1-29 — Fp big-endian 48 bytes (Fp::to_bytes src/fp.rs:211-227,
// from_bytes :179-208 with its canonicity check), Fp2 as c1 || c0, flag bits 7/6/5 of byte 0 =
// compressed / infinity / lexicographically-largest-y.  The subgroup check (is_torsion_free) is NOT part of
// the deserialize entry points; it is offered separately (b200_g{1,2}_check = is_on_curve + is_torsion_free,
// src/g1.rs:401-418, src/g2.rs:475-491), so deserialize + check == the reference's checked constructors.
#include "constants.cuh"
#include "ctx.cuh"
#include "curve.cuh"

using namespace b200;

namespace {

// Montgomery -> canonical integer: a * 1 * R^-1   (== montgomery_reduce(a, 0), src/fp.rs:214-217)
B200_DEV fp fp_from_mont(const fp &a) {
  fp one = fp_zero();
  one.v[0] = 1u;
  return fp_mul_c(a, one);
}
B200_DEV fp fp_to_mont(const fp &a) { return fp_mul_c(a, fp_const(K_R2)); }  // src/fp.rs:205
// canonical integer -> 48 big-endian bytes (16-byte aligned destination)
B200_DEV void fp_store_be(uint8_t *dst, const fp &c, uint32_t or_top_bits = 0) {
  uint32_t w[12];
#pragma unroll
  for (int k = 0; k < 12; k++) w[k] = __byte_perm(c.v[11 - k], 0, 0x0123);
  w[0] |= or_top_bits;  // byte 0 is the low byte of word 0 after the swap
  uint4 *q = reinterpret_cast<uint4 *>(dst);
  q[0] = make_uint4(w[0], w[1], w[2], w[3]);
  q[1] = make_uint4(w[4], w[5], w[6], w[7]);
  q[2] = make_uint4(w[8], w[9], w[10], w[11]);
}
B200_DEV fp fp_load_be(const uint8_t *src, uint32_t and_top_mask = 0xffffffffu) {
  const uint4 *q = reinterpret_cast<const uint4 *>(src);
  uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  uint32_t w[12] = {a.x & and_top_mask, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
  fp r;
#pragma unroll
  for (int k = 0; k < 12; k++) r.v[11 - k] = __byte_perm(w[k], 0, 0x0123);
  return r;
}
// integer compare: a >= k  (12 words)
B200_DEV bool words_ge(const fp &a, const uint32_t *k) {
  uint32_t t, borrow;
  ptx_sub_cc(t, a.v[0], k[0]);
#pragma unroll
  for (int i = 1; i < 12; i++) ptx_subc_cc(t, a.v[i], k[i]);
  ptx_subc(borrow, 0u, 0u);
  return borrow == 0;
}
// src/fp.rs:273-299 on a Montgomery-form element
B200_DEV bool fp_lex_largest(const fp &a) { return words_ge(fp_from_mont(a), K_HALF_P_PLUS1); }
B200_DEV bool fp2_lex_largest(const fp2 &a) {  // src/fp2.rs:171-180
  return fp_lex_largest(a.c1) || (fp_is_zero(a.c1) && fp_lex_largest(a.c0));
}
// square-and-multiply, MSB first, exponent = 12 plain words (src/fp.rs:309-321 / src/fp2.rs:322-336)
B200_DEV fp fp_pow(const fp &a, const uint32_t *e) {
  fp res = fp_one();
#pragma unroll 1
  for (int w = 11; w >= 0; w--) {
    uint32_t ew = e[w];
#pragma unroll 1
    for (int i = 31; i >= 0; i--) {
      res = fp_mul_c(res, res);
      if ((ew >> i) & 1) res = fp_mul_c(res, a);
    }
  }
  return res;
}
B200_DEV fp2 fp2_pow(const fp2 &a, const uint32_t *e) {
  fp2 res = fp2_one();
#pragma unroll 1
  for (int w = 11; w >= 0; w--) {
    uint32_t ew = e[w];
#pragma unroll 1
    for (int i = 31; i >= 0; i--) {
      res = S2(res);
      if ((ew >> i) & 1) res = M2(res, a);
    }
  }
  return res;
}
// src/fp.rs:324-343
B200_DEV bool fp_sqrt(const fp &a, fp &out) {
  out = fp_pow(a, K_EXP_SQRT);
  return fp_eq(fp_mul_c(out, out), a);
}
// src/fp2.rs:245-298 (Alg. 9 of eprint 2012/685)
B200_DEV bool fp2_sqrt(const fp2 &a, fp2 &out) {
  if (fp2_is_zero(a)) {
    out = fp2_zero();
    return true;
  }
  fp2 a1 = fp2_pow(a, K_EXP_P34);
  fp2 alpha = M2(S2(a1), a);
  fp2 x0 = M2(a1, a);
  fp2 cand;
  if (fp2_eq(alpha, fp2_neg(fp2_one())))
    cand = fp2{fp_neg(x0.c1), x0.c0};
  else
    cand = M2(fp2_pow(fp2_add(alpha, fp2_one()), K_EXP_P12), x0);
  out = cand;
  return fp2_eq(S2(cand), a);
}

// field-generic glue -----------------------------------------------------------------------------
B200_DEV void coord_store(uint8_t *dst, const fp &a, uint32_t flags) { fp_store_be(dst, fp_from_mont(a), flags); }
B200_DEV void coord_store(uint8_t *dst, const fp2 &a, uint32_t flags) {  // c1 || c0
  fp_store_be(dst, fp_from_mont(a.c1), flags);
  fp_store_be(dst + 48, fp_from_mont(a.c0));
}
B200_DEV bool coord_lex(const fp &a) { return fp_lex_largest(a); }
B200_DEV bool coord_lex(const fp2 &a) { return fp2_lex_largest(a); }
// returns canonicity (value < p) of every component; `out` in Montgomery form
B200_DEV bool coord_load(const uint8_t *src, fp &out, bool mask_flags) {
  fp c = fp_load_be(src, mask_flags ? 0xffffff1fu : 0xffffffffu);  // byte 0 = low byte of word 0 before the swap
  bool ok = !words_ge(c, FP_MOD);
  out = fp_to_mont(c);
  return ok;
}
B200_DEV bool coord_load(const uint8_t *src, fp2 &out, bool mask_flags) {
  bool ok1 = coord_load(src, out.c1, mask_flags);
  bool ok0 = coord_load(src + 48, out.c0, false);
  return ok0 && ok1;
}
B200_DEV bool coord_sqrt(const fp &a, fp &out) { return fp_sqrt(a, out); }
B200_DEV bool coord_sqrt(const fp2 &a, fp2 &out) { return fp2_sqrt(a, out); }
template <class F> B200_DEV F curve_b();
template <> B200_DEV fp curve_b<fp>() {  // 4  (src/g1.rs:176-183)
  fp one = fp_one();
  return fp_dbl(fp_dbl(one));
}
template <> B200_DEV fp2 curve_b<fp2>() {  // 4(1+u)  (src/g2.rs:177-194)
  fp four = curve_b<fp>();
  return fp2{four, four};
}
template <class F>
B200_DEV bool on_curve(const F &x, const F &y) {  // y^2 - x^3 == b   (src/g1.rs:414-418)
  return f_eq(f_sub(f_sqr(y), f_mul(f_sqr(x), x)), curve_b<F>());
}

// ---- subgroup membership (the check the reference's checked constructors add: src/g1.rs:264-267, :336)
// [x]P for the BLS parameter x = -0xd201000000010000 (src/g1.rs:777-793, src/g2.rs:915-932)
template <class F>
B200_DEV proj<F> proj_mul_by_x(const proj<F> &s) {
  proj<F> xself = proj_identity<F>(), acc = s;
  unsigned long long x = 0xd201000000010000ull >> 1;
#pragma unroll 1
  while (x != 0) {
    acc = proj_double(acc);
    if (x & 1) xself = proj_add(xself, acc);
    x >>= 1;
  }
  return proj_neg(xself);
}
template <class F>
B200_DEV bool proj_eq(const proj<F> &a, const proj<F> &b) {  // src/g1.rs:479-496
  bool az = f_is_zero(a.z), bz = f_is_zero(b.z);
  if (az || bz) return az && bz;
  return f_eq(f_mul(a.x, b.z), f_mul(b.x, a.z)) && f_eq(f_mul(a.y, b.z), f_mul(b.y, a.z));
}
// G1: endomorphism(P) == -[x^2]P   (src/g1.rs:401-410, eprint 2021/1130 sec. 6)
B200_DEV bool torsion_free(const affine<fp> &p) {
  proj<fp> pp = proj_from_affine(p);
  proj<fp> m = proj_neg(proj_mul_by_x(proj_mul_by_x(pp)));
  proj<fp> e = pp;
  e.x = fp_mul_c(e.x, fp_const(K_G1_BETA_REF));
  return proj_eq(m, e);
}
// G2: psi(P) == [x]P   (src/g2.rs:475-482, psi :847-888)
B200_DEV bool torsion_free(const affine<fp2> &p) {
  proj<fp2> pp = proj_from_affine(p);
  fp2 cx = fp2{fp_zero(), fp_const(K_PSI_X_U)}, cy = fp2{fp_const(K_PSI_Y_R), fp_const(K_PSI_Y_U)};
  proj<fp2> psi{M2(fp2_conj(pp.x), cx), M2(fp2_conj(pp.y), cy), fp2_conj(pp.z)};
  return proj_eq(psi, proj_mul_by_x(pp));
}
// status bit 0: is_on_curve, bit 1: is_torsion_free
template <class F>
__global__ void __launch_bounds__(128) k_point_checks(const char *xy, const uint8_t *inf, size_t n, uint8_t *status) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  affine<F> p = affine_load<F>(xy, inf, i);
  bool oc = p.inf || on_curve(p.x, p.y);
  status[i] = (oc ? 1 : 0) | (torsion_free(p) ? 2 : 0);
}

template <class F>
__global__ void __launch_bounds__(128) k_serialize(const char *xy, const uint8_t *inf, size_t n, int compressed,
                                                 uint8_t *out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr size_t FB = field_traits<F>::bytes;
  affine<F> p = affine_load<F>(xy, inf, i);
  F x = p.inf ? field_traits<F>::zero() : p.x;
  if (compressed) {
    uint32_t flags = 0x80u | (p.inf ? 0x40u : 0u) | ((!p.inf && coord_lex(p.y)) ? 0x20u : 0u);
    coord_store(out + FB * i, x, flags);
  } else {
    F y = p.inf ? field_traits<F>::zero() : p.y;
    coord_store(out + 2 * FB * i, x, p.inf ? 0x40u : 0u);
    coord_store(out + 2 * FB * i + FB, y, 0);
  }
}

// status bit 0: the reference's from_*_unchecked would return Some; bit 1: is_on_curve
template <class F>
__global__ void __launch_bounds__(128) k_deserialize(const uint8_t *in, size_t n, int compressed, char *oxy, uint8_t *oinf,
                                                   uint8_t *status) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr size_t FB = field_traits<F>::bytes;
  const uint8_t *src = in + (compressed ? FB : 2 * FB) * i;
  uint8_t b0 = src[0];
  bool cflag = (b0 >> 7) & 1, iflag = (b0 >> 6) & 1, sflag = (b0 >> 5) & 1;
  F x, y;
  bool canon = coord_load(src, x, true);
  bool ok, curve;
  affine<F> p;
  if (compressed) {
    if (iflag) {
      ok = canon && cflag && !sflag && f_is_zero(x);
      p = affine_identity<F>();
      curve = true;
    } else {
      bool sq = coord_sqrt(f_add(f_mul(f_sqr(x), x), curve_b<F>()), y);
      if (coord_lex(y) != sflag) y = f_neg(y);
      ok = canon && sq && cflag;
      p = affine<F>{x, y, false};
      curve = sq;
    }
  } else {
    bool canon_y = coord_load(src + FB, y, false);
    if (iflag) {
      ok = canon && canon_y && f_is_zero(x) && f_is_zero(y) && !cflag && !sflag;
      p = affine_identity<F>();
      curve = true;
    } else {
      ok = canon && canon_y && !cflag && !sflag;
      p = affine<F>{x, y, false};
      curve = on_curve(x, y);
    }
  }
  if (!ok) p = affine_identity<F>();
  affine_store<F>(oxy, oinf, i, p);
  status[i] = (ok ? 1 : 0) | ((ok && curve) ? 2 : 0);
}

inline unsigned nblk(size_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

template <class F>
int serialize_host(b200_ctx *ctx, const void *xy, const uint8_t *inf, size_t n, int compressed, uint8_t *out) {
  constexpr size_t FB = field_traits<F>::bytes;
  size_t ob = (compressed ? FB : 2 * FB) * n;
  int rc = stage_reserve(ctx, 2 * FB * n + n + ob + 8 * 256);
  if (rc != B200_OK) return rc;
  void *dxy = stage_take(ctx, 2 * FB * n), *di = inf ? stage_take(ctx, n) : nullptr, *dout = stage_take(ctx, ob);
  B200_CUDA(ctx, cudaMemcpyAsync(dxy, xy, 2 * FB * n, cudaMemcpyHostToDevice, ctx->stream));
  if (inf) B200_CUDA(ctx, cudaMemcpyAsync(di, inf, n, cudaMemcpyHostToDevice, ctx->stream));
  B200_LAUNCH(ctx, k_serialize<F>, nblk(n, 128), 128, 0, (const char *)dxy, (const uint8_t *)di, n, compressed, (uint8_t *)dout);
  B200_CUDA(ctx, cudaMemcpyAsync(out, dout, ob, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
template <class F>
int deserialize_host(b200_ctx *ctx, const uint8_t *in, size_t n, int compressed, void *oxy, uint8_t *oinf, uint8_t *status) {
  constexpr size_t FB = field_traits<F>::bytes;
  size_t ib = (compressed ? FB : 2 * FB) * n;
  int rc = stage_reserve(ctx, ib + 2 * FB * n + 2 * n + 8 * 256);
  if (rc != B200_OK) return rc;
  void *din = stage_take(ctx, ib), *dxy = stage_take(ctx, 2 * FB * n), *di = stage_take(ctx, n), *ds = stage_take(ctx, n);
  B200_CUDA(ctx, cudaMemcpyAsync(din, in, ib, cudaMemcpyHostToDevice, ctx->stream));
  B200_LAUNCH(ctx, k_deserialize<F>, nblk(n, 128), 128, 0, (const uint8_t *)din, n, compressed, (char *)dxy, (uint8_t *)di,
              (uint8_t *)ds);
  B200_CUDA(ctx, cudaMemcpyAsync(oxy, dxy, 2 * FB * n, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaMemcpyAsync(oinf, di, n, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaMemcpyAsync(status, ds, n, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

template <class F>
int checks_host(b200_ctx *ctx, const void *xy, const uint8_t *inf, size_t n, uint8_t *status) {
  constexpr size_t FB = field_traits<F>::bytes;
  int rc = stage_reserve(ctx, 2 * FB * n + 2 * n + 8 * 256);
  if (rc != B200_OK) return rc;
  void *dxy = stage_take(ctx, 2 * FB * n), *di = inf ? stage_take(ctx, n) : nullptr, *ds = stage_take(ctx, n);
  B200_CUDA(ctx, cudaMemcpyAsync(dxy, xy, 2 * FB * n, cudaMemcpyHostToDevice, ctx->stream));
  if (inf) B200_CUDA(ctx, cudaMemcpyAsync(di, inf, n, cudaMemcpyHostToDevice, ctx->stream));
  B200_LAUNCH(ctx, k_point_checks<F>, nblk(n, 128), 128, 0, (const char *)dxy, (const uint8_t *)di, n, (uint8_t *)ds);
  B200_CUDA(ctx, cudaMemcpyAsync(status, ds, n, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

}  // namespace

#define CHECK_CTX(ctx)                      \
  if ((ctx) == nullptr) return B200_EINVAL; \
  ctx_guard guard__(ctx);                   \
  if (!guard__.ok) return B200_ENODEV

extern "C" {

int b200_g1_serialize(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *inf, size_t n, int compressed, uint8_t *out) {
  CHECK_CTX(ctx);
  if (n && (!p || !out)) return B200_EINVAL;
  return n ? serialize_host<fp>(ctx, p, inf, n, compressed != 0, out) : B200_OK;
}
int b200_g2_serialize(b200_ctx *ctx, const b200_g2_affine *p, const uint8_t *inf, size_t n, int compressed, uint8_t *out) {
  CHECK_CTX(ctx);
  if (n && (!p || !out)) return B200_EINVAL;
  return n ? serialize_host<fp2>(ctx, p, inf, n, compressed != 0, out) : B200_OK;
}
int b200_g1_deserialize(b200_ctx *ctx, const uint8_t *in, size_t n, int compressed, b200_g1_affine *out, uint8_t *out_inf,
                        uint8_t *status) {
  CHECK_CTX(ctx);
  if (n && (!in || !out || !out_inf || !status)) return B200_EINVAL;
  return n ? deserialize_host<fp>(ctx, in, n, compressed != 0, out, out_inf, status) : B200_OK;
}
int b200_g2_deserialize(b200_ctx *ctx, const uint8_t *in, size_t n, int compressed, b200_g2_affine *out, uint8_t *out_inf,
                        uint8_t *status) {
  CHECK_CTX(ctx);
  if (n && (!in || !out || !out_inf || !status)) return B200_EINVAL;
  return n ? deserialize_host<fp2>(ctx, in, n, compressed != 0, out, out_inf, status) : B200_OK;
}

int b200_g1_check(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *inf, size_t n, uint8_t *status) {
  CHECK_CTX(ctx);
  if (n && (!p || !status)) return B200_EINVAL;
  return n ? checks_host<fp>(ctx, p, inf, n, status) : B200_OK;
}
int b200_g2_check(b200_ctx *ctx, const b200_g2_affine *p, const uint8_t *inf, size_t n, uint8_t *status) {
  CHECK_CTX(ctx);
  if (n && (!p || !status)) return B200_EINVAL;
  return n ? checks_host<fp2>(ctx, p, inf, n, status) : B200_OK;
}

}  // extern "C"
