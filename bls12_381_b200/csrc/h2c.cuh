// Batched hash-to-curve for G1 and G2 — the suites BLS12381G{1,2}_XMD:SHA-256_SSWU_{RO,NU}_ of
// RFC 9380, i.e. the reference's `HashToCurve<ExpandMsgXmd<Sha256>>::{hash_to_curve, encode_to_curve}`
// (src/hash_to_curve/mod.rs:81-109) for `G1Projective` / `G2Projective`.
//
// Replaces: src/hash_to_curve/expand_msg.rs (ExpandMsgXmd :230-300, oversize DST :64-84), map_g1.rs (from_okm :513, sgn0
// :535, map_to_curve_simple_swu :550, iso_map :589), map_g2.rs (:374, :382, :391, :457), src/g1.rs:800 clear_cofactor,
// src/g2.rs:890 psi2 / :938 clear_cofactor.  Same formulas in the same order, so the projective (x, y, z) limbs are
// identical to the reference's; the two fixed exponentiations use square-and-multiply instead of the addition chains of
// chain.rs (same field element).  SHA-256 (a third-party crate in the reference) is FIPS 180-4.
//
// Shape on the GPU: one thread per message, two kernels — k_h2c_expand (byte work: SHA-256 compressions, message read
// once from HBM) writes the uniform bytes, k_h2c_map_g{1,2} (integer-pipe work: ~1.2k / ~7k FpM per point, dominated by
// the square-root exponentiation and the cofactor clearing) reads them and writes projective points.  Like the other
// kernels of this library it is variable-time; hashing secret inputs stays on the CPU path.
#pragma once
#include "constants.cuh"
#include "curve.cuh"
#include "fr.cuh"
#include "h2c_constants.cuh"

namespace b200 {

#if defined(__CUDACC__) && !defined(B200_HOST_EMUL)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

// ------------------------------------------------------------------ SHA-256 (FIPS 180-4), usable on host and device
struct sha256_state {
  uint32_t h[8];
  uint32_t w[16];      // current block, big-endian words
  uint64_t len;        // bytes absorbed
};
B200_HD uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
B200_HD uint32_t sha_k(int i) {
  const uint32_t K[64] = {
      0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u,
      0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u,
      0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u,
      0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u,
      0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, 0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u,
      0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au,
      0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u,
      0xc67178f2u};
  return K[i];
}
B200_HD void sha256_init(sha256_state &s) {
  s.h[0] = 0x6a09e667u; s.h[1] = 0xbb67ae85u; s.h[2] = 0x3c6ef372u; s.h[3] = 0xa54ff53au;
  s.h[4] = 0x510e527fu; s.h[5] = 0x9b05688cu; s.h[6] = 0x1f83d9abu; s.h[7] = 0x5be0cd19u;
  for (int i = 0; i < 16; i++) s.w[i] = 0;
  s.len = 0;
}
B200_HD void sha256_compress(sha256_state &s) {
  uint32_t w[16];
  for (int i = 0; i < 16; i++) w[i] = s.w[i];
  uint32_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3], e = s.h[4], f = s.h[5], g = s.h[6], hh = s.h[7];
  for (int i = 0; i < 64; i++) {
    uint32_t wi;
    if (i < 16) {
      wi = w[i];
    } else {  // rolling 16-word message schedule
      uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
      uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
      wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
      w[i & 15] = wi;
    }
    uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25), ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + sha_k(i) + wi;
    uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22), maj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + maj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d; s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += hh;
}
B200_HD void sha256_byte(sha256_state &s, uint8_t v) {
  const int pos = (int)(s.len & 63);
  const int sh = 24 - 8 * (pos & 3);
  s.w[pos >> 2] = (s.w[pos >> 2] & ~(0xffu << sh)) | ((uint32_t)v << sh);
  s.len++;
  if ((s.len & 63) == 0) sha256_compress(s);
}
B200_HD void sha256_update(sha256_state &s, const uint8_t *p, size_t n) {
  for (size_t i = 0; i < n; i++) sha256_byte(s, p[i]);
}
B200_HD void sha256_final(sha256_state &s, uint8_t out[32]) {
  const uint64_t bits = s.len * 8;
  sha256_byte(s, 0x80);
  while ((s.len & 63) != 56) sha256_byte(s, 0);
  for (int i = 0; i < 8; i++) sha256_byte(s, (uint8_t)(bits >> (56 - 8 * i)));
  for (int i = 0; i < 8; i++)
    for (int k = 0; k < 4; k++) out[4 * i + k] = (uint8_t)(s.h[i] >> (24 - 8 * k));
}

// DST_prime = DST || I2OSP(len(DST), 1) with the RFC 9380 §5.3.3 reduction of an over-long DST
// (src/hash_to_curve/expand_msg.rs:64-84).  out holds up to 256 bytes; returns len(DST_prime).  Runs on the HOST, once
// per call (the tag is shared by the whole batch).
B200_HD int h2c_dst_prime(const uint8_t *dst, size_t dst_len, uint8_t *out) {
  int dl;
  if (dst_len > 255) {
    sha256_state hs;
    sha256_init(hs);
    const char salt[] = "H2C-OVERSIZE-DST-";
    sha256_update(hs, reinterpret_cast<const uint8_t *>(salt), 17);
    sha256_update(hs, dst, dst_len);
    sha256_final(hs, out);
    dl = 32;
  } else {
    for (size_t i = 0; i < dst_len; i++) out[i] = dst[i];
    dl = (int)dst_len;
  }
  out[dl] = (uint8_t)dl;
  return dl + 1;
}

// expand_message_xmd (src/hash_to_curve/expand_msg.rs:230-300 / RFC 9380 §5.3.1) for one message.
// Caller guarantees ell = ceil(len_in_bytes / 32) <= 255 and len_in_bytes <= 65535.
B200_HD void h2c_expand_message_xmd(const uint8_t *msg, size_t msg_len, const uint8_t *dst_prime, int dp_len,
                                    uint32_t len_in_bytes, uint8_t *out) {
  uint8_t b0[32], bi[32];
  sha256_state hs;
  sha256_init(hs);
  for (int i = 0; i < 64; i++) sha256_byte(hs, 0);  // Z_pad: one block of zeros
  sha256_update(hs, msg, msg_len);
  sha256_byte(hs, (uint8_t)(len_in_bytes >> 8));
  sha256_byte(hs, (uint8_t)len_in_bytes);
  sha256_byte(hs, 0);
  sha256_update(hs, dst_prime, dp_len);
  sha256_final(hs, b0);
  sha256_init(hs);
  sha256_update(hs, b0, 32);
  sha256_byte(hs, 1);
  sha256_update(hs, dst_prime, dp_len);
  sha256_final(hs, bi);
  const uint32_t ell = (len_in_bytes + 31) / 32;
  uint32_t off = 0;
  for (uint32_t i = 1; i <= ell; i++) {
    if (i > 1) {
      sha256_init(hs);
      for (int k = 0; k < 32; k++) sha256_byte(hs, b0[k] ^ bi[k]);
      sha256_byte(hs, (uint8_t)i);
      sha256_update(hs, dst_prime, dp_len);
      sha256_final(hs, bi);
    }
    for (int k = 0; k < 32 && off < len_in_bytes; k++) out[off++] = bi[k];
  }
}

// one thread per message: msgs[off[i] .. off[i+1]) -> okm[i * len_in_bytes ..)
static __global__ void __launch_bounds__(128) k_h2c_expand(const uint8_t *msgs, const uint64_t *off, size_t n,
                                                          const uint8_t *dst_prime, int dp_len, uint32_t len_in_bytes,
                                                          uint8_t *okm) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  h2c_expand_message_xmd(msgs + off[i], (size_t)(off[i + 1] - off[i]), dst_prime, dp_len, len_in_bytes,
                         okm + (size_t)len_in_bytes * i);
}

// ------------------------------------------------------------------ hash_to_field helpers
B200_DEV fp h2c_const(const uint32_t (*tab)[12], int i) { return fp_const(tab[i]); }
B200_DEV fp2 h2c_const2(const uint32_t (*tab)[12], int i) { return fp2{fp_const(tab[2 * i]), fp_const(tab[2 * i + 1])}; }

// 32 big-endian bytes -> the integer as a Montgomery-form element (value < 2^256 < p: always canonical)
B200_DEV fp h2c_fp_from_be32(const uint8_t *b) {
  fp raw = fp_zero();
#pragma unroll
  for (int w = 0; w < 8; w++) {
    const uint8_t *q = b + 28 - 4 * w;
    raw.v[w] = (uint32_t)q[0] << 24 | (uint32_t)q[1] << 16 | (uint32_t)q[2] << 8 | q[3];
  }
  return fp_mul_c(raw, fp_const(K_R2));  // x * R^2 * R^-1 = x R
}
// src/hash_to_curve/map_g1.rs:513-532
B200_DEV fp h2c_fp_from_okm(const uint8_t *okm) {
  fp db = h2c_fp_from_be32(okm), da = h2c_fp_from_be32(okm + 32);
  return fp_add(fp_mul_c(db, fp_const(H2C_F_2_256)), da);
}
// src/hash_to_curve/map_g1.rs:535-545: parity of the canonical integer
B200_DEV bool h2c_sgn0(const fp &a) {
  fp one_raw = fp_zero();
  one_raw.v[0] = 1;
  return (fp_mul_c(a, one_raw).v[0] & 1u) != 0;
}
// src/hash_to_curve/map_g2.rs:382-388
B200_DEV bool h2c_sgn0(const fp2 &a) { return h2c_sgn0(a.c0) || (fp_is_zero(a.c0) && h2c_sgn0(a.c1)); }

// a^e, e = nwords plain little-endian 32-bit words, MSB first
B200_DEV fp h2c_pow(const fp &a, const uint32_t *e, int nwords) {
  fp res = fp_one();
#pragma unroll 1
  for (int w = nwords - 1; w >= 0; w--) {
    const uint32_t ew = e[w];
#pragma unroll 1
    for (int i = 31; i >= 0; i--) {
      res = fp_sqr_c(res);
      if ((ew >> i) & 1) res = fp_mul_c(res, a);
    }
  }
  return res;
}
B200_DEV fp2 h2c_pow(const fp2 &a, const uint32_t *e, int nwords) {
  fp2 res = fp2_one();
#pragma unroll 1
  for (int w = nwords - 1; w >= 0; w--) {
    const uint32_t ew = e[w];
#pragma unroll 1
    for (int i = 31; i >= 0; i--) {
      res = S2(res);
      if ((ew >> i) & 1) res = M2(res, a);
    }
  }
  return res;
}

// group operations on E as single out-of-line copies (the complete formulas of curve.cuh are force-inlined)
static __device__ __noinline__ void h2c_add(proj<fp> *r, const proj<fp> *a, const proj<fp> *b) { *r = proj_add(*a, *b); }
static __device__ __noinline__ void h2c_dbl(proj<fp> *r, const proj<fp> *a) { *r = proj_double(*a); }
static __device__ __noinline__ void h2c_add(proj<fp2> *r, const proj<fp2> *a, const proj<fp2> *b) { *r = proj_add(*a, *b); }
static __device__ __noinline__ void h2c_dbl(proj<fp2> *r, const proj<fp2> *a) { *r = proj_double(*a); }
// [x]P, x = -0xd201000000010000 (src/g1.rs:777-795, src/g2.rs:915-932)
template <class F>
static __device__ __noinline__ proj<F> h2c_mul_by_x(const proj<F> &s) {
  proj<F> xself = proj_identity<F>(), acc = s;
  unsigned long long x = 0xd201000000010000ull >> 1;
#pragma unroll 1
  while (x != 0) {
    h2c_dbl(&acc, &acc);
    if (x & 1) h2c_add(&xself, &xself, &acc);
    x >>= 1;
  }
  return proj_neg(xself);
}

// ------------------------------------------------------------------ G1
// src/hash_to_curve/map_g1.rs:550-586
static __device__ __noinline__ proj<fp> h2c_g1_sswu(const fp &u) {
  const fp A = h2c_const(H2C_G1_SSWU_ELLP_A, 0), B = h2c_const(H2C_G1_SSWU_ELLP_B, 0), XI = h2c_const(H2C_G1_SSWU_XI, 0);
  fp usq = fp_sqr_c(u), xi_usq = fp_mul_c(XI, usq), xisq_u4 = fp_sqr_c(xi_usq);
  fp nd_common = fp_add(xisq_u4, xi_usq);
  fp x_den = fp_mul_c(A, fp_is_zero(nd_common) ? XI : fp_neg(nd_common));
  fp x0_num = fp_mul_c(B, fp_add(fp_one(), nd_common));
  fp x_densq = fp_sqr_c(x_den), gx_den = fp_mul_c(x_densq, x_den);
  fp gx0_num = fp_add(fp_mul_c(fp_add(fp_sqr_c(x0_num), fp_mul_c(A, x_densq)), x0_num), fp_mul_c(B, gx_den));
  fp u_v = fp_mul_c(gx0_num, gx_den), vsq = fp_sqr_c(gx_den);
  fp sqrt_candidate = fp_mul_c(u_v, h2c_pow(fp_mul_c(u_v, vsq), H2C_EXP_PM3DIV4, 12));
  bool gx0_square = fp_eq(fp_mul_c(fp_sqr_c(sqrt_candidate), gx_den), gx0_num);
  fp x1_num = fp_mul_c(x0_num, xi_usq);
  fp y1 = fp_mul_c(fp_mul_c(fp_mul_c(h2c_const(H2C_G1_SQRT_M_XI_CUBED, 0), usq), u), sqrt_candidate);
  fp x_num = gx0_square ? x0_num : x1_num;
  fp y = gx0_square ? sqrt_candidate : y1;
  if (h2c_sgn0(y) != h2c_sgn0(u)) y = fp_neg(y);
  return proj<fp>{x_num, fp_mul_c(y, x_den), x_den};
}
// Horner evaluation of one of the four isogeny polynomials (src/hash_to_curve/map_g1.rs:609-617)
template <int LEN>
B200_DEV fp h2c_g1_iso_poly(const uint32_t (*coeff)[12], const fp &x, const fp *zpows) {
  fp v = h2c_const(coeff, LEN - 1);
#pragma unroll 1
  for (int j = 0; j < LEN - 1; j++) v = fp_add(fp_mul_c(v, x), fp_mul_c(zpows[j], h2c_const(coeff, LEN - 2 - j)));
  return v;
}
// src/hash_to_curve/map_g1.rs:589-631
static __device__ __noinline__ proj<fp> h2c_g1_iso_map(const proj<fp> &u) {
  fp zpows[15];
  zpows[0] = u.z;
#pragma unroll 1
  for (int i = 1; i < 15; i++) zpows[i] = fp_mul_c(zpows[i - 1], u.z);
  fp xnum = h2c_g1_iso_poly<12>(H2C_G1_ISO11_XNUM, u.x, zpows);
  fp xden = h2c_g1_iso_poly<11>(H2C_G1_ISO11_XDEN, u.x, zpows);
  fp ynum = h2c_g1_iso_poly<16>(H2C_G1_ISO11_YNUM, u.x, zpows);
  fp yden = h2c_g1_iso_poly<16>(H2C_G1_ISO11_YDEN, u.x, zpows);
  xden = fp_mul_c(xden, u.z);
  ynum = fp_mul_c(ynum, u.y);
  yden = fp_mul_c(yden, u.z);
  return proj<fp>{fp_mul_c(xnum, yden), fp_mul_c(ynum, xden), fp_mul_c(xden, yden)};
}
B200_DEV proj<fp> h2c_g1_map_to_curve(const fp &u) { return h2c_g1_iso_map(h2c_g1_sswu(u)); }
// src/g1.rs:800-802: self - [x]self
static __device__ __noinline__ proj<fp> h2c_g1_clear_cofactor(const proj<fp> &p) {
  proj<fp> m = proj_neg(h2c_mul_by_x(p)), r;
  h2c_add(&r, &p, &m);
  return r;
}

// ------------------------------------------------------------------ G2
// src/hash_to_curve/map_g2.rs:391-454
static __device__ __noinline__ proj<fp2> h2c_g2_sswu(const fp2 &u) {
  const fp2 A = h2c_const2(H2C_G2_SSWU_ELLP_A, 0), B = h2c_const2(H2C_G2_SSWU_ELLP_B, 0), XI = h2c_const2(H2C_G2_SSWU_XI, 0);
  fp2 usq = S2(u), xi_usq = M2(XI, usq), xisq_u4 = S2(xi_usq);
  fp2 nd_common = fp2_add(xisq_u4, xi_usq);
  fp2 x_den = M2(A, fp2_is_zero(nd_common) ? XI : fp2_neg(nd_common));
  fp2 x0_num = M2(B, fp2_add(fp2_one(), nd_common));
  fp2 x_densq = S2(x_den), gx_den = M2(x_densq, x_den);
  fp2 gx0_num = fp2_add(M2(fp2_add(S2(x0_num), M2(A, x_densq)), x0_num), M2(B, gx_den));
  fp2 sqrt_candidate;
  {
    fp2 vsq = S2(gx_den), v_3 = M2(vsq, gx_den), v_4 = S2(vsq);
    fp2 uv_7 = M2(M2(gx0_num, v_3), v_4), uv_15 = M2(uv_7, S2(v_4));
    sqrt_candidate = M2(uv_7, h2c_pow(uv_15, H2C_EXP_P2M9DIV16, 24));
  }
  fp2 y = sqrt_candidate;
  fp2 tmp = fp2{fp_neg(sqrt_candidate.c1), sqrt_candidate.c0};
  if (fp2_eq(M2(S2(tmp), gx_den), gx0_num)) y = tmp;
  tmp = M2(sqrt_candidate, h2c_const2(H2C_G2_SSWU_RV1, 0));
  if (fp2_eq(M2(S2(tmp), gx_den), gx0_num)) y = tmp;
  tmp = fp2{tmp.c1, fp_neg(tmp.c0)};
  if (fp2_eq(M2(S2(tmp), gx_den), gx0_num)) y = tmp;
  fp2 gx1_num = M2(M2(gx0_num, xi_usq), xisq_u4);
  fp2 sc = M2(M2(sqrt_candidate, usq), u);
  bool eta_found = false;
#pragma unroll 1
  for (int k = 0; k < 4; k++) {
    fp2 t = M2(sc, h2c_const2(H2C_G2_SSWU_ETAS, k));
    bool found = fp2_eq(M2(S2(t), gx_den), gx1_num);
    if (found) y = t;
    eta_found = eta_found || found;
  }
  fp2 x_num = eta_found ? M2(x0_num, xi_usq) : x0_num;
  if (h2c_sgn0(u) != h2c_sgn0(y)) y = fp2_neg(y);
  return proj<fp2>{x_num, M2(y, x_den), x_den};
}
template <int LEN>
B200_DEV fp2 h2c_g2_iso_poly(const uint32_t (*coeff)[12], const fp2 &x, const fp2 *zpows) {
  fp2 v = h2c_const2(coeff, LEN - 1);
#pragma unroll 1
  for (int j = 0; j < LEN - 1; j++) v = fp2_add(M2(v, x), M2(zpows[j], h2c_const2(coeff, LEN - 2 - j)));
  return v;
}
// src/hash_to_curve/map_g2.rs:457-493
static __device__ __noinline__ proj<fp2> h2c_g2_iso_map(const proj<fp2> &u) {
  fp2 zpows[3];
  zpows[0] = u.z;
  zpows[1] = S2(u.z);
  zpows[2] = M2(zpows[1], u.z);
  fp2 xnum = h2c_g2_iso_poly<4>(H2C_G2_ISO3_XNUM, u.x, zpows);
  fp2 xden = h2c_g2_iso_poly<3>(H2C_G2_ISO3_XDEN, u.x, zpows);
  fp2 ynum = h2c_g2_iso_poly<4>(H2C_G2_ISO3_YNUM, u.x, zpows);
  fp2 yden = h2c_g2_iso_poly<4>(H2C_G2_ISO3_YDEN, u.x, zpows);
  xden = M2(xden, u.z);
  ynum = M2(ynum, u.y);
  yden = M2(yden, u.z);
  return proj<fp2>{M2(xnum, yden), M2(ynum, xden), M2(xden, yden)};
}
B200_DEV proj<fp2> h2c_g2_map_to_curve(const fp2 &u) { return h2c_g2_iso_map(h2c_g2_sswu(u)); }
// src/g2.rs:847-888 psi, :890-912 psi2
B200_DEV proj<fp2> h2c_psi(const proj<fp2> &s) {
  fp2 cx = fp2{fp_zero(), fp_const(K_PSI_X_U)}, cy = fp2{fp_const(K_PSI_Y_R), fp_const(K_PSI_Y_U)};
  return proj<fp2>{M2(fp2_conj(s.x), cx), M2(fp2_conj(s.y), cy), fp2_conj(s.z)};
}
B200_DEV proj<fp2> h2c_psi2(const proj<fp2> &s) {
  fp2 cx = fp2{fp_const(K_FROB6_C1_U), fp_zero()};  // 1 / 2^((p-1)/3): the limbs of src/g2.rs:893-900
  return proj<fp2>{M2(s.x, cx), fp2_neg(s.y), s.z};
}
// src/g2.rs:938-947, the same operator order
static __device__ __noinline__ proj<fp2> h2c_g2_clear_cofactor(const proj<fp2> &p) {
  proj<fp2> t1 = h2c_mul_by_x(p), t2 = h2c_psi(p), d, s, r;
  h2c_dbl(&d, &p);
  d = h2c_psi2(d);
  h2c_add(&s, &t1, &t2);
  s = h2c_mul_by_x(s);
  h2c_add(&r, &d, &s);
  t1 = proj_neg(t1);
  h2c_add(&r, &r, &t1);
  t2 = proj_neg(t2);
  h2c_add(&r, &r, &t2);
  d = proj_neg(p);
  h2c_add(&r, &r, &d);
  return r;
}

// ------------------------------------------------------------------ kernels: uniform bytes -> points, one thread each
// count = 2: hash_to_curve (src/hash_to_curve/mod.rs:86-92), count = 1: encode_to_curve (:103-108)
static __global__ void __launch_bounds__(128) k_h2c_map_g1(const uint8_t *okm, size_t n, int count, char *out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t *o = okm + (size_t)64 * count * i;
  proj<fp> p = h2c_g1_map_to_curve(h2c_fp_from_okm(o));
  if (count == 2) {
    proj<fp> q = h2c_g1_map_to_curve(h2c_fp_from_okm(o + 64));
    h2c_add(&p, &p, &q);
  }
  proj_store<fp>(out + 144 * i, h2c_g1_clear_cofactor(p));
}
static __global__ void __launch_bounds__(128) k_h2c_map_g2(const uint8_t *okm, size_t n, int count, char *out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t *o = okm + (size_t)128 * count * i;
  proj<fp2> p = h2c_g2_map_to_curve(fp2{h2c_fp_from_okm(o), h2c_fp_from_okm(o + 64)});  // map_g2.rs:374-378
  if (count == 2) {
    proj<fp2> q = h2c_g2_map_to_curve(fp2{h2c_fp_from_okm(o + 128), h2c_fp_from_okm(o + 192)});
    h2c_add(&p, &p, &q);
  }
  proj_store<fp2>(out + 288 * i, h2c_g2_clear_cofactor(p));
}
// the stages on their own (parity surface): kind 0 sswu (field element -> E' point), 1 iso_map, 2 map_to_curve,
// 3 clear_cofactor
template <class F>
__global__ void __launch_bounds__(128) k_h2c_stage(int kind, const char *in, size_t n, char *out) {
  constexpr size_t FB = field_traits<F>::bytes;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  proj<F> r;
  if (kind == 0 || kind == 2) {
    F u = field_traits<F>::load(in + FB * i);
    if constexpr (FB == 48) r = kind == 0 ? h2c_g1_sswu(u) : h2c_g1_map_to_curve(u);
    else r = kind == 0 ? h2c_g2_sswu(u) : h2c_g2_map_to_curve(u);
  } else {
    proj<F> p = proj_load<F>(in + 3 * FB * i);
    if constexpr (FB == 48) r = kind == 1 ? h2c_g1_iso_map(p) : h2c_g1_clear_cofactor(p);
    else r = kind == 1 ? h2c_g2_iso_map(p) : h2c_g2_clear_cofactor(p);
  }
  proj_store<F>(out + 3 * FB * i, r);
}

// ------------------------------------------------------------------ hash to field for Scalar
// src/hash_to_curve/map_scalar.rs:17-22: 48 big-endian uniform bytes, zero-extended to 512 bits, from_bytes_wide
B200_DEV fr h2c_fr_from_okm(const uint8_t *okm) {
  fr lo, hi = fr_zero();
#pragma unroll
  for (int w = 0; w < 8; w++) {  // little-endian word w of the 384-bit integer = bytes okm[44 - 4w .. 48 - 4w)
    const uint8_t *q = okm + 44 - 4 * w;
    lo.v[w] = (uint32_t)q[0] << 24 | (uint32_t)q[1] << 16 | (uint32_t)q[2] << 8 | q[3];
  }
#pragma unroll
  for (int w = 0; w < 4; w++) {
    const uint8_t *q = okm + 12 - 4 * w;
    hi.v[w] = (uint32_t)q[0] << 24 | (uint32_t)q[1] << 16 | (uint32_t)q[2] << 8 | q[3];
  }
  return fr_from_wide(lo, hi);
}
static __global__ void __launch_bounds__(128) k_h2c_fr_from_okm(const uint8_t *okm, size_t n, char *out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  fr_store(out + 32 * i, h2c_fr_from_okm(okm + 48 * i));
}

// The launch sequence of one batch (shared by capi_h2c.cu and the CPU test harness; `launch` as in fr_ntt.cuh).
// group 1 / 2; count 2 = hash_to_curve, 1 = encode_to_curve; okm = n * 64 * group * count bytes of scratch;
// dst_prime / dp_len from h2c_dst_prime (device-readable copy).  Returns the number of launches or a negative error.
inline uint32_t h2c_okm_bytes(int group, int count) { return 64u * (uint32_t)group * (uint32_t)count; }
template <class L>
int h2c_hash_run(L &&launch, int group, const uint8_t *msgs, const uint64_t *off, size_t n, const uint8_t *dst_prime,
                 int dp_len, int count, uint8_t *okm, char *out) {
  const unsigned grid = (unsigned)((n + 127) / 128);
  int rc = launch(k_h2c_expand, grid, 128u, msgs, off, n, dst_prime, dp_len, h2c_okm_bytes(group, count), okm);
  if (rc) return rc;
  if (group == 1)
    rc = launch(k_h2c_map_g1, grid, 128u, (const uint8_t *)okm, n, count, out);
  else
    rc = launch(k_h2c_map_g2, grid, 128u, (const uint8_t *)okm, n, count, out);
  return rc ? rc : 2;
}

}  // namespace b200
