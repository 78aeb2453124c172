// ORACLE — TEST INFRASTRUCTURE ONLY (see bls12_381_oracle.hpp).
//
// CPU restatement of the reference's hash-to-curve path (feature "experimental"):
//   src/hash_to_curve/expand_msg.rs  ExpandMsgXmd<Sha256>            (RFC 9380 §5.3.1, DST > 255 bytes §5.3.3)
//   src/hash_to_curve/mod.rs:35-67   HashToField::hash_to_field, :81-109 hash_to_curve / encode_to_curve
//   src/hash_to_curve/map_g1.rs      from_okm :513, sgn0 :535, map_to_curve_simple_swu :550, iso_map :589, clear_h :640
//   src/hash_to_curve/map_g2.rs      from_okm :374, sgn0 :382, map_to_curve_simple_swu :391, iso_map :457, clear_h :502
//   src/g1.rs:800 clear_cofactor, src/g2.rs:890 psi2, :938 clear_cofactor
// SHA-256 is a third-party dependency of the reference (crate `sha2`, not vendored in /root/reference): FIPS 180-4 is
// restated below and pinned by the RFC's expand_message_xmd vectors the reference tests hold (tests/expand_msg.rs) and
// by Python's hashlib in tests/test_oracle_h2c.py.  The addition chains of src/hash_to_curve/chain.rs compute
// x^((p-3)/4) and x^((p^2-9)/16); plain square-and-multiply over the same exponents gives the same field element.
#pragma once
#include "bls12_381_oracle.hpp"

namespace bls_oracle {

#include "h2c_constants.inc"

// ------------------------------------------------------------------ SHA-256 (FIPS 180-4)
struct Sha256 {
  uint32_t h[8];
  uint8_t buf[64];
  uint64_t len = 0;  // bytes absorbed
  Sha256() {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    std::memcpy(h, iv, 32);
  }
  static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  void block(const uint8_t *p) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
        0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
        0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
        0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
        0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
        0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
      uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
      uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g);
      uint32_t t1 = hh + S1 + ch + K[i] + w[i];
      uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), maj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + maj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  void update(const uint8_t *p, size_t n) {
    for (size_t i = 0; i < n; i++) {
      buf[len % 64] = p[i];
      len++;
      if (len % 64 == 0) block(buf);
    }
  }
  void finish(uint8_t out[32]) {
    uint64_t bits = len * 8;
    uint8_t pad = 0x80;
    update(&pad, 1);
    pad = 0;
    while (len % 64 != 56) update(&pad, 1);
    uint8_t lb[8];
    for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
    update(lb, 8);
    for (int i = 0; i < 8; i++)
      for (int k = 0; k < 4; k++) out[4 * i + k] = (uint8_t)(h[i] >> (24 - 8 * k));
  }
};

// src/hash_to_curve/expand_msg.rs:64-84 (ExpandMsgDst::for_xmd) + :230-300 (ExpandMsgXmd<Sha256>)
// returns false for the argument errors the reference panics on (ell > 255, len_in_bytes > 65535)
static inline bool expand_message_xmd_sha256(const uint8_t *msg, size_t msg_len, const uint8_t *dst_in, size_t dst_len,
                                             size_t len_in_bytes, uint8_t *out) {
  const size_t ell = (len_in_bytes + 31) / 32;
  if (ell > 255 || len_in_bytes > 65535) return false;
  uint8_t dst[256];
  size_t dl = dst_len;
  if (dst_len > 255) {
    Sha256 hd;
    hd.update((const uint8_t *)"H2C-OVERSIZE-DST-", 17);
    hd.update(dst_in, dst_len);
    hd.finish(dst);
    dl = 32;
  } else {
    std::memcpy(dst, dst_in, dst_len);
  }
  dst[dl] = (uint8_t)dl;  // DST_prime = DST || I2OSP(len(DST), 1)
  uint8_t b0[32], bi[32];
  {
    Sha256 h0;
    uint8_t z[64] = {0};
    h0.update(z, 64);
    h0.update(msg, msg_len);
    uint8_t l[3] = {(uint8_t)(len_in_bytes >> 8), (uint8_t)len_in_bytes, 0};
    h0.update(l, 3);
    h0.update(dst, dl + 1);
    h0.finish(b0);
  }
  {
    Sha256 h1;
    h1.update(b0, 32);
    uint8_t one = 1;
    h1.update(&one, 1);
    h1.update(dst, dl + 1);
    h1.finish(bi);
  }
  size_t off = 0;
  for (size_t i = 1; i <= ell; i++) {
    if (i > 1) {
      uint8_t x[32];
      for (int k = 0; k < 32; k++) x[k] = b0[k] ^ bi[k];
      Sha256 hi;
      hi.update(x, 32);
      uint8_t ib = (uint8_t)i;
      hi.update(&ib, 1);
      hi.update(dst, dl + 1);
      hi.finish(bi);
    }
    size_t take = len_in_bytes - off < 32 ? len_in_bytes - off : 32;
    std::memcpy(out + off, bi, take);
    off += take;
  }
  return true;
}

// ------------------------------------------------------------------ hash_to_field
// src/hash_to_curve/map_g1.rs:513-532: 64 big-endian bytes -> Fp as  hi * 2^256 + lo
static inline Fp fp_from_okm(const uint8_t okm[64]) {
  uint8_t bs[48] = {0};
  Fp db, da;
  std::memcpy(bs + 16, okm, 32);
  (void)fp_from_bytes(bs, db);
  std::memcpy(bs + 16, okm + 32, 32);
  (void)fp_from_bytes(bs, da);
  return fp_add(fp_mul(db, H2C_F_2_256), da);
}
// src/hash_to_curve/map_g1.rs:535-545: parity of the canonical integer
static inline bool fp_sgn0(const Fp &a) {
  u64 t[12] = {a.l[0], a.l[1], a.l[2], a.l[3], a.l[4], a.l[5], 0, 0, 0, 0, 0, 0};
  return (fp_montgomery_reduce(t).l[0] & 1) != 0;
}
// src/hash_to_curve/map_g2.rs:382-388
static inline bool fp2_sgn0(const Fp2 &a) { return fp_sgn0(a.c0) || (fp_is_zero(a.c0) && fp_sgn0(a.c1)); }
static inline Fp2 mk2(const Fp *c) { return Fp2{c[0], c[1]}; }

// ------------------------------------------------------------------ G1: SSWU onto E', 11-isogeny, cofactor
// src/hash_to_curve/map_g1.rs:550-586
static inline G1Projective g1_map_to_curve_simple_swu(const Fp &u) {
  const Fp A = H2C_G1_SSWU_ELLP_A[0], B = H2C_G1_SSWU_ELLP_B[0], XI = H2C_G1_SSWU_XI[0];
  Fp usq = fp_square(u), xi_usq = fp_mul(XI, usq), xisq_u4 = fp_square(xi_usq);
  Fp nd_common = fp_add(xisq_u4, xi_usq);
  Fp x_den = fp_mul(A, fp_is_zero(nd_common) ? XI : fp_neg(nd_common));
  Fp x0_num = fp_mul(B, fp_add(fp_one(), nd_common));
  Fp x_densq = fp_square(x_den), gx_den = fp_mul(x_densq, x_den);
  Fp gx0_num = fp_add(fp_mul(fp_add(fp_square(x0_num), fp_mul(A, x_densq)), x0_num), fp_mul(B, gx_den));
  Fp u_v = fp_mul(gx0_num, gx_den), vsq = fp_square(gx_den);
  Fp sqrt_candidate = fp_mul(u_v, fp_pow_vartime(fp_mul(u_v, vsq), H2C_EXP_PM3DIV4));  // chain_pm3div4 (chain.rs)
  bool gx0_square = fp_eq(fp_mul(fp_square(sqrt_candidate), gx_den), gx0_num);
  Fp x1_num = fp_mul(x0_num, xi_usq);
  Fp y1 = fp_mul(fp_mul(fp_mul(H2C_G1_SQRT_M_XI_CUBED[0], usq), u), sqrt_candidate);
  Fp x_num = gx0_square ? x0_num : x1_num;
  Fp y = gx0_square ? sqrt_candidate : y1;
  if (fp_sgn0(y) != fp_sgn0(u)) y = fp_neg(y);
  return G1Projective{x_num, fp_mul(y, x_den), x_den};
}
// src/hash_to_curve/map_g1.rs:589-631
static inline G1Projective g1_iso_map(const G1Projective &u) {
  const Fp *coeffs[4] = {H2C_G1_ISO11_XNUM, H2C_G1_ISO11_XDEN, H2C_G1_ISO11_YNUM, H2C_G1_ISO11_YDEN};
  const int lens[4] = {12, 11, 16, 16};
  Fp zpows[15];
  zpows[0] = u.z;
  for (int i = 1; i < 15; i++) zpows[i] = fp_mul(zpows[i - 1], u.z);
  Fp mapvals[4];
  for (int idx = 0; idx < 4; idx++) {
    const int clast = lens[idx] - 1;
    mapvals[idx] = coeffs[idx][clast];
    for (int jdx = 0; jdx < clast; jdx++)
      mapvals[idx] = fp_add(fp_mul(mapvals[idx], u.x), fp_mul(zpows[jdx], coeffs[idx][clast - 1 - jdx]));
  }
  mapvals[1] = fp_mul(mapvals[1], u.z);
  mapvals[2] = fp_mul(mapvals[2], u.y);
  mapvals[3] = fp_mul(mapvals[3], u.z);
  return G1Projective{fp_mul(mapvals[0], mapvals[3]), fp_mul(mapvals[2], mapvals[1]), fp_mul(mapvals[1], mapvals[3])};
}
static inline G1Projective g1_map_to_curve(const Fp &u) { return g1_iso_map(g1_map_to_curve_simple_swu(u)); }  // :635
// src/g1.rs:800-802: self - [x]self  (the reference's Sub is add(neg))
static inline G1Projective g1p_clear_cofactor(const G1Projective &p) { return g1p_add(p, g1p_neg(g1p_mul_by_x(p))); }

// ------------------------------------------------------------------ G2
// src/hash_to_curve/map_g2.rs:391-454
static inline Fp2 fp2_pow_words(const Fp2 &a, const u64 *by, int nwords) {
  Fp2 res = fp2_one();
  for (int e = nwords - 1; e >= 0; e--)
    for (int i = 63; i >= 0; i--) {
      res = fp2_square(res);
      if ((by[e] >> i) & 1) res = fp2_mul(res, a);
    }
  return res;
}
static inline G2Projective g2_map_to_curve_simple_swu(const Fp2 &u) {
  const Fp2 A = mk2(H2C_G2_SSWU_ELLP_A), B = mk2(H2C_G2_SSWU_ELLP_B), XI = mk2(H2C_G2_SSWU_XI), RV1 = mk2(H2C_G2_SSWU_RV1);
  Fp2 usq = fp2_square(u), xi_usq = fp2_mul(XI, usq), xisq_u4 = fp2_square(xi_usq);
  Fp2 nd_common = fp2_add(xisq_u4, xi_usq);
  Fp2 x_den = fp2_mul(A, fp2_is_zero(nd_common) ? XI : fp2_neg(nd_common));
  Fp2 x0_num = fp2_mul(B, fp2_add(fp2_one(), nd_common));
  Fp2 x_densq = fp2_square(x_den), gx_den = fp2_mul(x_densq, x_den);
  Fp2 gx0_num = fp2_add(fp2_mul(fp2_add(fp2_square(x0_num), fp2_mul(A, x_densq)), x0_num), fp2_mul(B, gx_den));
  Fp2 sqrt_candidate;
  {
    Fp2 vsq = fp2_square(gx_den), v_3 = fp2_mul(vsq, gx_den), v_4 = fp2_square(vsq);
    Fp2 uv_7 = fp2_mul(fp2_mul(gx0_num, v_3), v_4), uv_15 = fp2_mul(uv_7, fp2_square(v_4));
    sqrt_candidate = fp2_mul(uv_7, fp2_pow_words(uv_15, H2C_EXP_P2M9DIV16, 12));  // chain_p2m9div16 (chain.rs)
  }
  Fp2 y = sqrt_candidate;
  Fp2 tmp = Fp2{fp_neg(sqrt_candidate.c1), sqrt_candidate.c0};
  if (fp2_eq(fp2_mul(fp2_square(tmp), gx_den), gx0_num)) y = tmp;
  tmp = fp2_mul(sqrt_candidate, RV1);
  if (fp2_eq(fp2_mul(fp2_square(tmp), gx_den), gx0_num)) y = tmp;
  tmp = Fp2{tmp.c1, fp_neg(tmp.c0)};
  if (fp2_eq(fp2_mul(fp2_square(tmp), gx_den), gx0_num)) y = tmp;
  Fp2 gx1_num = fp2_mul(fp2_mul(gx0_num, xi_usq), xisq_u4);
  Fp2 sc = fp2_mul(fp2_mul(sqrt_candidate, usq), u);
  bool eta_found = false;
  for (int k = 0; k < 4; k++) {
    Fp2 t = fp2_mul(sc, mk2(H2C_G2_SSWU_ETAS + 2 * k));
    bool found = fp2_eq(fp2_mul(fp2_square(t), gx_den), gx1_num);
    if (found) y = t;
    eta_found = eta_found || found;
  }
  Fp2 x_num = eta_found ? fp2_mul(x0_num, xi_usq) : x0_num;
  if (fp2_sgn0(u) != fp2_sgn0(y)) y = fp2_neg(y);
  return G2Projective{x_num, fp2_mul(y, x_den), x_den};
}
// src/hash_to_curve/map_g2.rs:457-493
static inline G2Projective g2_iso_map(const G2Projective &u) {
  const Fp *coeffs[4] = {H2C_G2_ISO3_XNUM, H2C_G2_ISO3_XDEN, H2C_G2_ISO3_YNUM, H2C_G2_ISO3_YDEN};
  const int lens[4] = {4, 3, 4, 4};
  Fp2 zsq = fp2_square(u.z);
  Fp2 zpows[3] = {u.z, zsq, fp2_mul(zsq, u.z)};
  Fp2 mapvals[4];
  for (int idx = 0; idx < 4; idx++) {
    const int clast = lens[idx] - 1;
    mapvals[idx] = mk2(coeffs[idx] + 2 * clast);
    for (int jdx = 0; jdx < clast; jdx++)
      mapvals[idx] = fp2_add(fp2_mul(mapvals[idx], u.x), fp2_mul(zpows[jdx], mk2(coeffs[idx] + 2 * (clast - 1 - jdx))));
  }
  mapvals[1] = fp2_mul(mapvals[1], u.z);
  mapvals[2] = fp2_mul(mapvals[2], u.y);
  mapvals[3] = fp2_mul(mapvals[3], u.z);
  return G2Projective{fp2_mul(mapvals[0], mapvals[3]), fp2_mul(mapvals[2], mapvals[1]), fp2_mul(mapvals[1], mapvals[3])};
}
static inline G2Projective g2_map_to_curve(const Fp2 &u) { return g2_iso_map(g2_map_to_curve_simple_swu(u)); }  // :497
// src/g2.rs:890-912
static inline G2Projective g2p_psi2(const G2Projective &s) {
  static const Fp2 cx = {{{0xcd03c9e48671f071ULL, 0x5dab22461fcda5d2ULL, 0x587042afd3851b95ULL, 0x8eb60ebe01bacb9eULL,
                           0x03f97d6e83d050d2ULL, 0x18f0206554638741ULL}},
                         {{0, 0, 0, 0, 0, 0}}};
  return G2Projective{fp2_mul(s.x, cx), fp2_neg(s.y), s.z};
}
// src/g2.rs:938-947, the same operator order (each `-` is add(neg))
static inline G2Projective g2p_clear_cofactor(const G2Projective &p) {
  G2Projective t1 = g2p_mul_by_x(p), t2 = g2p_psi(p);
  G2Projective r = g2p_add(g2p_psi2(g2p_double(p)), g2p_mul_by_x(g2p_add(t1, t2)));
  r = g2p_add(r, g2p_neg(t1));
  r = g2p_add(r, g2p_neg(t2));
  return g2p_add(r, g2p_neg(p));
}

// ------------------------------------------------------------------ hash_to_curve / encode_to_curve
// src/hash_to_curve/mod.rs:41-66 (hash_to_field: count elements of InputLength bytes each) and :86-108
static inline bool g1_hash(const uint8_t *msg, size_t msg_len, const uint8_t *dst, size_t dst_len, bool encode,
                           G1Projective &out) {
  uint8_t okm[128];
  const int count = encode ? 1 : 2;
  if (!expand_message_xmd_sha256(msg, msg_len, dst, dst_len, 64 * count, okm)) return false;
  G1Projective p = g1_map_to_curve(fp_from_okm(okm));
  if (!encode) p = g1p_add(p, g1_map_to_curve(fp_from_okm(okm + 64)));
  out = g1p_clear_cofactor(p);
  return true;
}
static inline bool g2_hash(const uint8_t *msg, size_t msg_len, const uint8_t *dst, size_t dst_len, bool encode,
                           G2Projective &out) {
  uint8_t okm[256];
  const int count = encode ? 1 : 2;
  if (!expand_message_xmd_sha256(msg, msg_len, dst, dst_len, 128 * count, okm)) return false;
  auto elt = [&](int i) { return Fp2{fp_from_okm(okm + 128 * i), fp_from_okm(okm + 128 * i + 64)}; };  // map_g2.rs:374
  G2Projective p = g2_map_to_curve(elt(0));
  if (!encode) p = g2p_add(p, g2_map_to_curve(elt(1)));
  out = g2p_clear_cofactor(p);
  return true;
}

// ------------------------------------------------------------------ hash to field for Scalar
// src/hash_to_curve/map_scalar.rs:10-23: 48 uniform bytes, big-endian, zero-extended to 64 and reduced like
// Scalar::from_bytes_wide (src/scalar.rs:300-331)
static inline Scalar fr_from_okm(const uint8_t okm[48]) {
  uint8_t bs[64] = {0};
  std::memcpy(bs + 16, okm, 48);
  for (int i = 0; i < 32; i++) {
    uint8_t t = bs[i];
    bs[i] = bs[63 - i];
    bs[63 - i] = t;
  }
  return fr_from_bytes_wide(bs);
}

}  // namespace bls_oracle
