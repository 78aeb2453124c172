// bls12_381.hpp — C++ host-side mirror of the zkcrypto/bls12_381 public surface for the accelerated path,
// written above the C ABI (include/bls12381_b200.h).  The reference is a Rust crate and this image has no
// Rust toolchain, so the host layer a Rust user would get from the `-sys` shim is stated
// here in C++ with the same names, argument meaning and error behaviour:
//
//   reference (Rust)                                        this header
//   -------------------------------------------------------------------------------------------------
//   &G1Projective * &Scalar            src/g1.rs:556        G1Projective::mul_batch(engine, pts, scalars)
//   iter.map(|(p,s)| p*s).sum()        src/g1.rs:573,:161   G1Projective::msm(engine, bases, scalars)
//   G1Projective::batch_normalize      src/g1.rs:806        G1Projective::batch_normalize(engine, p, q)
//   pairing(&G1Affine,&G2Affine)->Gt   src/pairings.rs:607  pairing_batch(engine, ps, qs)
//   multi_miller_loop(&[(&p,&q)])      src/pairings.rs:554  multi_miller_loop(engine, ps, qs)
//   MillerLoopResult::final_exponentiation :48              MillerLoopResult::final_exponentiation(engine)
//   products of pairings (Groth16 / BLS batch verification)  pairing_products(engine, ps, qs, terms)
//   Gt ==, +, -, double, identity       src/pairings.rs:228-337   Gt::operator==, add, neg, dbl, identity
//
// Value types are byte-compatible with the reference's in-memory values (Montgomery limbs), `Copy`-like
// PODs; fallible calls throw b200::Error carrying the ABI's code and text instead of Rust's CtOption/panic
// (batch_normalize keeps the reference's `assert_eq!(p.len(), q.len())`, src/g1.rs:807, as an exception).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/bls12381_b200.h"

namespace bls12_381 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

// One engine = one GPU (b200_ctx).  Send + Sync in the reference's sense: calls are serialised inside.
class Engine {
 public:
  explicit Engine(int device = -1) {
    int rc = b200_ctx_create(device, &ctx_);
    if (rc != B200_OK) throw Error(rc, std::string("b200_ctx_create: ") + b200_strerror(rc));
  }
  ~Engine() { b200_ctx_destroy(ctx_); }
  Engine(const Engine &) = delete;
  Engine &operator=(const Engine &) = delete;
  b200_ctx *raw() const { return ctx_; }
  void check(int rc, const char *what) const {
    if (rc != B200_OK) throw Error(rc, std::string(what) + ": " + b200_strerror(rc) + " (" + b200_last_error(ctx_) + ")");
  }

 private:
  b200_ctx *ctx_ = nullptr;
};

struct Scalar {
  b200_scalar bytes;  // == Scalar::to_bytes(): canonical little-endian, < q   (src/scalar.rs:284)
};

// An element of the scalar field in the reference's in-memory form, Scalar([u64; 4]) Montgomery limbs (src/scalar.rs:24),
// with the batched operators of the GPU path.  Fr::to_bytes gives the `Scalar` above (what msm / mul_batch consume).
struct Fr {
  b200_fr v;

  static std::vector<Fr> op(const Engine &e, int opcode, const std::vector<Fr> &a, const std::vector<Fr> *b) {
    if (b && b->size() != a.size()) throw Error(B200_EINVAL, "Fr: length mismatch");
    std::vector<Fr> out(a.size());
    e.check(b200_fr_op(e.raw(), opcode, &a.data()->v, b ? &b->data()->v : nullptr, a.size(), &out.data()->v), "fr_op");
    return out;
  }
  static std::vector<Fr> mul(const Engine &e, const std::vector<Fr> &a, const std::vector<Fr> &b) { return op(e, B200_OP_MUL, a, &b); }
  static std::vector<Fr> add(const Engine &e, const std::vector<Fr> &a, const std::vector<Fr> &b) { return op(e, B200_OP_ADD, a, &b); }
  static std::vector<Fr> sub(const Engine &e, const std::vector<Fr> &a, const std::vector<Fr> &b) { return op(e, B200_OP_SUB, a, &b); }
  static std::vector<Fr> square(const Engine &e, const std::vector<Fr> &a) { return op(e, B200_OP_SQUARE, a, nullptr); }
  static std::vector<Fr> neg(const Engine &e, const std::vector<Fr> &a) { return op(e, B200_OP_NEG, a, nullptr); }
  static std::vector<Fr> invert(const Engine &e, const std::vector<Fr> &a) { return op(e, B200_OP_INVERT, a, nullptr); }  // 0 -> 0
  static std::vector<Scalar> to_bytes(const Engine &e, const std::vector<Fr> &a) {  // src/scalar.rs:284
    std::vector<Scalar> out(a.size());
    e.check(b200_fr_to_bytes(e.raw(), &a.data()->v, a.size(), &out.data()->bytes), "fr_to_bytes");
    return out;
  }
  // src/scalar.rs:256: ok[i] = 0 for a non-canonical encoding (the reference's CtOption::None)
  static std::vector<Fr> from_bytes(const Engine &e, const std::vector<Scalar> &in, std::vector<uint8_t> &ok) {
    std::vector<Fr> out(in.size());
    ok.resize(in.size());
    e.check(b200_fr_from_bytes(e.raw(), &in.data()->bytes, in.size(), &out.data()->v, ok.data()), "fr_from_bytes");
    return out;
  }
  // in-place transform of a.size() = 2^k elements over Scalar::ROOT_OF_UNITY (src/scalar.rs:200); throws EINVAL otherwise
  static void ntt(const Engine &e, std::vector<Fr> &a, bool inverse = false, bool coset = false) {
    size_t n = a.size();
    int k = 0;
    while (((size_t)1 << k) < n) k++;
    if (n == 0 || ((size_t)1 << k) != n) throw Error(B200_EINVAL, "Fr::ntt: length must be a power of two");
    e.check(b200_fr_ntt(e.raw(), &a.data()->v, k, inverse, coset, &a.data()->v), "fr_ntt");
  }
};

struct G1Affine {
  b200_g1_affine xy;
  uint8_t infinity;  // Choice; identity() is (0, 1, infinity = 1)   (src/g1.rs:187-193)
};
struct G2Affine {
  b200_g2_affine xy;
  uint8_t infinity;
};

struct G1Projective {
  b200_g1_projective v;

  // out[i] = pts[i] * scalars[i]  — limb-exact with G1Projective::multiply (src/g1.rs:754-774)
  static std::vector<G1Projective> mul_batch(const Engine &e, const std::vector<G1Projective> &pts,
                                             const std::vector<Scalar> &scalars) {
    if (pts.size() != scalars.size()) throw Error(B200_EINVAL, "mul_batch: length mismatch");
    std::vector<G1Projective> out(pts.size());
    e.check(b200_g1_mul_batch(e.raw(), &pts.data()->v, &scalars.data()->bytes, pts.size(), &out.data()->v), "g1_mul_batch");
    return out;
  }
  // sum_i bases[i] * scalars[i]   (src/g1.rs:573-579 + Sum :161-171)
  static G1Projective msm(const Engine &e, const std::vector<G1Affine> &bases, const std::vector<Scalar> &scalars) {
    if (bases.size() != scalars.size()) throw Error(B200_EINVAL, "msm: length mismatch");
    std::vector<b200_g1_affine> xy(bases.size());
    std::vector<uint8_t> inf(bases.size());
    for (size_t i = 0; i < bases.size(); i++) {
      xy[i] = bases[i].xy;
      inf[i] = bases[i].infinity;
    }
    G1Projective out;
    e.check(b200_g1_msm(e.raw(), xy.data(), inf.data(), &scalars.data()->bytes, bases.size(), &out.v), "g1_msm");
    return out;
  }
  // G1Projective::batch_normalize(p, q): panics (throws) if p.len() != q.len()   (src/g1.rs:806-839)
  static void batch_normalize(const Engine &e, const std::vector<G1Projective> &p, std::vector<G1Affine> &q) {
    if (p.size() != q.size()) throw Error(B200_EINVAL, "batch_normalize: assertion `left == right` failed");
    std::vector<b200_g1_affine> xy(p.size());
    std::vector<uint8_t> inf(p.size());
    e.check(b200_g1_batch_normalize(e.raw(), &p.data()->v, p.size(), xy.data(), inf.data()), "g1_batch_normalize");
    for (size_t i = 0; i < p.size(); i++) q[i] = G1Affine{xy[i], inf[i]};
  }
};

struct G2Projective {
  b200_g2_projective v;
  static std::vector<G2Projective> mul_batch(const Engine &e, const std::vector<G2Projective> &pts,
                                             const std::vector<Scalar> &scalars) {
    if (pts.size() != scalars.size()) throw Error(B200_EINVAL, "mul_batch: length mismatch");
    std::vector<G2Projective> out(pts.size());
    e.check(b200_g2_mul_batch(e.raw(), &pts.data()->v, &scalars.data()->bytes, pts.size(), &out.data()->v), "g2_mul_batch");
    return out;
  }
  static G2Projective msm(const Engine &e, const std::vector<G2Affine> &bases, const std::vector<Scalar> &scalars) {
    if (bases.size() != scalars.size()) throw Error(B200_EINVAL, "msm: length mismatch");
    std::vector<b200_g2_affine> xy(bases.size());
    std::vector<uint8_t> inf(bases.size());
    for (size_t i = 0; i < bases.size(); i++) {
      xy[i] = bases[i].xy;
      inf[i] = bases[i].infinity;
    }
    G2Projective out;
    e.check(b200_g2_msm(e.raw(), xy.data(), inf.data(), &scalars.data()->bytes, bases.size(), &out.v), "g2_msm");
    return out;
  }
  static void batch_normalize(const Engine &e, const std::vector<G2Projective> &p, std::vector<G2Affine> &q) {
    if (p.size() != q.size()) throw Error(B200_EINVAL, "batch_normalize: assertion `left == right` failed");
    std::vector<b200_g2_affine> xy(p.size());
    std::vector<uint8_t> inf(p.size());
    e.check(b200_g2_batch_normalize(e.raw(), &p.data()->v, p.size(), xy.data(), inf.data()), "g2_batch_normalize");
    for (size_t i = 0; i < p.size(); i++) q[i] = G2Affine{xy[i], inf[i]};
  }
};

// <G{1,2}Projective as HashToCurve<ExpandMsgXmd<Sha256>>>::hash_to_curve / encode_to_curve for a batch of messages
// (src/hash_to_curve/mod.rs:86-108); `dst` is the domain separation tag shared by the batch
namespace detail {
inline void pack(const std::vector<std::string> &msgs, std::vector<uint8_t> &cat, std::vector<uint64_t> &off) {
  off.assign(1, 0);
  for (const auto &m : msgs) {
    cat.insert(cat.end(), m.begin(), m.end());
    off.push_back(cat.size());
  }
}
}  // namespace detail
inline std::vector<G1Projective> hash_to_curve_g1(const Engine &e, const std::vector<std::string> &msgs, const std::string &dst,
                                                  bool encode = false) {
  std::vector<uint8_t> cat;
  std::vector<uint64_t> off;
  detail::pack(msgs, cat, off);
  std::vector<G1Projective> out(msgs.size());
  e.check(b200_g1_hash_to_curve(e.raw(), cat.data(), off.data(), msgs.size(), (const uint8_t *)dst.data(), dst.size(), encode,
                                &out.data()->v), "g1_hash_to_curve");
  return out;
}
inline std::vector<G2Projective> hash_to_curve_g2(const Engine &e, const std::vector<std::string> &msgs, const std::string &dst,
                                                  bool encode = false) {
  std::vector<uint8_t> cat;
  std::vector<uint64_t> off;
  detail::pack(msgs, cat, off);
  std::vector<G2Projective> out(msgs.size());
  e.check(b200_g2_hash_to_curve(e.raw(), cat.data(), off.data(), msgs.size(), (const uint8_t *)dst.data(), dst.size(), encode,
                                &out.data()->v), "g2_hash_to_curve");
  return out;
}

struct Gt {
  b200_fp12 v;  // canonical Fp12 (src/pairings.rs:211): limbs identical to the crate's Gt.0, so == is bytewise
  bool operator==(const Gt &o) const { return std::memcmp(&v, &o.v, sizeof v) == 0; }
  bool operator!=(const Gt &o) const { return !(*this == o); }
  // Gt::identity() = Fp12::one()  (src/pairings.rs:228-231)
  static Gt identity() {
    Gt g;
    std::memset(&g.v, 0, sizeof g.v);
    const uint64_t one[6] = {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull,
                             0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull};  // R = 2^384 mod p
    std::memcpy(g.v.c[0].l, one, sizeof one);
    return g;
  }
  // the group law of Gt is multiplication in Fp12: a + b (src/pairings.rs:270-276), -a = conjugate (:253-259), double (:233-236)
  static Gt tower(const Engine &e, int op, const Gt &a, const Gt *b) {
    Gt out;
    e.check(b200_tower_op(e.raw(), 12, op, reinterpret_cast<const uint64_t *>(&a.v), b ? reinterpret_cast<const uint64_t *>(&b->v) : nullptr,
                          reinterpret_cast<uint64_t *>(&out.v), 1), "tower_op");
    return out;
  }
  Gt add(const Engine &e, const Gt &o) const { return tower(e, B200_OP_MUL, *this, &o); }
  Gt neg(const Engine &e) const { return tower(e, B200_OP_CONJUGATE, *this, nullptr); }
  Gt dbl(const Engine &e) const { return tower(e, B200_OP_SQUARE, *this, nullptr); }
  // out[i] = &g[i] * &scalars[i]   (src/pairings.rs:296-323)
  static std::vector<Gt> mul_batch(const Engine &e, const std::vector<Gt> &g, const std::vector<Scalar> &scalars) {
    if (g.size() != scalars.size()) throw Error(B200_EINVAL, "Gt::mul_batch: length mismatch");
    std::vector<Gt> out(g.size());
    e.check(b200_gt_mul_batch(e.raw(), &g.data()->v, &scalars.data()->bytes, g.size(), &out.data()->v), "gt_mul_batch");
    return out;
  }
};

struct MillerLoopResult {
  b200_fp12 v;  // src/pairings.rs:26 ; Default = Fp12::one()
  Gt final_exponentiation(const Engine &e) const {  // src/pairings.rs:48-176
    Gt out;
    e.check(b200_final_exponentiation_batch(e.raw(), &v, 1, &out.v), "final_exponentiation");
    return out;
  }
};

namespace detail {
inline void split(const std::vector<G1Affine> &ps, const std::vector<G2Affine> &qs, std::vector<b200_g1_affine> &pxy,
                  std::vector<uint8_t> &pinf, std::vector<b200_g2_affine> &qxy, std::vector<uint8_t> &qinf) {
  if (ps.size() != qs.size()) throw Error(B200_EINVAL, "pairing: length mismatch");
  pxy.resize(ps.size());
  pinf.resize(ps.size());
  qxy.resize(ps.size());
  qinf.resize(ps.size());
  for (size_t i = 0; i < ps.size(); i++) {
    pxy[i] = ps[i].xy;
    pinf[i] = ps[i].infinity;
    qxy[i] = qs[i].xy;
    qinf[i] = qs[i].infinity;
  }
}
}  // namespace detail

// out[i] = pairing(&ps[i], &qs[i])   (src/pairings.rs:607-653; identity on either side -> Gt::identity())
inline std::vector<Gt> pairing_batch(const Engine &e, const std::vector<G1Affine> &ps, const std::vector<G2Affine> &qs) {
  std::vector<b200_g1_affine> pxy;
  std::vector<b200_g2_affine> qxy;
  std::vector<uint8_t> pinf, qinf;
  detail::split(ps, qs, pxy, pinf, qxy, qinf);
  std::vector<Gt> out(ps.size());
  e.check(b200_pairing_batch(e.raw(), pxy.data(), pinf.data(), qxy.data(), qinf.data(), ps.size(), &out.data()->v),
          "pairing_batch");
  return out;
}
// multi_miller_loop(&[(&p_i, &G2Prepared::from(q_i))])   (src/pairings.rs:554-603)
inline MillerLoopResult multi_miller_loop(const Engine &e, const std::vector<G1Affine> &ps, const std::vector<G2Affine> &qs) {
  std::vector<b200_g1_affine> pxy;
  std::vector<b200_g2_affine> qxy;
  std::vector<uint8_t> pinf, qinf;
  detail::split(ps, qs, pxy, pinf, qxy, qinf);
  MillerLoopResult out;
  e.check(b200_multi_miller_loop(e.raw(), pxy.data(), pinf.data(), qxy.data(), qinf.data(), ps.size(), &out.v),
          "multi_miller_loop");
  return out;
}

// n_products independent products of `terms` consecutive pairs each — the shape of Groth16 / BLS batch verification:
//   (0..n).map(|i| multi_miller_loop(&terms[i]).final_exponentiation())   with ONE squaring per bit shared by the terms of a
// product (src/pairings.rs:554-603), ps.size() == qs.size() == terms * n_products
inline std::vector<Gt> pairing_products(const Engine &e, const std::vector<G1Affine> &ps, const std::vector<G2Affine> &qs, size_t terms) {
  if (terms == 0 || ps.size() % terms != 0) throw Error(B200_EINVAL, "pairing_products: the number of pairs must be a multiple of terms");
  std::vector<b200_g1_affine> pxy;
  std::vector<b200_g2_affine> qxy;
  std::vector<uint8_t> pinf, qinf;
  detail::split(ps, qs, pxy, pinf, qxy, qinf);
  std::vector<Gt> out(ps.size() / terms);
  e.check(b200_pairing_product_batch(e.raw(), pxy.data(), pinf.data(), qxy.data(), qinf.data(), terms, out.size(), 1, &out.data()->v),
          "pairing_product_batch");
  return out;
}

static_assert(sizeof(G1Projective) == 144 && sizeof(G2Projective) == 288 && sizeof(Gt) == 576 && sizeof(Scalar) == 32,
              "value types must be layout-compatible with the ABI structs");

}  // namespace bls12_381
