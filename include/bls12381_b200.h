/* bls12381_b200 — C ABI of the B200-native BLS12-381 hot path (libbls12381_b200.so).
 *
 * This is the drop-in boundary for the data-parallel hot path of zkcrypto/bls12_381 v0.8.0
 * (SURVEY.md §8b): batched scalar multiplication, G1/G2 multi-scalar multiplication, batched
 * multi_miller_loop / final_exponentiation / pairing, and batch_normalize.  The reference crate has
 * no FFI of its own (#![deny(unsafe_code)], src/lib.rs:17); each entry point below names the
 * reference function it replaces, and INTEGRATION.md shows the Rust `-sys` binding that calls it.
 *
 * Data layouts mirror the reference's in-memory values exactly (all little-endian):
 *   Fp     = Fp([u64;6])   Montgomery form, R = 2^384, canonical (< p)          src/fp.rs:15
 *   Fp2    = {c0, c1}                                                           src/fp2.rs:11
 *   Fp12   = 12 Fp in the order c0.c0.c0, c0.c0.c1, c0.c1.c0, ... c1.c2.c1      src/fp12.rs:13
 *   scalar = Scalar::to_bytes(): canonical 32-byte little-endian integer < q    src/scalar.rs:284
 * Affine points are passed as coordinate arrays plus an optional infinity-flag array (one byte per
 * point, non-zero = identity; NULL = no identities).  A set flag means G*Affine::identity()
 * (x = 0, y = 1) whatever the coordinate bytes hold.  Rust structs are not repr(C), so the Rust shim
 * marshals field by field into these layouts.
 *
 * Conventions: every function returns 0 on success or a negative B200_E* code; nothing aborts,
 * throws or longjmps across the boundary.  The caller owns all buffers; the library keeps no pointer
 * after return.  A b200_ctx owns one CUDA device, one stream and grow-only scratch memory; calls on
 * one ctx are serialised by an internal mutex, distinct ctxs are independent (Send + Sync on the Rust
 * side).  Functions without a suffix take HOST pointers (copies included); `_dev` variants take
 * DEVICE pointers on the ctx's device, enqueue on the ctx stream and synchronise it before returning
 * unless stated otherwise.  No CPU fallback exists: without a usable CUDA device every call fails
 * with B200_ENODEV.
 */
#ifndef BLS12381_B200_H
#define BLS12381_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[6]; } b200_fp;                 /* 48 B  */
typedef struct { b200_fp c0, c1; } b200_fp2;               /* 96 B  */
typedef struct { b200_fp c[12]; } b200_fp12;               /* 576 B: Gt / MillerLoopResult */
typedef struct { b200_fp x, y; } b200_g1_affine;           /* 96 B  G1Affine coords   src/g1.rs:28  */
typedef struct { b200_fp x, y, z; } b200_g1_projective;    /* 144 B G1Projective      src/g1.rs:442 */
typedef struct { b200_fp2 x, y; } b200_g2_affine;          /* 192 B G2Affine coords   src/g2.rs:29  */
typedef struct { b200_fp2 x, y, z; } b200_g2_projective;   /* 288 B G2Projective      src/g2.rs:495 */
typedef struct { uint8_t b[32]; } b200_scalar;             /* Scalar::to_bytes()      src/scalar.rs:284 */

typedef struct b200_ctx b200_ctx;

enum {
  B200_OK = 0,
  B200_EINVAL = -1,  /* NULL pointer / bad size / bad op */
  B200_ENODEV = -2,  /* no usable CUDA device */
  B200_ECUDA = -3,   /* CUDA runtime error (see b200_last_error) */
  B200_ENOMEM = -4,
  B200_ENCCL = -5    /* NCCL missing (libnccl.so.2 could not be loaded) or an NCCL call failed */
};

/* ---- context ------------------------------------------------------------------------------- */
int b200_ctx_create(int device /* CUDA ordinal, -1 = current */, b200_ctx **out);
/* same, but the ctx enqueues its work on the CALLER's stream (a cudaStream_t of `device`), which stays owned by the caller
 * and must outlive the ctx — lets a host runtime (e.g. a PyTorch process) keep one stream for its own copies / events and
 * the library's kernels.  stream = NULL is b200_ctx_create. */
int b200_ctx_create_on_stream(int device, void *stream, b200_ctx **out);
void b200_ctx_destroy(b200_ctx *ctx);
const char *b200_strerror(int code);
const char *b200_last_error(const b200_ctx *ctx); /* text of the last CUDA error on this ctx */
int b200_ctx_device(const b200_ctx *ctx);
/* stream the _dev calls are enqueued on (a cudaStream_t), for callers that time with CUDA events */
void *b200_ctx_stream(const b200_ctx *ctx);
/* number of kernels this ctx has launched since creation (bench.py's gpu_launches) */
uint64_t b200_ctx_launch_count(const b200_ctx *ctx);
/* per-kernel timing: when on, every kernel launch of this ctx is bracketed by CUDA events on the ctx
 * stream.  get_timing synchronises, writes up to `max` durations (ms) and the kernel names joined by
 * '\n' into `names`, and returns the number of records since the last set_timing call (>= 0). */
int b200_ctx_set_timing(b200_ctx *ctx, int on);
int b200_ctx_get_timing(b200_ctx *ctx, char *names, size_t names_len, float *ms, int max);
/* tuning knobs by name (defaults are the measured best on B200; DESIGN.md has the numbers).  Unknown key or bad value -> B200_EINVAL.
 *   MSM:      "msm_window" (0 = auto, else 2..24) · "g1_glv" (0 off = default, 1 on, 2 on for window-sharded calls) ·
 *             "msm_affine_levels" (-1 auto, 0..3 batched-affine tree levels before the bucket kernel; default 0) · "g1_prefetch" (0|1) ·
 *             "msm_reduce" (bucket reduction: 0 one thread per chunk, 1 lane-cooperative for the exposed last window group = default
 *             (also -1), 2 lane-cooperative everywhere) · "msm_reduce_min_chunk" (1..64 buckets per thread at least, default 8) ·
 *             "msm_tail_groups" (2-3 local windows: 0 one group = default, 1 one window per group) ·
 *             "g2_acc_blocks" (G2 bucket kernel: 2 registers, 3 shared-memory accumulator built for 3 blocks/SM, 4 shared-memory accumulator
 *             at 2 blocks/SM = default)
 *   pairing:  "pairing_variant" (0 = six lanes per pairing at every batch size = default, 7 the same explicitly, 4 = one thread per pairing) ·
 *             "coop_warps" (1..12 warps per block of the six-lane kernels) · "coop_split" (Miller loop and final exponentiation as two
 *             launches, default 1) · "coop_chunks" (1..64 whole-wave chunks on two streams for batches of more than two waves, default 3) ·
 *             "coop_prepare_max" (largest batch whose G2 line coefficients come from the six-lanes-per-Q kernel, default 5000) ·
 *             "pairing_chunks" (1..64 chunks of a batch of the one-thread-per-pairing kernels)
 *   scalar multiplication batches: "mul_groups" (-1 thread per item, 0 auto = default, 1..5 items per warp of the six-lane group kernel,
 *             6 one warp per item) · "mul_groups_max_n" (largest batch for the group kernel, default 9000) */
int b200_ctx_set_tuning(b200_ctx *ctx, const char *key, int value);
/* MSM tuning: window bits c (0 = automatic from n, else 2..24); B200_OK or B200_EINVAL (same as set_tuning("msm_window")) */
int b200_ctx_set_msm_window(b200_ctx *ctx, int c);

/* ---- field tower, batched (diagnostic / parity surface for src/fp.rs, fp2.rs, fp6.rs, fp12.rs) - */
/* level: 1 Fp, 2 Fp2, 6 Fp6, 12 Fp12.  op: */
enum {
  B200_OP_MUL = 0, B200_OP_ADD = 1, B200_OP_SUB = 2, B200_OP_SQUARE = 3, B200_OP_NEG = 4,
  B200_OP_INVERT = 5, B200_OP_FROBENIUS = 6, B200_OP_CONJUGATE = 7, B200_OP_MUL_BY_NONRESIDUE = 8,
  B200_OP_CYCLOTOMIC_SQUARE = 9,
  B200_OP_INVERT_FAST = 10 /* Fp only: binary-GCD inverse (csrc/fp_inv.cuh), same value as B200_OP_INVERT */
};
/* out[i] = op(a[i], b[i]); b may be NULL for unary ops; arrays of n elements of 6*level u64 each */
int b200_tower_op(b200_ctx *ctx, int level, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);

/* ---- G1 ------------------------------------------------------------------------------------- */
/* out[i] = p[i] * s[i]  — G1Projective::multiply, src/g1.rs:754-774 (via Mul<&Scalar> :556-562).
 * Raw (x,y,z) limbs are bit-identical to the reference (same complete formulas, same 255 steps). */
int b200_g1_mul_batch(b200_ctx *ctx, const b200_g1_projective *p, const b200_scalar *s, size_t n, b200_g1_projective *out);
/* element-wise group ops, limb-exact: double :638, add :670, add_mixed :715 */
int b200_g1_double_batch(b200_ctx *ctx, const b200_g1_projective *p, size_t n, b200_g1_projective *out);
int b200_g1_add_batch(b200_ctx *ctx, const b200_g1_projective *p, const b200_g1_projective *q, size_t n, b200_g1_projective *out);
int b200_g1_add_mixed_batch(b200_ctx *ctx, const b200_g1_projective *p, const b200_g1_affine *q, const uint8_t *q_inf, size_t n, b200_g1_projective *out);
/* G1Projective::batch_normalize, src/g1.rs:806-839 (same values as per-point G1Affine::from :49-63) */
int b200_g1_batch_normalize(b200_ctx *ctx, const b200_g1_projective *p, size_t n, b200_g1_affine *out, uint8_t *out_inf);
/* MSM: out = sum_i points[i] * scalars[i]  — what the reference expresses as
 * points.zip(scalars).map(|(p,s)| p*s).sum()  (src/g1.rs:573-579, :161-171).  The result is the same
 * GROUP ELEMENT; it is returned in projective form and is bit-identical after to_affine. */
int b200_g1_msm(b200_ctx *ctx, const b200_g1_affine *points, const uint8_t *inf /* nullable */, const b200_scalar *scalars, size_t n, b200_g1_projective *out);

/* ---- G2 (same shapes over Fp2; src/g2.rs:825-845, :709, :741, :786, :951-984, :609-615/:162) --- */
int b200_g2_mul_batch(b200_ctx *ctx, const b200_g2_projective *p, const b200_scalar *s, size_t n, b200_g2_projective *out);
int b200_g2_double_batch(b200_ctx *ctx, const b200_g2_projective *p, size_t n, b200_g2_projective *out);
int b200_g2_add_batch(b200_ctx *ctx, const b200_g2_projective *p, const b200_g2_projective *q, size_t n, b200_g2_projective *out);
int b200_g2_add_mixed_batch(b200_ctx *ctx, const b200_g2_projective *p, const b200_g2_affine *q, const uint8_t *q_inf, size_t n, b200_g2_projective *out);
int b200_g2_batch_normalize(b200_ctx *ctx, const b200_g2_projective *p, size_t n, b200_g2_affine *out, uint8_t *out_inf);
int b200_g2_msm(b200_ctx *ctx, const b200_g2_affine *points, const uint8_t *inf, const b200_scalar *scalars, size_t n, b200_g2_projective *out);

/* ---- pairings (src/pairings.rs) ------------------------------------------------------------- */
/* out[i] = Miller loop of (p[i], q[i]) as in pairing() :607-646 — unprepared, with its identity
 * handling (either side identity -> Fp12::one()).  MillerLoopResult limbs are bit-identical. */
int b200_miller_loop_batch(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *p_inf, const b200_g2_affine *q, const uint8_t *q_inf, size_t n, b200_fp12 *out);
/* out[i] = MillerLoopResult(in[i]).final_exponentiation()  :48-176  (f^(3(p^12-1)/r)) */
int b200_final_exponentiation_batch(b200_ctx *ctx, const b200_fp12 *in, size_t n, b200_fp12 *out);
/* out[i] = pairing(&p[i], &q[i])  :607-653  — n independent Gt values */
int b200_pairing_batch(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *p_inf, const b200_g2_affine *q, const uint8_t *q_inf, size_t n, b200_fp12 *gt_out);
/* out = multi_miller_loop(&[(p_i, G2Prepared::from(q_i))])  :554-603 — ONE MillerLoopResult for the
 * product.  Computed as the Fp12 product of the per-pair Miller values, which is the same field
 * element the reference's shared-squaring loop produces (f <- f^2 * prod_t l_t is multiplicative), so
 * the limbs are bit-identical; identity terms contribute one(), like the reference's skip :566-569. */
int b200_multi_miller_loop(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *p_inf, const b200_g2_affine *q, const uint8_t *q_inf, size_t n, b200_fp12 *out);
/* host-pointer form of b200_pairing_product_batch_dev (see there): product i = pairs [i * terms, (i + 1) * terms) */
int b200_pairing_product_batch(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *p_inf, const b200_g2_affine *q, const uint8_t *q_inf, size_t terms, size_t n_products, int final_exp, b200_fp12 *out);

/* ---- G2Prepared (src/pairings.rs:498-546): coeffs = n x 68 x (Fp2, Fp2, Fp2) = 19 584 B per Q, in
 * the order the Miller loop consumes them; the identity is prepared as the generator (the caller keeps q_inf, :528-544).
 * b200_multi_miller_loop_prepared = multi_miller_loop(&[(&p_i, &prepared_i)]) (:554-603), limb-exact; terms with p_i or
 * q_i at infinity contribute one().  The _dev variants keep the coefficients resident in HBM for fixed verifying keys. */
int b200_g2_prepare(b200_ctx *ctx, const b200_g2_affine *q, const uint8_t *q_inf, size_t n, b200_fp2 *coeffs);
int b200_multi_miller_loop_prepared(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *p_inf, const b200_fp2 *coeffs, const uint8_t *q_inf, size_t n, b200_fp12 *out);
int b200_g2_prepare_dev(b200_ctx *ctx, const void *q, const void *q_inf, size_t n, void *coeffs);
int b200_miller_loop_prepared_batch_dev(b200_ctx *ctx, const void *p, const void *p_inf, const void *coeffs, const void *q_inf, size_t n, void *out);

/* ---- point (de)serialization on the device (wire format src/notes/serialization.rs) -----
 * serialize: out[i] = G1Affine::to_compressed (48 B) / to_uncompressed (96 B)  src/g1.rs:221-260
 *            (G2: 96 / 192 B, Fp2 as c1 || c0                                   src/g2.rs:254-299)
 * deserialize: G1Affine::from_compressed_unchecked / from_uncompressed_unchecked (src/g1.rs:275-390,
 *            src/g2.rs:313-464).  status[i] bit 0 = the reference constructor would return Some (canonical
 *            field encodings, consistent flags, x on the curve for compressed input); bit 1 = is_on_curve
 *            (src/g1.rs:414).  Rejected inputs yield the identity.  The subgroup test is b200_g{1,2}_check. */
/* status[i] bit 0 = G1Affine::is_on_curve (src/g1.rs:414), bit 1 = is_torsion_free (src/g1.rs:401-410: endomorphism(P) == -[x^2]P;
 * G2 src/g2.rs:475-482: psi(P) == [x]P).  deserialize + check == from_compressed / from_uncompressed (checked). */
int b200_g1_check(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *inf, size_t n, uint8_t *status);
int b200_g2_check(b200_ctx *ctx, const b200_g2_affine *p, const uint8_t *inf, size_t n, uint8_t *status);
int b200_g1_serialize(b200_ctx *ctx, const b200_g1_affine *p, const uint8_t *inf, size_t n, int compressed, uint8_t *out);
int b200_g2_serialize(b200_ctx *ctx, const b200_g2_affine *p, const uint8_t *inf, size_t n, int compressed, uint8_t *out);
int b200_g1_deserialize(b200_ctx *ctx, const uint8_t *in, size_t n, int compressed, b200_g1_affine *out, uint8_t *out_inf, uint8_t *status);
int b200_g2_deserialize(b200_ctx *ctx, const uint8_t *in, size_t n, int compressed, b200_g2_affine *out, uint8_t *out_inf, uint8_t *status);

/* ---- device-pointer variants (inputs already resident in HBM; used by bench.py `value`) ------- */
int b200_g1_mul_batch_dev(b200_ctx *ctx, const void *p, const void *s, size_t n, void *out);
int b200_g2_mul_batch_dev(b200_ctx *ctx, const void *p, const void *s, size_t n, void *out);
int b200_g1_batch_normalize_dev(b200_ctx *ctx, const void *p, size_t n, void *out_xy, void *out_inf);
int b200_g2_batch_normalize_dev(b200_ctx *ctx, const void *p, size_t n, void *out_xy, void *out_inf);
int b200_g1_msm_dev(b200_ctx *ctx, const void *points, const void *inf, const void *scalars, size_t n, void *out);
int b200_g2_msm_dev(b200_ctx *ctx, const void *points, const void *inf, const void *scalars, size_t n, void *out);
/* window-sharded MSM (north_star: "MSM shards by scalar-window across up to 8 GPUs"): processes only
 * windows w with w % n_shards == shard and writes sum_w 2^(c*w) * S_w for those windows — a partial
 * group element; the partials of all shards add up (complete add) to the full MSM. */
int b200_g1_msm_shard_dev(b200_ctx *ctx, const void *points, const void *inf, const void *scalars, size_t n, int shard, int n_shards, void *out);
int b200_g2_msm_shard_dev(b200_ctx *ctx, const void *points, const void *inf, const void *scalars, size_t n, int shard, int n_shards, void *out);
/* out = sum_i parts[i] (complete projective adds, src/g1.rs:161-171): combines gathered partials */
int b200_g1_sum_dev(b200_ctx *ctx, const void *parts, size_t n, void *out);
int b200_g2_sum_dev(b200_ctx *ctx, const void *parts, size_t n, void *out);

/* ---- multi-GPU ("the ctx owns CUDA streams, scratch device memory and NCCL communicators"; §8e) ---------------
 * (a) one process (or thread) per GPU.  Rank 0 calls b200_comm_unique_id and ships the 128 bytes to the other ranks
 * (MPI, socket, torch.distributed ...); every rank then calls b200_ctx_comm_init on its own ctx.  After that
 * b200_g{1,2}_msm_sharded_dev is a COLLECTIVE call (all ranks, same n, same mode): every rank computes its shard,
 * the 144 / 288-byte partial sums are exchanged with one ncclAllGather over NVLink and added with complete additions
 * (NCCL has no elliptic-curve reduction) — shard, collective and combine are enqueued on the ctx stream back to back,
 * the host synchronises once at the end.  `out` (device) holds the full sum on EVERY rank.
 *   B200_SHARD_WINDOWS: north_star's scalar-window sharding — every rank holds ALL n points and scalars and handles the
 *                       windows w = rank (mod world).
 *   B200_SHARD_POINTS:  every rank holds all n points/scalars too (same pointers semantics) but only READS its
 *                       contiguous slice [rank n / world, (rank+1) n / world): all windows of a slice of the points.
 * Pairing batches shard by pair index with no collective: the caller passes each rank its slice. */
#define B200_COMM_ID_BYTES 128
#define B200_SHARD_WINDOWS 0
#define B200_SHARD_POINTS 1
int b200_comm_unique_id(uint8_t *id /* B200_COMM_ID_BYTES */);
int b200_ctx_comm_init(b200_ctx *ctx, const uint8_t *id /* B200_COMM_ID_BYTES */, int rank, int world);
int b200_ctx_comm_destroy(b200_ctx *ctx);
int b200_ctx_comm_rank(const b200_ctx *ctx);
int b200_ctx_comm_world(const b200_ctx *ctx);
int b200_g1_msm_sharded_dev(b200_ctx *ctx, const void *points, const void *inf, const void *scalars, size_t n, int mode, void *out);
int b200_g2_msm_sharded_dev(b200_ctx *ctx, const void *points, const void *inf, const void *scalars, size_t n, int mode, void *out);
/* (b) one process, n_gpus devices (0 = all visible): owns one ctx, one NCCL communicator (ncclCommInitAll) and one host
 * thread per device.  b200_multi_g{1,2}_msm take HOST pointers: with B200_SHARD_POINTS (default) every device receives
 * only its slice of the points and scalars (H2D scales with the number of GPUs), with B200_SHARD_WINDOWS all of them. */
typedef struct b200_multi b200_multi;
int b200_multi_create(int n_gpus, b200_multi **out);
void b200_multi_destroy(b200_multi *m);
int b200_multi_gpus(const b200_multi *m);
b200_ctx *b200_multi_ctx(b200_multi *m, int i);
int b200_multi_set_sharding(b200_multi *m, int mode);
int b200_multi_g1_msm(b200_multi *m, const b200_g1_affine *points, const uint8_t *inf, const b200_scalar *scalars, size_t n, b200_g1_projective *out);
int b200_multi_g2_msm(b200_multi *m, const b200_g2_affine *points, const uint8_t *inf, const b200_scalar *scalars, size_t n, b200_g2_projective *out);

int b200_miller_loop_batch_dev(b200_ctx *ctx, const void *p, const void *p_inf, const void *q, const void *q_inf, size_t n, void *out);
/* multi_miller_loop over n terms -> ONE MillerLoopResult, device pointers (src/pairings.rs:554-603) */
int b200_multi_miller_loop_dev(b200_ctx *ctx, const void *p, const void *p_inf, const void *q, const void *q_inf, size_t n, void *out);
/* n_products independent products of `terms` consecutive pairs each, with ONE squaring of the Miller accumulator per bit for
 * all terms of a product (the shared-squaring loop of src/pairings.rs:554-603): out[i] = multi_miller_loop(pairs of product i)
 * (final_exp = 0) or its final_exponentiation (final_exp = 1).  The Groth16 / BLS batch-verification shape. */
int b200_pairing_product_batch_dev(b200_ctx *ctx, const void *p, const void *p_inf, const void *q, const void *q_inf, size_t terms, size_t n_products, int final_exp, void *out);
int b200_final_exponentiation_batch_dev(b200_ctx *ctx, const void *in, size_t n, void *out);
int b200_pairing_batch_dev(b200_ctx *ctx, const void *p, const void *p_inf, const void *q, const void *q_inf, size_t n, void *gt_out);
/* out = product of the n Fp12 values (MillerLoopResult `+`, src/pairings.rs:179-186) */
int b200_fp12_product_dev(b200_ctx *ctx, const void *in, size_t n, void *out);

/* test surface for the GLV scalar decomposition used by the G1 MSM (csrc/glv.cuh): k = k1 + k2*lambda (mod q),
 * |k1|, |k2| < 2^127.  k1k2[i] = k1 magnitude (16 B LE) | k2 magnitude (16 B LE); signs[i] bit 0 = k1 < 0,
 * bit 1 = k2 < 0. */
int b200_glv_decompose(b200_ctx *ctx, const b200_scalar *scalars, size_t n, uint8_t *k1k2, uint8_t *signs);

/* ---- Gt * Scalar (`&Gt * &Scalar`, src/pairings.rs:296-323): out[i] = g[i]^scalars[i] in Fp12, double-and-add over the
 * canonical 32-byte little-endian scalar, limb-identical to the reference for any Fp12 input.  Gt Add / Neg / double /
 * Sum are b200_tower_op (MUL / CONJUGATE / SQUARE at level 12) and b200_fp12_product_dev. */
int b200_gt_mul_batch(b200_ctx *ctx, const b200_fp12 *g, const b200_scalar *scalars, size_t n, b200_fp12 *out);
int b200_gt_mul_batch_dev(b200_ctx *ctx, const void *g, const void *scalars, size_t n, void *out);

/* ---- scalar field Fr: batched arithmetic and the NTT -------------------------------------
 * Elements are b200_fr = Scalar([u64; 4]) (src/scalar.rs:24): little-endian limbs, Montgomery form R = 2^256,
 * canonical (< q).  b200_fr_op: out[i] = a[i] (op) b[i]; op = B200_OP_MUL (src/scalar.rs:554), _ADD (:600), _SUB (:582),
 * _SQUARE (:341), _NEG (:613), _INVERT (:408; 0 -> 0 where the reference returns CtOption::None), B200_OP_DOUBLE
 * (:249); b is ignored (may be NULL) for the unary ops.
 * b200_fr_to_bytes = Scalar::to_bytes (:284): canonical 32-byte little-endian integers — the b200_scalar the MSM and
 * scalar-multiplication entry points consume.  b200_fr_from_bytes = Scalar::from_bytes (:256): ok[i] = 1 and the
 * Montgomery limbs when the encoding is canonical, else ok[i] = 0 and zero limbs.
 * b200_fr_ntt: the transform every FFT over this field is built on (the reference exports its parameters:
 * ROOT_OF_UNITY :200, S = 32 :191, MULTIPLICATIVE_GENERATOR = 7 :100), n = 2^log_n, natural order in and out:
 *   forward  out[k] = sum_j in[j] w^(jk),  w = ROOT_OF_UNITY^(2^(32 - log_n));
 *   inverse  out[j] = n^-1 sum_k in[k] w^(-jk);
 *   coset    forward first scales in[j] by g^j, inverse finally scales out[j] by g^-j (g = 7): evaluation on /
 *            interpolation from the coset g<w>  (bellman EvaluationDomain::coset_fft / icoset_fft).
 * 0 <= log_n <= 28.  `in` and `out` may be the same buffer.  The n/2-entry twiddle table (16 n bytes of device
 * memory) is cached in the ctx per log_n. */
typedef struct { uint64_t l[4]; } b200_fr;      /* 32 B  src/scalar.rs:24 */
#define B200_OP_DOUBLE 11
int b200_fr_op(b200_ctx *ctx, int op, const b200_fr *a, const b200_fr *b, size_t n, b200_fr *out);
int b200_fr_to_bytes(b200_ctx *ctx, const b200_fr *a, size_t n, b200_scalar *out);
int b200_fr_from_bytes(b200_ctx *ctx, const b200_scalar *in, size_t n, b200_fr *out, uint8_t *ok);
int b200_fr_ntt(b200_ctx *ctx, const b200_fr *in, int log_n, int inverse, int coset, b200_fr *out);
/* device-pointer variants (32-byte aligned device memory; run on the ctx stream, not synchronised) */
int b200_fr_op_dev(b200_ctx *ctx, int op, const void *a, const void *b, size_t n, void *out);
int b200_fr_ntt_dev(b200_ctx *ctx, const void *in, int log_n, int inverse, int coset, void *out);

/* ---- batched hash to curve: the RFC 9380 suites BLS12381G1_XMD:SHA-256_SSWU_{RO,NU}_ and
 * BLS12381G2_XMD:SHA-256_SSWU_{RO,NU}_, i.e. <G1Projective as HashToCurve<ExpandMsgXmd<Sha256>>>::hash_to_curve /
 * encode_to_curve (src/hash_to_curve/mod.rs:86-92, :103-108) and the G2 equivalents.
 * Messages are passed concatenated: message i = msgs[offsets[i] .. offsets[i+1]) (offsets has n + 1 entries, ascending,
 * offsets[0] may be non-zero); the domain separation tag `dst` is shared by the batch (longer than 255 bytes: reduced
 * as RFC 9380 5.3.3 / src/hash_to_curve/expand_msg.rs:64-84).  encode = 0: hash_to_curve (random oracle, two field
 * elements), 1: encode_to_curve (non-uniform, one).  out[i] = the projective point, limb-identical to the reference's.
 * b200_expand_message_xmd_sha256: out[i * len_in_bytes ..] = expand_message_xmd(msg_i, dst, len_in_bytes)
 * (src/hash_to_curve/expand_msg.rs:230-300); B200_EINVAL where the reference panics (ceil(len/32) > 255).
 * b200_h2c_stage: the steps on their own (parity surface).  group 1 / 2; kind 0 map_to_curve_simple_swu (field element ->
 * point of the isogenous curve, src/hash_to_curve/map_g1.rs:550 / map_g2.rs:391), 1 iso_map (:589 / :457), 2 map_to_curve
 * (:635 / :497), 3 clear_cofactor (src/g1.rs:800, src/g2.rs:938); in: n field elements (kind 0, 2) or n projective
 * points (kind 1, 3); out: n projective points. */
int b200_expand_message_xmd_sha256(b200_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len, size_t len_in_bytes, uint8_t *out);
int b200_g1_hash_to_curve(b200_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len, int encode, b200_g1_projective *out);
int b200_g2_hash_to_curve(b200_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len, int encode, b200_g2_projective *out);
int b200_h2c_stage(b200_ctx *ctx, int group, int kind, const void *in, size_t n, void *out);
/* hash to field for Scalar (src/hash_to_curve/map_scalar.rs:10-23, src/hash_to_curve/mod.rs:41-66): b200_fr_from_okm maps
 * n x 48 uniform bytes to n scalars (Scalar::from_okm); b200_fr_hash_to_field writes `count` scalars per message,
 * out[i * count + c], from expand_message_xmd(msg_i, dst, 48 * count).  1 <= count <= 170. */
int b200_fr_from_okm(b200_ctx *ctx, const uint8_t *okm, size_t n, b200_fr *out);
int b200_fr_hash_to_field(b200_ctx *ctx, const uint8_t *msgs, const uint64_t *offsets, size_t n, const uint8_t *dst, size_t dst_len, int count, b200_fr *out);

/* ---- measurement helper: dependent-free IMAD.WIDE.U32 stream on all SMs; returns achieved
 * 32x32+64 multiply-adds per second (the integer roofline denominator) ------------ */
int b200_imad_peak(b200_ctx *ctx, int iters, double *imad_per_sec, double *ms);
/* mode 0 = the above (mad.wide.u32 -> IMAD.WIDE.U32); mode 1 = the same 64-bit multiply-accumulate as a mad.lo.cc / madc.hi
 * pair (two multiplier instructions per product), counted in the same unit — the audit of the roofline denominator */
int b200_imad_peak_mode(b200_ctx *ctx, int iters, int mode, double *imad_per_sec, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* BLS12381_B200_H */
