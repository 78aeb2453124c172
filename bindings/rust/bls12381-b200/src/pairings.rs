//! The pairing half of the boundary: `pairing` (src/pairings.rs:607), `multi_miller_loop` (:554),
//! `MillerLoopResult::final_exponentiation` (:48) and `impl Engine / MultiMillerLoop for Bls12` (:795-824) on the GPU.
//!
//! THE MARSHALLING PROBLEM.  `Gt(pub(crate) Fp12)` (src/pairings.rs:211) and `MillerLoopResult(pub(crate) Fp12)` (:26) have
//! no public constructor and no byte encoding, so — unlike points and scalars, which travel as their canonical byte
//! encodings — an Fp12 computed outside the crate cannot be turned into a `bls12_381::Gt`.  Two mechanisms, both here:
//!
//!  (A) feature "raw-gt" (recommended): the THREE-LINE accessor patch below against bls12_381 0.8.0.  With it
//!      `Engine::pairing_batch` returns the crate's own `Vec<Gt>`, and `Bls12B200` is a `pairing::Engine` /
//!      `MultiMillerLoop` whose associated types are the crate's (`Gt`, `MillerLoopResult`, `G2Prepared`), so a prover
//!      generic over `E: MultiMillerLoop` only swaps `Bls12` for `Bls12B200`.
//!
//!      --- a/src/pairings.rs
//!      +++ b/src/pairings.rs
//!      @@ impl MillerLoopResult {            (after :47)
//!      +    /// Raw limbs for accelerator backends: the 12 Fp of the Fp12 in struct order c0.c0.c0 .. c1.c2.c1 (Montgomery form).
//!      +    pub fn from_raw_unchecked(l: [[u64; 6]; 12]) -> Self { MillerLoopResult(Fp12::from_raw_unchecked(l)) }
//!      +    pub fn to_raw(&self) -> [[u64; 6]; 12] { self.0.to_raw() }
//!      @@ impl Gt {                          (after :226)
//!      +    pub fn from_raw_unchecked(l: [[u64; 6]; 12]) -> Self { Gt(Fp12::from_raw_unchecked(l)) }
//!      +    pub fn to_raw(&self) -> [[u64; 6]; 12] { self.0.to_raw() }
//!      --- a/src/fp12.rs
//!      +++ b/src/fp12.rs
//!      @@ impl Fp12 {
//!      +    pub(crate) fn from_raw_unchecked(l: [[u64; 6]; 12]) -> Self {
//!      +        let f = |i: usize| Fp::from_raw_unchecked(l[i]);          // src/fp.rs:302
//!      +        let f2 = |i: usize| Fp2 { c0: f(i), c1: f(i + 1) };
//!      +        Fp12 { c0: Fp6 { c0: f2(0), c1: f2(2), c2: f2(4) }, c1: Fp6 { c0: f2(6), c1: f2(8), c2: f2(10) } }
//!      +    }
//!      +    pub(crate) fn to_raw(&self) -> [[u64; 6]; 12] { /* the same twelve `.0` arrays in the same order */ }
//!
//!      (`Fp.0` is `pub(crate)`, src/fp.rs:15; `b200_fp12` IS that array of twelve `[u64; 6]`, include/bls12381_b200.h.)
//!
//!  (B) without the patch: the opaque `B200Gt` / `B200MillerLoopResult` below hold the 576 bytes the GPU produced and
//!      implement the operations callers perform on the crate's types — `==` (limbs are canonical, so equality is
//!      bytewise), `+` / `-` / `Neg` / `double` / `Sum` (src/pairings.rs:253-337; Fp12 mul / conjugate on the device),
//!      `* Scalar` (:296-323, `b200_gt_mul_batch`), `identity()`, `generator()`, `final_exponentiation()`.  A verifier
//!      that only compares pairing products (`e(a,b) * e(c,d) == Gt::identity()`, Groth16 / BLS signatures) needs nothing
//!      else.  They are NOT `bls12_381::Gt`; code that must hand a `Gt` to a third crate needs (A).
//!
//! NOT compiled in this repository's build container (no Rust toolchain); the C++ mirror of the same call sequences is
//! compiled and run by tests/test_host_cpp.py, and the device entry points are parity-tested against the oracle.
use crate::{check, Engine, Error};
use bls12_381::{G1Affine, G2Affine, Scalar};
use bls12381_b200_sys as sys;
use core::ops::{Add, AddAssign, Mul, Neg, Sub};

const FP_ZERO: sys::b200_fp = sys::b200_fp { l: [0; 6] };
/// R = 2^384 mod p, i.e. Fp::one() (src/fp.rs:83-90)
const FP_ONE: sys::b200_fp = sys::b200_fp {
    l: [0x760900000002fffd, 0xebf4000bc40c0002, 0x5f48985753c758ba, 0x77ce585370525745, 0x5c071a97a256ec6d, 0x15f65ec3fa80e493],
};
// b200_tower_op codes (include/bls12381_b200.h)
const OP_MUL: i32 = 0;
const OP_SQUARE: i32 = 3;
const OP_CONJUGATE: i32 = 7;

fn fp12_one() -> sys::b200_fp12 {
    let mut c = [FP_ZERO; 12];
    c[0] = FP_ONE;
    sys::b200_fp12 { c }
}
fn fp12_eq(a: &sys::b200_fp12, b: &sys::b200_fp12) -> bool {
    a.c.iter().zip(b.c.iter()).all(|(x, y)| x.l == y.l) // canonical Montgomery limbs: equal values <=> equal limbs
}

/// (points, infinity flags) in the ABI's layout, decoded on the GPU from the canonical uncompressed encodings
fn marshal_pairs(eng: &Engine, ps: &[G1Affine], qs: &[G2Affine])
    -> Result<(Vec<sys::b200_g1_affine>, Vec<u8>, Vec<sys::b200_g2_affine>, Vec<u8>), Error> {
    assert_eq!(ps.len(), qs.len());
    let n = ps.len();
    let z2 = sys::b200_fp2 { c0: FP_ZERO, c1: FP_ZERO };
    let pb: Vec<u8> = ps.iter().flat_map(|p| p.to_uncompressed()).collect();
    let qb: Vec<u8> = qs.iter().flat_map(|q| q.to_uncompressed()).collect();
    let mut pxy = vec![sys::b200_g1_affine { x: FP_ZERO, y: FP_ZERO }; n];
    let mut qxy = vec![sys::b200_g2_affine { x: z2, y: z2 }; n];
    let (mut pinf, mut qinf, mut st) = (vec![0u8; n], vec![0u8; n], vec![0u8; n]);
    unsafe {
        check(sys::b200_g1_deserialize(eng.0, pb.as_ptr(), n, 0, pxy.as_mut_ptr(), pinf.as_mut_ptr(), st.as_mut_ptr()))?;
        check(sys::b200_g2_deserialize(eng.0, qb.as_ptr(), n, 0, qxy.as_mut_ptr(), qinf.as_mut_ptr(), st.as_mut_ptr()))?;
    }
    Ok((pxy, pinf, qxy, qinf))
}

/// The value of `multi_miller_loop` / one Miller loop: an Fp12 before the final exponentiation (src/pairings.rs:26).
#[derive(Clone, Copy, Debug)]
pub struct B200MillerLoopResult(pub sys::b200_fp12);
/// An element of the target group as computed on the GPU (src/pairings.rs:211); limbs identical to the crate's `Gt.0`.
#[derive(Clone, Copy, Debug)]
pub struct B200Gt(pub sys::b200_fp12);

impl PartialEq for B200Gt { fn eq(&self, o: &Self) -> bool { fp12_eq(&self.0, &o.0) } }
impl Eq for B200Gt {}
impl PartialEq for B200MillerLoopResult { fn eq(&self, o: &Self) -> bool { fp12_eq(&self.0, &o.0) } }
impl Eq for B200MillerLoopResult {}
impl Default for B200MillerLoopResult { fn default() -> Self { B200MillerLoopResult(fp12_one()) } } // src/pairings.rs:28-32

fn tower12(op: i32, a: &sys::b200_fp12, b: Option<&sys::b200_fp12>) -> sys::b200_fp12 {
    let eng = Engine::global();
    let mut out = fp12_one();
    let bp = b.map_or(core::ptr::null(), |x| x.c.as_ptr() as *const u64);
    let rc = unsafe { sys::b200_tower_op(eng.0, 12, op, a.c.as_ptr() as *const u64, bp, out.c.as_mut_ptr() as *mut u64, 1) };
    assert_eq!(rc, sys::B200_OK, "b200_tower_op");
    out
}

impl B200Gt {
    pub fn identity() -> Self { B200Gt(fp12_one()) }                                   // src/pairings.rs:228-231
    pub fn is_identity(&self) -> bool { *self == Self::identity() }
    pub fn double(&self) -> Self { B200Gt(tower12(OP_SQUARE, &self.0, None)) }          // :233-236
    /// `Gt::generator()` = e(G1::generator(), G2::generator())  (src/pairings.rs:359-475, checked :827-832)
    pub fn generator() -> Self {
        Engine::global().pairing_batch_raw(&[G1Affine::generator()], &[G2Affine::generator()]).unwrap()[0]
    }
    pub fn to_raw(&self) -> [[u64; 6]; 12] { core::array::from_fn(|i| self.0.c[i].l) }
}
impl Neg for B200Gt { type Output = B200Gt; fn neg(self) -> B200Gt { B200Gt(tower12(OP_CONJUGATE, &self.0, None)) } } // :253-259
impl Add for B200Gt { type Output = B200Gt; fn add(self, r: B200Gt) -> B200Gt { B200Gt(tower12(OP_MUL, &self.0, Some(&r.0))) } } // :270-276
impl Sub for B200Gt { type Output = B200Gt; fn sub(self, r: B200Gt) -> B200Gt { self + (-r) } }
impl AddAssign for B200Gt { fn add_assign(&mut self, r: B200Gt) { *self = *self + r; } }
impl core::iter::Sum for B200Gt { fn sum<I: Iterator<Item = B200Gt>>(it: I) -> B200Gt { it.fold(B200Gt::identity(), |a, b| a + b) } }
impl<'a> Mul<&'a Scalar> for B200Gt {                                                                     // :296-323
    type Output = B200Gt;
    fn mul(self, s: &Scalar) -> B200Gt {
        let sc = sys::b200_scalar { b: s.to_bytes() };
        let mut out = fp12_one();
        let rc = unsafe { sys::b200_gt_mul_batch(Engine::global().0, &self.0, &sc, 1, &mut out) };
        assert_eq!(rc, sys::B200_OK, "b200_gt_mul_batch");
        B200Gt(out)
    }
}
impl Add for B200MillerLoopResult {                                                                        // :179-186
    type Output = B200MillerLoopResult;
    fn add(self, r: B200MillerLoopResult) -> B200MillerLoopResult { B200MillerLoopResult(tower12(OP_MUL, &self.0, Some(&r.0))) }
}
impl B200MillerLoopResult {
    /// src/pairings.rs:48-176 on the GPU
    pub fn final_exponentiation(&self) -> B200Gt {
        let mut out = fp12_one();
        let rc = unsafe { sys::b200_final_exponentiation_batch(Engine::global().0, &self.0, 1, &mut out) };
        assert_eq!(rc, sys::B200_OK, "b200_final_exponentiation_batch");
        B200Gt(out)
    }
}

impl Engine {
    /// The process-wide engine the operator impls above run on (device `B200_DEVICE`, default 0).
    pub fn global() -> &'static Engine {
        static E: std::sync::OnceLock<Engine> = std::sync::OnceLock::new();
        E.get_or_init(|| {
            let dev = std::env::var("B200_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0);
            Engine::new(dev).expect("no usable B200 (the accelerated path has no CPU fallback)")
        })
    }

    /// `ps.iter().zip(qs).map(|(p, q)| pairing(p, q))` (src/pairings.rs:607-653) for the whole batch on the GPU.
    pub fn pairing_batch_raw(&self, ps: &[G1Affine], qs: &[G2Affine]) -> Result<Vec<B200Gt>, Error> {
        let (pxy, pinf, qxy, qinf) = marshal_pairs(self, ps, qs)?;
        let mut out = vec![fp12_one(); ps.len()];
        check(unsafe { sys::b200_pairing_batch(self.0, pxy.as_ptr(), pinf.as_ptr(), qxy.as_ptr(), qinf.as_ptr(), ps.len(), out.as_mut_ptr()) })?;
        Ok(out.into_iter().map(B200Gt).collect())
    }

    /// `multi_miller_loop(&[(&p_i, &G2Prepared::from(q_i))])` (src/pairings.rs:554-603): ONE MillerLoopResult for all terms;
    /// terms with an identity on either side are skipped (:566-569).  The G2 line coefficients are computed on the GPU.
    pub fn multi_miller_loop_raw(&self, ps: &[G1Affine], qs: &[G2Affine]) -> Result<B200MillerLoopResult, Error> {
        let (pxy, pinf, qxy, qinf) = marshal_pairs(self, ps, qs)?;
        let mut out = fp12_one();
        check(unsafe { sys::b200_multi_miller_loop(self.0, pxy.as_ptr(), pinf.as_ptr(), qxy.as_ptr(), qinf.as_ptr(), ps.len(), &mut out) })?;
        Ok(B200MillerLoopResult(out))
    }

    /// Groth16 / BLS batch verification shape: `n_products` independent products of `terms` pairings each, one final
    /// exponentiation per product — `(0..n).map(|i| multi_miller_loop(&terms[i]).final_exponentiation())`.
    pub fn pairing_products_raw(&self, ps: &[G1Affine], qs: &[G2Affine], terms: usize) -> Result<Vec<B200Gt>, Error> {
        assert!(terms > 0 && ps.len() % terms == 0);
        let (pxy, pinf, qxy, qinf) = marshal_pairs(self, ps, qs)?;
        let n = ps.len() / terms;
        let mut out = vec![fp12_one(); n];
        check(unsafe { sys::b200_pairing_product_batch(self.0, pxy.as_ptr(), pinf.as_ptr(), qxy.as_ptr(), qinf.as_ptr(), terms, n, 1, out.as_mut_ptr()) })?;
        Ok(out.into_iter().map(B200Gt).collect())
    }
}

/// Mechanism (A): with the accessor patch the results ARE the crate's types.
#[cfg(feature = "raw-gt")]
mod raw_gt {
    use super::*;
    use bls12_381::{G2Prepared, Gt, MillerLoopResult};
    use pairing::{Engine as PairingEngine, MultiMillerLoop};

    impl From<B200Gt> for Gt { fn from(g: B200Gt) -> Gt { Gt::from_raw_unchecked(g.to_raw()) } }
    impl From<B200MillerLoopResult> for MillerLoopResult {
        fn from(m: B200MillerLoopResult) -> MillerLoopResult { MillerLoopResult::from_raw_unchecked(core::array::from_fn(|i| m.0.c[i].l)) }
    }
    impl Engine {
        pub fn pairing_batch(&self, ps: &[G1Affine], qs: &[G2Affine]) -> Result<Vec<Gt>, Error> {
            Ok(self.pairing_batch_raw(ps, qs)?.into_iter().map(Gt::from).collect())
        }
    }

    /// `G2Prepared` for the accelerated engine: the affine point (the 68 line-coefficient triples are recomputed on the
    /// device — 8 us of GPU time against moving 19 584 B per point over PCIe; `b200_g2_prepare_dev` keeps them resident in
    /// HBM for fixed verifying keys).
    #[derive(Clone, Debug)]
    pub struct B200G2Prepared(pub G2Affine);
    impl From<G2Affine> for B200G2Prepared { fn from(q: G2Affine) -> Self { B200G2Prepared(q) } }

    /// Drop-in for `bls12_381::Bls12` where code is generic over `pairing::{Engine, MultiMillerLoop}` (src/pairings.rs:795-824).
    #[derive(Clone, Debug)]
    pub struct Bls12B200;
    impl PairingEngine for Bls12B200 {
        type Fr = Scalar;
        type G1 = bls12_381::G1Projective;
        type G1Affine = G1Affine;
        type G2 = bls12_381::G2Projective;
        type G2Affine = G2Affine;
        type Gt = Gt;
        fn pairing(p: &G1Affine, q: &G2Affine) -> Gt { super::Engine::global().pairing_batch(&[*p], &[*q]).unwrap()[0] }
    }
    impl MultiMillerLoop for Bls12B200 {
        type G2Prepared = B200G2Prepared;
        type Result = MillerLoopResult;
        fn multi_miller_loop(terms: &[(&G1Affine, &B200G2Prepared)]) -> MillerLoopResult {
            let ps: Vec<G1Affine> = terms.iter().map(|t| *t.0).collect();
            let qs: Vec<G2Affine> = terms.iter().map(|t| (t.1).0).collect();
            super::Engine::global().multi_miller_loop_raw(&ps, &qs).unwrap().into()
        }
    }
    // keeps the name G2Prepared importable next to the crate's own
    #[allow(dead_code)]
    type _Unused = G2Prepared;
}
#[cfg(feature = "raw-gt")]
pub use raw_gt::{B200G2Prepared, Bls12B200};

/// Without the patch `Bls12B200` still offers the two entry points a verifier calls, on the opaque types.
#[cfg(not(feature = "raw-gt"))]
#[derive(Clone, Debug)]
pub struct Bls12B200;
#[cfg(not(feature = "raw-gt"))]
impl Bls12B200 {
    pub fn pairing(p: &G1Affine, q: &G2Affine) -> B200Gt { Engine::global().pairing_batch_raw(&[*p], &[*q]).unwrap()[0] }
    pub fn multi_miller_loop(terms: &[(&G1Affine, &G2Affine)]) -> B200MillerLoopResult {
        let ps: Vec<G1Affine> = terms.iter().map(|t| *t.0).collect();
        let qs: Vec<G2Affine> = terms.iter().map(|t| *t.1).collect();
        Engine::global().multi_miller_loop_raw(&ps, &qs).unwrap()
    }
}
