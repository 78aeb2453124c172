//! Safe wrapper: the accelerated hot path of `bls12_381` (v0.8) on a B200 through `libbls12381_b200.so`.
//!
//! The reference crate denies `unsafe` and keeps `Fp`/`Fp2`/`Fp12` private, so this wrapper marshals ONLY through
//! the crate's public, canonical byte encodings — no patch to `bls12_381` is needed:
//!   * points go in as `G1Affine::to_uncompressed()` / `G2Affine::to_uncompressed()` bytes and are decoded on the
//!     GPU (`b200_g{1,2}_deserialize`, the device-side `from_uncompressed_unchecked`, src/g1.rs:275-320);
//!   * scalars go in as `Scalar::to_bytes()` (src/scalar.rs:284);
//!   * results come back as uncompressed bytes (`b200_g{1,2}_batch_normalize` + `b200_g{1,2}_serialize`) and are
//!     rebuilt with `from_uncompressed_unchecked` (the GPU result is on the curve and in the subgroup by
//!     construction).
//! NOT compiled in the build container of this repository (no Rust toolchain there); the same call sequence is
//! exercised from Python in tests/test_gpu_parity.py::test_msm_through_byte_encodings.
use bls12_381::{G1Affine, G1Projective, G2Affine, G2Projective, Scalar};
use bls12381_b200_sys as sys;
use core::ffi::c_int;

pub mod pairings;
pub use pairings::{B200Gt, B200MillerLoopResult, Bls12B200};

#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub struct Error(pub c_int);

pub(crate) fn check(rc: c_int) -> Result<(), Error> {
    if rc == sys::B200_OK { Ok(()) } else { Err(Error(rc)) }
}

/// One engine = one GPU (`b200_ctx`): stream + scratch memory.  Calls are serialised inside the library.
pub struct Engine(pub(crate) *mut sys::b200_ctx);
unsafe impl Send for Engine {}
unsafe impl Sync for Engine {}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { sys::b200_ctx_destroy(self.0) }
    }
}

impl Engine {
    /// `device` = CUDA ordinal, -1 = current.  Fails with `Error(B200_ENODEV)` without a GPU (no CPU fallback).
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut h = core::ptr::null_mut();
        check(unsafe { sys::b200_ctx_create(device, &mut h) })?;
        Ok(Engine(h))
    }

    /// `bases.iter().zip(scalars).map(|(p, s)| p * s).sum::<G1Projective>()` (src/g1.rs:573-579, :161-171) on the GPU.
    pub fn g1_msm(&self, bases: &[G1Affine], scalars: &[Scalar]) -> Result<G1Projective, Error> {
        assert_eq!(bases.len(), scalars.len());
        let n = bases.len();
        let bytes: Vec<u8> = bases.iter().flat_map(|p| p.to_uncompressed()).collect();
        let s: Vec<sys::b200_scalar> = scalars.iter().map(|s| sys::b200_scalar { b: s.to_bytes() }).collect();
        let mut xy = vec![sys::b200_g1_affine { x: sys::b200_fp { l: [0; 6] }, y: sys::b200_fp { l: [0; 6] } }; n];
        let (mut inf, mut status) = (vec![0u8; n], vec![0u8; n]);
        unsafe {
            check(sys::b200_g1_deserialize(self.0, bytes.as_ptr(), n, 0, xy.as_mut_ptr(), inf.as_mut_ptr(), status.as_mut_ptr()))?;
            debug_assert!(status.iter().all(|&s| s & 1 == 1));
            let mut out = core::mem::MaybeUninit::<sys::b200_g1_projective>::uninit();
            check(sys::b200_g1_msm(self.0, xy.as_ptr(), inf.as_ptr(), s.as_ptr(), n, out.as_mut_ptr()))?;
            let out = out.assume_init();
            let mut axy = xy[0..0].to_vec();
            axy.push(sys::b200_g1_affine { x: sys::b200_fp { l: [0; 6] }, y: sys::b200_fp { l: [0; 6] } });
            let mut ainf = [0u8; 1];
            check(sys::b200_g1_batch_normalize(self.0, &out, 1, axy.as_mut_ptr(), ainf.as_mut_ptr()))?;
            let mut enc = [0u8; 96];
            check(sys::b200_g1_serialize(self.0, axy.as_ptr(), ainf.as_ptr(), 1, 0, enc.as_mut_ptr()))?;
            Ok(G1Projective::from(G1Affine::from_uncompressed_unchecked(&enc).unwrap()))
        }
    }

    /// Same for G2 (src/g2.rs:626-632, :162-172).
    pub fn g2_msm(&self, bases: &[G2Affine], scalars: &[Scalar]) -> Result<G2Projective, Error> {
        assert_eq!(bases.len(), scalars.len());
        let n = bases.len();
        let bytes: Vec<u8> = bases.iter().flat_map(|p| p.to_uncompressed()).collect();
        let s: Vec<sys::b200_scalar> = scalars.iter().map(|s| sys::b200_scalar { b: s.to_bytes() }).collect();
        let z = sys::b200_fp { l: [0; 6] };
        let z2 = sys::b200_fp2 { c0: z, c1: z };
        let mut xy = vec![sys::b200_g2_affine { x: z2, y: z2 }; n];
        let (mut inf, mut status) = (vec![0u8; n], vec![0u8; n]);
        unsafe {
            check(sys::b200_g2_deserialize(self.0, bytes.as_ptr(), n, 0, xy.as_mut_ptr(), inf.as_mut_ptr(), status.as_mut_ptr()))?;
            let mut out = core::mem::MaybeUninit::<sys::b200_g2_projective>::uninit();
            check(sys::b200_g2_msm(self.0, xy.as_ptr(), inf.as_ptr(), s.as_ptr(), n, out.as_mut_ptr()))?;
            let out = out.assume_init();
            let mut axy = [sys::b200_g2_affine { x: z2, y: z2 }];
            let mut ainf = [0u8; 1];
            check(sys::b200_g2_batch_normalize(self.0, &out, 1, axy.as_mut_ptr(), ainf.as_mut_ptr()))?;
            let mut enc = [0u8; 192];
            check(sys::b200_g2_serialize(self.0, axy.as_ptr(), ainf.as_ptr(), 1, 0, enc.as_mut_ptr()))?;
            Ok(G2Projective::from(G2Affine::from_uncompressed_unchecked(&enc).unwrap()))
        }
    }

    /// Batched subgroup / on-curve validation of untrusted encodings: what `G1Affine::from_compressed` checks
    /// (src/g1.rs:330-336), for `n` points at once.  Returns `true` per point that is a valid element of G1.
    pub fn g1_validate_compressed(&self, encodings: &[[u8; 48]]) -> Result<Vec<bool>, Error> {
        let n = encodings.len();
        let flat: Vec<u8> = encodings.iter().flatten().copied().collect();
        let mut xy = vec![sys::b200_g1_affine { x: sys::b200_fp { l: [0; 6] }, y: sys::b200_fp { l: [0; 6] } }; n];
        let (mut inf, mut st, mut chk) = (vec![0u8; n], vec![0u8; n], vec![0u8; n]);
        unsafe {
            check(sys::b200_g1_deserialize(self.0, flat.as_ptr(), n, 1, xy.as_mut_ptr(), inf.as_mut_ptr(), st.as_mut_ptr()))?;
            check(sys::b200_g1_check(self.0, xy.as_ptr(), inf.as_ptr(), n, chk.as_mut_ptr()))?;
        }
        Ok((0..n).map(|i| st[i] & 1 == 1 && chk[i] == 3).collect())
    }

    /// In-place FFT over `Scalar::ROOT_OF_UNITY` (src/scalar.rs:200) of `a.len() = 2^k` scalars — forward, inverse (with
    /// the 1/n scaling) and the coset variants bellman's `EvaluationDomain` uses.  Marshalled through
    /// `Scalar::to_bytes` / `from_bytes` and converted to/from Montgomery limbs on the GPU.
    pub fn fft(&self, a: &mut [Scalar], inverse: bool, coset: bool) -> Result<(), Error> {
        let n = a.len();
        assert!(n.is_power_of_two());
        let k = n.trailing_zeros() as c_int;
        let enc: Vec<sys::b200_scalar> = a.iter().map(|s| sys::b200_scalar { b: s.to_bytes() }).collect();
        let mut limbs = vec![sys::b200_fr { l: [0; 4] }; n];
        let mut ok = vec![0u8; n];
        let mut out = vec![sys::b200_scalar { b: [0; 32] }; n];
        unsafe {
            check(sys::b200_fr_from_bytes(self.0, enc.as_ptr(), n, limbs.as_mut_ptr(), ok.as_mut_ptr()))?;
            check(sys::b200_fr_ntt(self.0, limbs.as_ptr(), k, inverse as c_int, coset as c_int, limbs.as_mut_ptr()))?;
            check(sys::b200_fr_to_bytes(self.0, limbs.as_ptr(), n, out.as_mut_ptr()))?;
        }
        for (x, y) in a.iter_mut().zip(&out) {
            *x = Scalar::from_bytes(&y.b).unwrap();
        }
        Ok(())
    }

    /// `<G1Projective as HashToCurve<ExpandMsgXmd<Sha256>>>::hash_to_curve(msg, dst)` for every message
    /// (src/hash_to_curve/mod.rs:86-92; feature "experimental" of the reference), one GPU thread per message.
    pub fn g1_hash_to_curve(&self, msgs: &[&[u8]], dst: &[u8]) -> Result<Vec<G1Projective>, Error> {
        let n = msgs.len();
        let mut off = Vec::with_capacity(n + 1);
        let mut cat = Vec::new();
        off.push(0u64);
        for m in msgs {
            cat.extend_from_slice(m);
            off.push(cat.len() as u64);
        }
        let z = sys::b200_fp { l: [0; 6] };
        let mut pr = vec![sys::b200_g1_projective { x: z, y: z, z }; n];
        let mut xy = vec![sys::b200_g1_affine { x: z, y: z }; n];
        let mut inf = vec![0u8; n];
        let mut enc = vec![0u8; 96 * n];
        unsafe {
            check(sys::b200_g1_hash_to_curve(self.0, cat.as_ptr(), off.as_ptr(), n, dst.as_ptr(), dst.len(), 0, pr.as_mut_ptr()))?;
            check(sys::b200_g1_batch_normalize(self.0, pr.as_ptr(), n, xy.as_mut_ptr(), inf.as_mut_ptr()))?;
            check(sys::b200_g1_serialize(self.0, xy.as_ptr(), inf.as_ptr(), n, 0, enc.as_mut_ptr()))?;
        }
        Ok(enc.chunks_exact(96)
            .map(|c| G1Projective::from(G1Affine::from_uncompressed_unchecked(c.try_into().unwrap()).unwrap()))
            .collect())
    }
}
