//! `impl InnerProduct` over the stateless host-pointer entry points (L1 of `include/ripp_b200.h`).
use crate::{check, ctx, pack};
use ark_bls12_381::{Bls12_381, Fr, G1Projective, G2Projective};
use ark_ec::pairing::PairingOutput;
use ark_inner_products::{Error, InnerProduct};
use ripp_b200_sys as sys;

/// Replaces `PairingInnerProduct<Bls12_381>` (`inner_products/src/lib.rs:52-116`): the batched multi-Miller loop, the
/// Fq12 product tree and ONE final exponentiation on the GPU.
#[derive(Copy, Clone)]
pub struct GpuPairingInnerProduct;

impl InnerProduct for GpuPairingInnerProduct {
    type LeftMessage = G1Projective;
    type RightMessage = G2Projective;
    type Output = PairingOutput<Bls12_381>;

    fn inner_product(left: &[G1Projective], right: &[G2Projective]) -> Result<Self::Output, Error> {
        let (a, b) = (pack::pack_g1_jac(left), pack::pack_g2_jac(right));
        let mut out = [0u64; 72];
        let c = ctx();
        let st = unsafe {
            sys::ripp_pairing_ip(c.raw(), a.as_ptr() as *const _, left.len(), b.as_ptr() as *const _, right.len(),
                                 out.as_mut_ptr() as *mut _)
        };
        check(st, left.len(), right.len())?;
        Ok(PairingOutput(pack::get_fq12(&out)))
    }
}

/// Replaces `MultiexponentiationInnerProduct<G1Projective>` (`inner_products/src/lib.rs:118-142`).
#[derive(Copy, Clone)]
pub struct GpuMultiexponentiationInnerProductG1;

impl InnerProduct for GpuMultiexponentiationInnerProductG1 {
    type LeftMessage = G1Projective;
    type RightMessage = Fr;
    type Output = G1Projective;

    fn inner_product(left: &[G1Projective], right: &[Fr]) -> Result<Self::Output, Error> {
        let (a, s) = (pack::pack_g1_jac(left), pack::pack_fr(right));
        let mut out = [0u64; 18];
        let c = ctx();
        let st = unsafe {
            sys::ripp_msm_g1(c.raw(), a.as_ptr() as *const _, left.len(), s.as_ptr() as *const _, right.len(),
                             out.as_mut_ptr() as *mut _)
        };
        check(st, left.len(), right.len())?;
        Ok(pack::get_g1_jac(&out))
    }
}

/// Replaces `MultiexponentiationInnerProduct<G2Projective>`.
#[derive(Copy, Clone)]
pub struct GpuMultiexponentiationInnerProductG2;

impl InnerProduct for GpuMultiexponentiationInnerProductG2 {
    type LeftMessage = G2Projective;
    type RightMessage = Fr;
    type Output = G2Projective;

    fn inner_product(left: &[G2Projective], right: &[Fr]) -> Result<Self::Output, Error> {
        let (a, s) = (pack::pack_g2_jac(left), pack::pack_fr(right));
        let mut out = [0u64; 36];
        let c = ctx();
        let st = unsafe {
            sys::ripp_msm_g2(c.raw(), a.as_ptr() as *const _, left.len(), s.as_ptr() as *const _, right.len(),
                             out.as_mut_ptr() as *mut _)
        };
        check(st, left.len(), right.len())?;
        Ok(pack::get_g2_jac(&out))
    }
}

/// Replaces `ScalarInnerProduct<Fr>` (`inner_products/src/lib.rs:144-166`).
#[derive(Copy, Clone)]
pub struct GpuScalarInnerProduct;

impl InnerProduct for GpuScalarInnerProduct {
    type LeftMessage = Fr;
    type RightMessage = Fr;
    type Output = Fr;

    fn inner_product(left: &[Fr], right: &[Fr]) -> Result<Self::Output, Error> {
        let (a, b) = (pack::pack_fr(left), pack::pack_fr(right));
        let mut out = [0u64; 4];
        let c = ctx();
        let st = unsafe {
            sys::ripp_scalar_ip(c.raw(), a.as_ptr() as *const _, left.len(), b.as_ptr() as *const _, right.len(),
                                out.as_mut_ptr() as *mut _)
        };
        check(st, left.len(), right.len())?;
        Ok(pack::get_fr(&out))
    }
}
