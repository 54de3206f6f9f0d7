//! `impl DoublyHomomorphicCommitment` for AFGHO16 and Pedersen over the GPU inner products.
//!
//! `setup` keeps the reference's own code path (`random_generators`, `dh_commitments/src/lib.rs:59-61`): the draws come
//! from the caller's `Rng`, so seeded tests see exactly the reference's keys.  `setup_from_exponents` is the
//! GPU key generator for callers that draw exponents instead of points (`ripp_fixed_base_msm_g{1,2}_dev`).
use crate::inner_products::{GpuMultiexponentiationInnerProductG1, GpuMultiexponentiationInnerProductG2, GpuPairingInnerProduct};
use crate::resident::DeviceVec;
use crate::{check, ctx};
use ark_bls12_381::{Bls12_381, Fr, G1Affine, G1Projective, G2Affine, G2Projective};
use ark_dh_commitments::{random_generators, DoublyHomomorphicCommitment, Error};
use ark_ec::pairing::PairingOutput;
use ark_inner_products::InnerProduct;
use ark_std::rand::Rng;
use ripp_b200_sys as sys;

/// `AFGHOCommitmentG1<Bls12_381>` (`afgho16/mod.rs:20-33`): message in G1, key in G2, commit = IP(m, k).
#[derive(Clone)]
pub struct GpuAFGHOCommitmentG1;
impl DoublyHomomorphicCommitment for GpuAFGHOCommitmentG1 {
    type Scalar = Fr;
    type Message = G1Projective;
    type Key = G2Projective;
    type Output = PairingOutput<Bls12_381>;
    fn setup<R: Rng>(rng: &mut R, size: usize) -> Result<Vec<Self::Key>, Error> {
        Ok(random_generators(rng, size))
    }
    fn commit(k: &[Self::Key], m: &[Self::Message]) -> Result<Self::Output, Error> {
        Ok(GpuPairingInnerProduct::inner_product(m, k)?)
    }
}

/// `AFGHOCommitmentG2<Bls12_381>` (`afgho16/mod.rs:35-48`): message in G2, key in G1, commit = IP(k, m).
#[derive(Clone)]
pub struct GpuAFGHOCommitmentG2;
impl DoublyHomomorphicCommitment for GpuAFGHOCommitmentG2 {
    type Scalar = Fr;
    type Message = G2Projective;
    type Key = G1Projective;
    type Output = PairingOutput<Bls12_381>;
    fn setup<R: Rng>(rng: &mut R, size: usize) -> Result<Vec<Self::Key>, Error> {
        Ok(random_generators(rng, size))
    }
    fn commit(k: &[Self::Key], m: &[Self::Message]) -> Result<Self::Output, Error> {
        Ok(GpuPairingInnerProduct::inner_product(k, m)?)
    }
}

/// `PedersenCommitment<G1Projective>` (`pedersen/mod.rs:14-27`): commit = MSM(keys, messages).
#[derive(Clone)]
pub struct GpuPedersenCommitmentG1;
impl DoublyHomomorphicCommitment for GpuPedersenCommitmentG1 {
    type Scalar = Fr;
    type Message = Fr;
    type Key = G1Projective;
    type Output = G1Projective;
    fn setup<R: Rng>(rng: &mut R, size: usize) -> Result<Vec<Self::Key>, Error> {
        Ok(random_generators(rng, size))
    }
    fn commit(k: &[Self::Key], m: &[Self::Message]) -> Result<Self::Output, Error> {
        Ok(GpuMultiexponentiationInnerProductG1::inner_product(k, m)?)
    }
}

/// `PedersenCommitment<G2Projective>`.
#[derive(Clone)]
pub struct GpuPedersenCommitmentG2;
impl DoublyHomomorphicCommitment for GpuPedersenCommitmentG2 {
    type Scalar = Fr;
    type Message = Fr;
    type Key = G2Projective;
    type Output = G2Projective;
    fn setup<R: Rng>(rng: &mut R, size: usize) -> Result<Vec<Self::Key>, Error> {
        Ok(random_generators(rng, size))
    }
    fn commit(k: &[Self::Key], m: &[Self::Message]) -> Result<Self::Output, Error> {
        Ok(GpuMultiexponentiationInnerProductG2::inner_product(k, m)?)
    }
}

/// Key vector `[e_i * g1]` on the device from caller-drawn exponents (fixed-base windowed table on the GPU).
pub fn g1_keys_from_exponents(exponents: &[Fr]) -> Result<DeviceVec<G1Affine>, Error> {
    let s = DeviceVec::<Fr>::upload_fr(exponents)?;
    let out = DeviceVec::<G1Affine>::alloc(exponents.len())?;
    let c = ctx();
    let st = unsafe { sys::ripp_fixed_base_msm_g1_dev(c.raw(), std::ptr::null(), s.ptr(), exponents.len(), out.ptr_mut()) };
    check(st, exponents.len(), exponents.len())?;
    Ok(out)
}
/// Key vector `[e_i * g2]` on the device.
pub fn g2_keys_from_exponents(exponents: &[Fr]) -> Result<DeviceVec<G2Affine>, Error> {
    let s = DeviceVec::<Fr>::upload_fr(exponents)?;
    let out = DeviceVec::<G2Affine>::alloc(exponents.len())?;
    let c = ctx();
    let st = unsafe { sys::ripp_fixed_base_msm_g2_dev(c.raw(), std::ptr::null(), s.ptr(), exponents.len(), out.ptr_mut()) };
    check(st, exponents.len(), exponents.len())?;
    Ok(out)
}
