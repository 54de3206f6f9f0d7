//! SIPP on the GPU (`sipp/src/lib.rs`): there is no trait seam in the reference, so the unit of replacement is the
//! function.  Instantiation: E = BLS12-381, D = Blake2s (BASELINE configs[0]).
use crate::{check, ctx, pack};
use ark_bls12_381::{Bls12_381, Fr, G1Affine, G2Affine};
use ark_ec::pairing::PairingOutput;
use ark_inner_products::Error;
use ark_serialize::CanonicalDeserialize;
use ark_sipp::Proof;
use ripp_b200_sys as sys;
use std::os::raw::c_void;

/// `product_of_pairings_with_coeffs` (`sipp/src/lib.rs:184-217`): prod_i e(r_i a_i, b_i).
pub fn product_of_pairings_with_coeffs(a: &[G1Affine], b: &[G2Affine], r: &[Fr]) -> Result<PairingOutput<Bls12_381>, Error> {
    assert_eq!(a.len(), b.len());
    assert_eq!(a.len(), r.len());
    let (aw, bw, rw) = (pack::pack_g1_aff(a), pack::pack_g2_aff(b), pack::pack_fr(r));
    let mut out = [0u64; 72];
    let c = ctx();
    let st = unsafe {
        sys::ripp_sipp_product_with_coeffs(c.raw(), aw.as_ptr() as *const c_void, bw.as_ptr() as *const c_void,
                                           rw.as_ptr() as *const c_void, a.len(), out.as_mut_ptr() as *mut c_void)
    };
    drop(c);
    check(st, a.len(), b.len())?;
    Ok(PairingOutput(pack::get_fq12(&out)))
}

/// `SIPP::<Bls12_381, Blake2s>::prove` (`sipp/src/lib.rs:42-106`).  Returns the log2(n) pairs (z_l, z_r).
pub fn prove(a: &[G1Affine], b: &[G2Affine], r: &[Fr], value: PairingOutput<Bls12_381>) -> Result<Proof<Bls12_381>, Error> {
    assert_eq!(a.len(), b.len());
    let (aw, bw, rw) = (pack::pack_g1_aff(a), pack::pack_g2_aff(b), pack::pack_fr(r));
    let mut vw = Vec::with_capacity(72);
    pack::put_fq12(&mut vw, &value.0);
    let rounds = a.len().trailing_zeros() as usize;
    let mut buf = vec![0u8; 2 * 576 * rounds.max(1)];
    let mut len = 0usize;
    let c = ctx();
    let st = unsafe {
        sys::ripp_sipp_prove(c.raw(), aw.as_ptr() as *const c_void, bw.as_ptr() as *const c_void, rw.as_ptr() as *const c_void,
                             a.len(), vw.as_ptr() as *const c_void, buf.as_mut_ptr(), buf.len(), &mut len)
    };
    drop(c);
    check(st, a.len(), b.len())?;
    let mut rd = &buf[..len];
    let mut gt_elems = Vec::with_capacity(rounds);
    for _ in 0..rounds {
        let zl = PairingOutput::<Bls12_381>::deserialize_uncompressed_unchecked(&mut rd)?;
        let zr = PairingOutput::<Bls12_381>::deserialize_uncompressed_unchecked(&mut rd)?;
        gt_elems.push((zl, zr));
    }
    Ok(Proof { gt_elems }) // `gt_elems` is private in the reference (sipp/src/lib.rs:33): the maintainer adds `pub(crate)` or a constructor
}

/// `SIPP::<Bls12_381, Blake2s>::verify` (`sipp/src/lib.rs:109-180`).
pub fn verify(a: &[G1Affine], b: &[G2Affine], r: &[Fr], value: PairingOutput<Bls12_381>, proof: &Proof<Bls12_381>)
              -> Result<bool, Error> {
    use ark_serialize::CanonicalSerialize;
    let (aw, bw, rw) = (pack::pack_g1_aff(a), pack::pack_g2_aff(b), pack::pack_fr(r));
    let mut vw = Vec::with_capacity(72);
    pack::put_fq12(&mut vw, &value.0);
    let mut pb = Vec::new();
    for (zl, zr) in &proof.gt_elems {
        zl.serialize_uncompressed(&mut pb)?;
        zr.serialize_uncompressed(&mut pb)?;
    }
    let mut accept = 0i32;
    let c = ctx();
    let st = unsafe {
        sys::ripp_sipp_verify(c.raw(), aw.as_ptr() as *const c_void, bw.as_ptr() as *const c_void, rw.as_ptr() as *const c_void,
                              a.len(), vw.as_ptr() as *const c_void, pb.as_ptr(), pb.len(), &mut accept)
    };
    drop(c);
    check(st, a.len(), b.len())?;
    Ok(accept == 1)
}
