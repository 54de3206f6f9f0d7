//! `SIPP<Bls12_377, Blake2s>` -- the reference's own instantiation (`sipp/src/lib.rs:228-254`,
//! `sipp/examples/scaling-ipp.rs:10`) -- on the GPU (`ripp377_*` of `include/ripp_b200.h`).
//! Same packing as `sipp.rs`; ark-bls12-377's `Fq` is 6 x u64 limbs (377 bits), `Fr` 4 x u64 (253 bits).
use crate::{check, ctx};
use ark_bls12_377::{Bls12_377, Fq, Fq12, Fq2, Fq6, Fr, G1Affine, G2Affine};
use ark_ec::pairing::PairingOutput;
use ark_ff::{BigInt, Fp};
use ark_inner_products::Error;
use ark_serialize::{CanonicalDeserialize, CanonicalSerialize};
use ark_sipp::Proof;
use ripp_b200_sys as sys;
use std::marker::PhantomData;
use std::os::raw::c_void;

fn put_fq(o: &mut Vec<u64>, a: &Fq) {
    o.extend_from_slice(&(a.0).0);
}
fn put_fq2(o: &mut Vec<u64>, a: &Fq2) {
    put_fq(o, &a.c0);
    put_fq(o, &a.c1);
}
fn get_fq(w: &[u64]) -> Fq {
    let mut l = [0u64; 6];
    l.copy_from_slice(&w[..6]);
    Fp(BigInt(l), PhantomData)
}
fn get_fq2(w: &[u64]) -> Fq2 {
    Fq2::new(get_fq(&w[..6]), get_fq(&w[6..12]))
}
fn get_fq12(w: &[u64]) -> Fq12 {
    let f6 = |w: &[u64]| Fq6::new(get_fq2(&w[..12]), get_fq2(&w[12..24]), get_fq2(&w[24..36]));
    Fq12::new(f6(&w[..36]), f6(&w[36..72]))
}
fn pack_g1(v: &[G1Affine]) -> Vec<u64> {
    let mut o = Vec::with_capacity(12 * v.len());
    for p in v {
        if p.infinity {
            o.extend_from_slice(&[0u64; 12]);
        } else {
            put_fq(&mut o, &p.x);
            put_fq(&mut o, &p.y);
        }
    }
    o
}
fn pack_g2(v: &[G2Affine]) -> Vec<u64> {
    let mut o = Vec::with_capacity(24 * v.len());
    for p in v {
        if p.infinity {
            o.extend_from_slice(&[0u64; 24]);
        } else {
            put_fq2(&mut o, &p.x);
            put_fq2(&mut o, &p.y);
        }
    }
    o
}
fn pack_fr(v: &[Fr]) -> Vec<u64> {
    v.iter().flat_map(|s| (s.0).0).collect()
}
fn pack_gt(f: &Fq12) -> Vec<u64> {
    let mut o = Vec::with_capacity(72);
    for c in [&f.c0, &f.c1] {
        put_fq2(&mut o, &c.c0);
        put_fq2(&mut o, &c.c1);
        put_fq2(&mut o, &c.c2);
    }
    o
}

/// `product_of_pairings_with_coeffs::<Bls12_377>` (`sipp/src/lib.rs:184-217`).
pub fn product_of_pairings_with_coeffs(a: &[G1Affine], b: &[G2Affine], r: &[Fr]) -> Result<PairingOutput<Bls12_377>, Error> {
    assert_eq!(a.len(), b.len());
    assert_eq!(a.len(), r.len());
    let (aw, bw, rw) = (pack_g1(a), pack_g2(b), pack_fr(r));
    let mut out = [0u64; 72];
    let c = ctx();
    let st = unsafe {
        sys::ripp377_sipp_product_with_coeffs(c.raw(), aw.as_ptr() as *const c_void, bw.as_ptr() as *const c_void,
                                              rw.as_ptr() as *const c_void, a.len(), out.as_mut_ptr() as *mut c_void)
    };
    drop(c);
    check(st, a.len(), b.len())?;
    Ok(PairingOutput(get_fq12(&out)))
}

/// `SIPP::<Bls12_377, Blake2s>::prove` (`sipp/src/lib.rs:42-106`).
pub fn prove(a: &[G1Affine], b: &[G2Affine], r: &[Fr], value: PairingOutput<Bls12_377>) -> Result<Proof<Bls12_377>, Error> {
    assert_eq!(a.len(), b.len());
    let (aw, bw, rw, vw) = (pack_g1(a), pack_g2(b), pack_fr(r), pack_gt(&value.0));
    let rounds = a.len().trailing_zeros() as usize;
    let mut buf = vec![0u8; 2 * 576 * rounds.max(1)];
    let mut len = 0usize;
    let c = ctx();
    let st = unsafe {
        sys::ripp377_sipp_prove(c.raw(), aw.as_ptr() as *const c_void, bw.as_ptr() as *const c_void, rw.as_ptr() as *const c_void,
                                a.len(), vw.as_ptr() as *const c_void, buf.as_mut_ptr(), buf.len(), &mut len)
    };
    drop(c);
    check(st, a.len(), b.len())?;
    let mut rd = &buf[..len];
    let mut gt_elems = Vec::with_capacity(rounds);
    for _ in 0..rounds {
        let zl = PairingOutput::<Bls12_377>::deserialize_uncompressed_unchecked(&mut rd)?;
        let zr = PairingOutput::<Bls12_377>::deserialize_uncompressed_unchecked(&mut rd)?;
        gt_elems.push((zl, zr));
    }
    Ok(Proof { gt_elems }) // `gt_elems` is private in the reference (sipp/src/lib.rs:33): the maintainer adds `pub(crate)` or a constructor
}

/// `SIPP::<Bls12_377, Blake2s>::verify` (`sipp/src/lib.rs:109-180`).
pub fn verify(a: &[G1Affine], b: &[G2Affine], r: &[Fr], value: PairingOutput<Bls12_377>, proof: &Proof<Bls12_377>)
              -> Result<bool, Error> {
    let (aw, bw, rw, vw) = (pack_g1(a), pack_g2(b), pack_fr(r), pack_gt(&value.0));
    let mut pb = Vec::new();
    for (zl, zr) in &proof.gt_elems {
        zl.serialize_uncompressed(&mut pb)?;
        zr.serialize_uncompressed(&mut pb)?;
    }
    let mut accept = 0i32;
    let c = ctx();
    let st = unsafe {
        sys::ripp377_sipp_verify(c.raw(), aw.as_ptr() as *const c_void, bw.as_ptr() as *const c_void, rw.as_ptr() as *const c_void,
                                 a.len(), vw.as_ptr() as *const c_void, pb.as_ptr(), pb.len(), &mut accept)
    };
    drop(c);
    check(st, a.len(), b.len())?;
    Ok(accept == 1)
}
