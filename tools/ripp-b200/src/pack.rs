//! Field-by-field copies between arkworks values and the ABI's packed limb arrays.
//!
//! arkworks structs are not `repr(C)`, so nothing is transmuted: every coordinate is copied limb by limb.
//! Layouts (u64 words): Fr 4; Fq 6; Fq2 = c0 | c1 (12); G1 affine = x | y (12), identity = all zero; G2 affine = x | y (24);
//! G1 Jacobian = x | y | z (18); G2 Jacobian (36); Fq12 = c0.c0.c0 .. c1.c2.c1 in tower order (72).
use ark_bls12_381::{Fq, Fq12, Fq2, Fq6, Fr, G1Affine, G1Projective, G2Affine, G2Projective};
use ark_ff::{BigInt, Fp, Zero};
use std::marker::PhantomData;

#[inline]
pub fn put_fq(out: &mut Vec<u64>, a: &Fq) {
    out.extend_from_slice(&(a.0).0);
}
#[inline]
pub fn put_fq2(out: &mut Vec<u64>, a: &Fq2) {
    put_fq(out, &a.c0);
    put_fq(out, &a.c1);
}
#[inline]
pub fn get_fq(w: &[u64]) -> Fq {
    let mut l = [0u64; 6];
    l.copy_from_slice(&w[..6]);
    Fp(BigInt(l), PhantomData) // already Montgomery form: no conversion
}
#[inline]
pub fn get_fq2(w: &[u64]) -> Fq2 {
    Fq2::new(get_fq(&w[..6]), get_fq(&w[6..12]))
}
pub fn pack_fr(v: &[Fr]) -> Vec<u64> {
    let mut out = Vec::with_capacity(4 * v.len());
    for s in v {
        out.extend_from_slice(&(s.0).0);
    }
    out
}
pub fn get_fr(w: &[u64]) -> Fr {
    let mut l = [0u64; 4];
    l.copy_from_slice(&w[..4]);
    Fp(BigInt(l), PhantomData)
}
pub fn pack_g1_jac(v: &[G1Projective]) -> Vec<u64> {
    let mut out = Vec::with_capacity(18 * v.len());
    for p in v {
        put_fq(&mut out, &p.x);
        put_fq(&mut out, &p.y);
        put_fq(&mut out, &p.z);
    }
    out
}
pub fn pack_g2_jac(v: &[G2Projective]) -> Vec<u64> {
    let mut out = Vec::with_capacity(36 * v.len());
    for p in v {
        put_fq2(&mut out, &p.x);
        put_fq2(&mut out, &p.y);
        put_fq2(&mut out, &p.z);
    }
    out
}
pub fn pack_g1_aff(v: &[G1Affine]) -> Vec<u64> {
    let mut out = Vec::with_capacity(12 * v.len());
    for p in v {
        if p.infinity {
            out.extend_from_slice(&[0u64; 12]);
        } else {
            put_fq(&mut out, &p.x);
            put_fq(&mut out, &p.y);
        }
    }
    out
}
pub fn pack_g2_aff(v: &[G2Affine]) -> Vec<u64> {
    let mut out = Vec::with_capacity(24 * v.len());
    for p in v {
        if p.infinity {
            out.extend_from_slice(&[0u64; 24]);
        } else {
            put_fq2(&mut out, &p.x);
            put_fq2(&mut out, &p.y);
        }
    }
    out
}
pub fn get_g1_jac(w: &[u64]) -> G1Projective {
    G1Projective::new_unchecked(get_fq(&w[..6]), get_fq(&w[6..12]), get_fq(&w[12..18]))
}
pub fn get_g2_jac(w: &[u64]) -> G2Projective {
    G2Projective::new_unchecked(get_fq2(&w[..12]), get_fq2(&w[12..24]), get_fq2(&w[24..36]))
}
pub fn get_g1_aff(w: &[u64]) -> G1Affine {
    if w[..12].iter().all(|&x| x == 0) {
        G1Affine::identity()
    } else {
        G1Affine::new_unchecked(get_fq(&w[..6]), get_fq(&w[6..12]))
    }
}
pub fn get_g2_aff(w: &[u64]) -> G2Affine {
    if w[..24].iter().all(|&x| x == 0) {
        G2Affine::identity()
    } else {
        G2Affine::new_unchecked(get_fq2(&w[..12]), get_fq2(&w[12..24]))
    }
}
pub fn get_fq12(w: &[u64]) -> Fq12 {
    let f6 = |w: &[u64]| Fq6::new(get_fq2(&w[..12]), get_fq2(&w[12..24]), get_fq2(&w[24..36]));
    Fq12::new(f6(&w[..36]), f6(&w[36..72]))
}
pub fn put_fq12(out: &mut Vec<u64>, f: &Fq12) {
    for c in [&f.c0, &f.c1] {
        put_fq2(out, &c.c0);
        put_fq2(out, &c.c1);
        put_fq2(out, &c.c2);
    }
}
/// `true` when every element is the additive identity (used only by debug assertions).
pub fn all_zero_fr(v: &[Fr]) -> bool {
    v.iter().all(|s| s.is_zero())
}
