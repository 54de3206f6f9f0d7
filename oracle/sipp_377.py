"""ORACLE (test infrastructure, NOT product code) -- SIPP over BLS12-377 + Blake2s, the reference's own instantiation
(sipp/src/lib.rs:228-254; sipp/examples/scaling-ipp.rs:10).  Same protocol code as oracle/protocols.py's SIPP section
(sipp/src/lib.rs:42-217), on oracle/bls12_377.py and ark-ec's default point serialisation.  PARITY UNPINNED."""
from . import bls12_377 as E
from .encoding import FiatShamirRng, blake2s, ser_vec


def product_of_pairings_with_coeffs(a, b, r):
    """sipp/src/lib.rs:184-217."""
    return E.multi_pairing([E.g1_mul(p, s) for p, s in zip(a, r)], b)


def _rng(a, b, r, value):
    """lib.rs:56-60: the tuple (a, b, r, value) serialised uncompressed seeds the Fiat-Shamir RNG."""
    return FiatShamirRng(ser_vec(a, E.ser_g1) + ser_vec(b, E.ser_g2) + ser_vec(r, E.ser_fr) + E.ser_gt(value), blake2s)


def sipp_prove(a, b, r, value):
    """lib.rs:42-106 -> [(z_l, z_r)]."""
    assert len(a) == len(b) and bin(len(a)).count("1") == 1
    rng = _rng(a, b, r, value)
    a = [E.g1_mul(p, s) for p, s in zip(a, r)]
    b = list(b)
    length = len(a)
    proof = []
    while length != 1:
        length //= 2
        a_l, a_r, b_l, b_r = a[:length], a[length:], b[:length], b[length:]
        z_l, z_r = E.multi_pairing(a_r, b_l), E.multi_pairing(a_l, b_r)
        proof.append((z_l, z_r))
        rng.absorb(E.ser_gt(z_l) + E.ser_gt(z_r))
        x = rng.next_u128()
        xi = E.fr_inv(x)
        a = [E.g1_add(E.g1_mul(hi, x), lo) for hi, lo in zip(a_r, a_l)]
        b = [E.g2_add(E.g2_mul(hi, xi), lo) for hi, lo in zip(b_r, b_l)]
    return proof


def sipp_verify(a, b, r, claimed_value, proof):
    """lib.rs:109-180."""
    length = len(a)
    assert bin(length).count("1") == 1 and length >= 2 and length == len(b) and 1 << len(proof) == length
    k = len(proof)
    rng = _rng(a, b, r, claimed_value)
    xs = []
    for z_l, z_r in proof:
        rng.absorb(E.ser_gt(z_l) + E.ser_gt(z_r))
        xs.append(rng.next_u128())
    xinv = [E.fr_inv(x) for x in xs]
    z = claimed_value
    for (z_l, z_r), x, xi in zip(proof, xs, xinv):
        z = E.gt_mul(z, E.gt_mul(E.gt_pow(z_l, x), E.gt_pow(z_r, xi)))
    s, si = [1] * length, [1] * length
    for j, (x, xi) in enumerate(zip(xs, xinv)):
        for i in range(length):
            if i & (1 << (k - j - 1)):
                s[i] = s[i] * x % E.R
                si[i] = si[i] * xi % E.R
    s = [x * ri % E.R for x, ri in zip(s, r)]
    return E.pairing(E.msm(a, s, E.g1_add, E.g1_mul), E.msm(b, si, E.g2_add, E.g2_mul)) == z


def ser_proof(proof):
    return b"".join(E.ser_gt(zl) + E.ser_gt(zr) for zl, zr in proof)


def points(tag, n, group):
    """Deterministic subgroup points s * generator, s from the synthetic-scalar hash of oracle/synth.py (reduced mod r)."""
    from .synth import scalar

    g, mul = (E.G1_GEN, E.g1_mul) if group == 1 else (E.G2_GEN, E.g2_mul)
    return [mul(g, scalar(tag, i) % E.R) for i in range(n)]
