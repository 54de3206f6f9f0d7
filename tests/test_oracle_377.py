"""Pins of the BLS12-377 oracle (oracle/bls12_377.py, oracle/sipp_377.py) by derivation and identities: the curve is
DERIVED from the BLS12 family polynomials, so primality, curve orders, the twist and the pairing's bilinearity are checks,
not recollections.  SIPP on the reference's own instantiation (sipp/src/lib.rs:228-254) round-trips and rejects tampering."""
import random

from oracle import bls12_377 as E
from oracle import sipp_377 as S

rnd = random.Random(377)


def _probable_prime(n):
    return all(pow(a, n - 1, n) == 1 for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37))


def test_parameters_follow_from_x():
    assert E.P.bit_length() == 377 and E.R.bit_length() == 253
    assert _probable_prime(E.P) and _probable_prime(E.R)
    assert (E.P + 1 - (E.X + 1)) % E.R == 0  # #E(Fp) = p + 1 - t, t = x + 1, divisible by r
    assert (E.P**12 - 1) % E.R == 0 and all((E.P**k - 1) % E.R for k in (1, 2, 3, 4, 6))  # embedding degree 12
    assert pow(E.P - E.BETA, (E.P - 1) // 2, E.P) == E.P - 1  # -5 is a non-residue: Fq2 is a field
    assert (E.X**2 - 1) ** 2 % E.R == (-(E.X**2 - 1) - 1) % E.R  # lambda = x^2 - 1 is a cube root of unity mod r


def test_generators_and_twist():
    assert E.g1_is_on_curve(E.G1_GEN) and E._mul_raw(E.G1_GEN, E.R, E.g1_add) is None
    assert E.g2_is_on_curve(E.G2_GEN) and E._mul_raw(E.G2_GEN, E.R, E.g2_add) is None
    assert E.f2_mul(E.B2, E.XI) == E.F2_ONE  # D-type twist: b' = b / xi
    # the derived G1 generator is the one arkworks uses for BLS12-377 (x coordinate recalled from ark-bls12-377)
    assert hex(E.G1_GEN[0]).startswith("0x8848defe740a67c8fc6225bf87ff5485951e2caa9d41bb188282c8bd37cb5cd5481512ffcd394eeab9b16eb21be9ef")


def test_pairing_bilinear_nondegenerate_order_r():
    a, b = rnd.randrange(1, E.R), rnd.randrange(1, E.R)
    e0 = E.pairing(E.G1_GEN, E.G2_GEN)
    assert e0 != E.F12_ONE and E.f12_pow(e0, E.R) == E.F12_ONE
    assert E.pairing(E.g1_mul(E.G1_GEN, a), E.g2_mul(E.G2_GEN, b)) == E.gt_pow(e0, a * b % E.R)
    f = E.miller_loop(E.G1_GEN, E.G2_GEN)
    assert E.final_exponentiation(f) == E.final_exponentiation_naive(f)
    assert E.pairing(None, E.G2_GEN) == E.F12_ONE


def test_serialisation_flags():
    p = E.g1_mul(E.G1_GEN, 5)
    s, sn = E.ser_g1(p), E.ser_g1(E.g1_neg(p))
    assert len(s) == 96 and len(sn) == 96
    assert (s[95] & 0x80) != (sn[95] & 0x80) and s[:48] == sn[:48]  # the sign flag separates y from -y
    assert E.ser_g1(None)[95] == 0x40 and not any(E.ser_g1(None)[:95])
    q = E.g2_mul(E.G2_GEN, 7)
    t, tn = E.ser_g2(q), E.ser_g2(E.g2_neg(q))
    assert len(t) == 192 and (t[191] & 0x80) != (tn[191] & 0x80)
    assert len(E.ser_gt(E.pairing(p, q))) == 576


def test_sipp_round_trip_and_tamper():
    """sipp/src/lib.rs:233-254 at n = 8 (the reference's test uses 32 with the same structure)."""
    n = 8
    a, b = S.points("s377-a", n, 1), S.points("s377-b", n, 2)
    r = [rnd.randrange(E.R) for _ in range(n)]
    z = S.product_of_pairings_with_coeffs(a, b, r)
    proof = S.sipp_prove(a, b, r, z)
    assert len(proof) == 3 and S.sipp_verify(a, b, r, z, proof)
    assert not S.sipp_verify(a, b, r, E.gt_mul(z, z), proof)
    bad = list(proof)
    bad[1] = (bad[1][1], bad[1][0])
    assert not S.sipp_verify(a, b, r, z, bad)
