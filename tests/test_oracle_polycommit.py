"""The oracle's restatement of the polynomial-commitment application (oracle/poly_commit.py;
ip_proofs/src/applications/poly_commit/mod.rs:379-472 are the reference's own round-trip tests)."""
import random

from oracle import bls12_381 as E
from oracle import poly_commit as PC
from oracle import synth as OS

rnd = random.Random(31)
ALPHA, BETA = OS.scalar("pc-alpha", 0), OS.scalar("pc-beta", 0)


def test_bivariate_degrees_and_form():
    U = PC.UnivariatePolynomialCommitment
    assert U.bivariate_degrees(65535) == (15, 4095)   # the reference's UNIVARIATE_DEGREE (mod.rs:388)
    assert U.bivariate_degrees(1048575) == (63, 16383)
    assert U.bivariate_degrees(56) == (1, 31)
    assert U.bivariate_degrees(3) == (1, 1)
    coeffs = list(range(1, 8))
    form = U.bivariate_form((1, 3), coeffs)
    assert form == [[1, 2, 3, 4], [5, 6, 7, 0]]
    z = rnd.randrange(E.R)
    assert PC.bivariate_evaluate(form, (pow(z, 4, E.R), z)) == PC.poly_eval(coeffs, z)


def test_kzg_round_trip():
    powers, v_srs = PC.KZG.setup(5, ALPHA, BETA)
    poly = [rnd.randrange(E.R) for _ in range(6)]
    com = PC.KZG.commit(powers, poly)
    z = rnd.randrange(E.R)
    proof = PC.KZG.open(powers, poly, z)
    assert PC.KZG.verify(v_srs, com, z, PC.poly_eval(poly, z), proof)
    assert not PC.KZG.verify(v_srs, com, z, (PC.poly_eval(poly, z) + 1) % E.R, proof)


def test_bivariate_round_trip():
    """mod.rs:395-433 at x_degree = y_degree = 3."""
    srs = PC.BivariatePolynomialCommitment.setup(3, 3, ALPHA, BETA)
    v_srs = srs[0].get_verifier_key()
    ys = [[rnd.randrange(E.R) for _ in range(4)] for _ in range(4)]
    com, y_coms = PC.BivariatePolynomialCommitment.commit(srs, ys)
    point = (rnd.randrange(E.R), rnd.randrange(E.R))
    ev = PC.bivariate_evaluate(ys, point)
    proof = PC.BivariatePolynomialCommitment.open(srs, ys, y_coms, point)
    assert PC.BivariatePolynomialCommitment.verify(v_srs, com, point, ev, proof)
    assert not PC.BivariatePolynomialCommitment.verify(v_srs, com, point, (ev + 1) % E.R, proof)
    assert not PC.BivariatePolynomialCommitment.verify(v_srs, com, (point[1], point[0]), ev, proof)


def test_transparent_round_trip():
    """transparent.rs:338-393 at x_degree = y_degree = 3 (no trusted setup: Pedersen first tier, two GIPAs with SSM)."""
    T = PC.TransparentBivariatePolynomialCommitment
    ck = (OS.g1_points("pct-ck1", 4), OS.g2_points("pct-ck2", 4))
    ys = [[rnd.randrange(E.R) for _ in range(4)] for _ in range(3)]  # one Y polynomial short: zero padded
    com, y_coms = T.commit(ck, ys)
    point = (rnd.randrange(E.R), rnd.randrange(E.R))
    ev = PC.bivariate_evaluate(ys, point)
    proof = T.open(ck, ys, y_coms, point)
    assert T.verify(ck, com, point, ev, proof)
    assert not T.verify(ck, com, point, (ev + 1) % E.R, proof)
    U = PC.TransparentUnivariatePolynomialCommitment
    assert U.bivariate_degrees(65535) == (63, 1023) and U.bivariate_degrees(15) == (1, 7)
    coeffs = [rnd.randrange(E.R) for _ in range(14)]
    ck2 = (OS.g1_points("pct-ck1", 8), OS.g2_points("pct-ck2", 2))
    com, y_coms = U.commit(ck2, coeffs)
    z = rnd.randrange(E.R)
    assert U.verify(ck2, com, z, PC.poly_eval(coeffs, z), U.open(ck2, coeffs, y_coms, z))
