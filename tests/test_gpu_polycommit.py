"""Polynomial-commitment application on the GPU (ripp_b200/poly_commit.py) against the oracle's restatement at the
reference's bivariate test size (poly_commit/mod.rs:395-433: x_degree = y_degree = 7), and a prove -> verify round trip
at the reference's univariate test size (mod.rs:388,435-471: degree 65535) entirely through the C ABI."""
import random

import pytest

from oracle import bls12_381 as E
from oracle import poly_commit as OPC
from oracle import synth as OS
from ripp_b200 import poly_commit as PC

pytestmark = pytest.mark.gpu
rnd = random.Random(77)
ALPHA, BETA = OS.scalar("pc-alpha", 0), OS.scalar("pc-beta", 0)


def test_kzg_matches_oracle(ctx):
    powers, v_srs = PC.KZG.setup(9, ALPHA, BETA, ctx)
    o_powers, o_v = OPC.KZG.setup(9, ALPHA, BETA)
    assert v_srs == o_v
    poly = [rnd.randrange(E.R) for _ in range(8)]  # shorter than the powers: padded (mod.rs:86-87)
    z = rnd.randrange(E.R)
    com, proof = PC.KZG.commit(powers, poly, ctx), PC.KZG.open(powers, poly, z, ctx)
    assert com == OPC.KZG.commit(o_powers, poly) and proof == OPC.KZG.open(o_powers, poly, z)
    ev = OPC.poly_eval(poly, z)
    assert PC.KZG.verify(v_srs, com, z, ev, proof, ctx)
    assert not PC.KZG.verify(v_srs, com, z, (ev + 1) % E.R, proof, ctx)
    assert not PC.KZG.verify(v_srs, com, (z + 1) % E.R, ev, proof, ctx)


def test_bivariate_matches_oracle(ctx):
    xd = yd = 7
    srs = PC.BivariatePolynomialCommitment.setup(xd, yd, ALPHA, BETA, ctx)
    o_srs = OPC.BivariatePolynomialCommitment.setup(xd, yd, ALPHA, BETA)
    ys = [[rnd.randrange(E.R) for _ in range(yd + 1)] for _ in range(xd)]  # one Y polynomial short: zero padded (mod.rs:190-193)
    com, y_coms = PC.BivariatePolynomialCommitment.commit(srs, ys, ctx)
    o_com, o_y_coms = OPC.BivariatePolynomialCommitment.commit(o_srs, ys)
    assert com == o_com and y_coms == o_y_coms
    point = (rnd.randrange(E.R), rnd.randrange(E.R))
    ev = OPC.bivariate_evaluate(ys, point)
    proof = PC.BivariatePolynomialCommitment.open(srs, ys, y_coms, point, ctx)
    o_proof = OPC.BivariatePolynomialCommitment.open(o_srs, ys, o_y_coms, point)
    assert proof == OPC.ser_opening_proof(o_proof)
    v_srs = srs["v_srs"]
    assert v_srs == o_srs[0].get_verifier_key()
    assert OPC.BivariatePolynomialCommitment.verify(v_srs, o_com, point, ev, o_proof)
    assert PC.BivariatePolynomialCommitment.verify(v_srs, com, point, ev, proof, ctx)
    assert not PC.BivariatePolynomialCommitment.verify(v_srs, com, point, (ev + 1) % E.R, proof, ctx)
    assert not PC.BivariatePolynomialCommitment.verify(v_srs, com, (point[1], point[0]), ev, proof, ctx)


def test_univariate_round_trip_at_reference_size(ctx):
    degree = 65535
    U = PC.UnivariatePolynomialCommitment
    assert U.bivariate_degrees(degree) == (15, 4095)
    srs = U.setup(degree, ALPHA, BETA, ctx)
    poly = [rnd.randrange(E.R) for _ in range(degree + 1)]
    com, y_coms = U.commit(srs, poly, ctx)
    z = rnd.randrange(E.R)
    ev = OPC.poly_eval(poly, z)
    proof = U.open(srs, poly, y_coms, z, ctx)
    assert U.verify(srs["v_srs"], degree, com, z, ev, proof, ctx)
    assert not U.verify(srs["v_srs"], degree, com, z, (ev + 1) % E.R, proof, ctx)
    assert not U.verify(srs["v_srs"], degree, com, (z + 1) % E.R, ev, proof, ctx)


def test_transparent_bivariate_matches_oracle(ctx):
    """transparent.rs:338-393 at x_degree = y_degree = 7: the two GIPAs with structured scalar message."""
    xd = yd = 7
    first, second = OS.g1_points("pct-ck1", yd + 1), OS.g2_points("pct-ck2", xd + 1)
    T, OT = PC.TransparentBivariatePolynomialCommitment, OPC.TransparentBivariatePolynomialCommitment
    ck = T.setup_from_points(first, second, ctx)
    ys = [[rnd.randrange(E.R) for _ in range(yd + 1)] for _ in range(xd + 1)]
    com, y_coms = T.commit(ck, ys, ctx)
    o_com, o_y_coms = OT.commit((first, second), ys)
    assert com == o_com and y_coms == o_y_coms
    point = (rnd.randrange(E.R), rnd.randrange(E.R))
    ev = OPC.bivariate_evaluate(ys, point)
    proof = T.open(ck, ys, y_coms, point, ctx)
    o_proof = OT.open((first, second), ys, o_y_coms, point)
    assert proof == OPC.ser_transparent_opening_proof(o_proof)
    assert OT.verify((first, second), o_com, point, ev, o_proof)
    assert T.verify(ck, com, point, ev, proof, ctx)
    assert not T.verify(ck, com, point, (ev + 1) % E.R, proof, ctx)
    assert not T.verify(ck, com, (point[1], point[0]), ev, proof, ctx)


def test_transparent_univariate_round_trip(ctx):
    """transparent.rs:395-430 shape at degree 4095 -> (x_degree, y_degree) = (15, 255); keys generated on the GPU."""
    from ripp_b200 import synth

    degree = 4095
    U = PC.TransparentUnivariatePolynomialCommitment
    xd, yd = U.bivariate_degrees(degree)
    assert (xd, yd) == (15, 255)
    ck = {"first": (synth.g1_points_dev(ctx, "pct-ck1", yd + 1), yd + 1), "second": (synth.g2_points_dev(ctx, "pct-ck2", xd + 1), xd + 1)}
    poly = [rnd.randrange(E.R) for _ in range(degree + 1)]
    com, y_coms = U.commit(ck, poly, ctx)
    z = rnd.randrange(E.R)
    ev = OPC.poly_eval(poly, z)
    proof = U.open(ck, poly, y_coms, z, ctx)
    assert U.verify(ck, com, z, ev, proof, ctx)
    assert not U.verify(ck, com, z, (ev + 1) % E.R, proof, ctx)
