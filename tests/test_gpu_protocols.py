"""GIPA / TIPA / TIPA-SSM / aggregate_proofs on the GPU: proof bytes must equal the oracle's
serialisation of its own proof on the same seeded inputs, and the oracle's verifier must accept.
Mirrors the reference's n = 8 round-trip tests (gipa.rs:470-561, tipa/mod.rs:450-579,
structured_scalar_message.rs:360-423) and the aggregation example."""
import random

import pytest

from oracle import bls12_381 as E
from oracle import protocols as O
from oracle import synth as OS
from ripp_b200 import _lib, codec as C
from ripp_b200.ip_proofs import GIPA, TIPA, InnerProductArgumentError, aggregate_proofs

pytestmark = pytest.mark.gpu
rnd = random.Random(5)
N = 8


def _frs(n):
    return [rnd.randrange(E.R) for _ in range(n)]


def _oracle_gipa(kind):
    G1, G2, GT, Fr = O.G1T, O.G2T, O.GTT, O.FrT
    MSM1 = O.MultiexponentiationInnerProduct(G1)
    table = {
        _lib.GIPA_PAIRING: (O.PairingInnerProduct, O.AFGHOCommitmentG1, O.AFGHOCommitmentG2, O.IdentityCommitment(GT)),
        _lib.GIPA_MULTIEXP_PEDERSEN: (MSM1, O.AFGHOCommitmentG1, O.PedersenCommitment(G1), O.IdentityCommitment(G1)),
        _lib.GIPA_MULTIEXP_SSM: (MSM1, O.AFGHOCommitmentG1, O.SSMPlaceholderCommitment, O.IdentityCommitment(G1)),
        _lib.GIPA_SCALAR_PEDERSEN_G2_G2: (O.ScalarInnerProduct, O.PedersenCommitment(G2), O.PedersenCommitment(G2), O.IdentityCommitment(Fr)),
        _lib.GIPA_SCALAR_PEDERSEN_G2_G1: (O.ScalarInnerProduct, O.PedersenCommitment(G2), O.PedersenCommitment(G1), O.IdentityCommitment(Fr)),
        _lib.GIPA_SCALAR_SSM: (O.ScalarInnerProduct, O.PedersenCommitment(G2), O.SSMPlaceholderCommitment, O.IdentityCommitment(Fr)),
        _lib.GIPA_SCALAR_SSM_G1: (O.ScalarInnerProduct, O.PedersenCommitment(G1), O.SSMPlaceholderCommitment, O.IdentityCommitment(Fr)),
    }
    return table[kind]


def _inputs(kind, n, seed=0):
    ta, tb, tv, tw = {
        _lib.GIPA_PAIRING: ("G1", "G2", "G2", "G1"),
        _lib.GIPA_MULTIEXP_PEDERSEN: ("G1", "Fr", "G2", "G1"),
        _lib.GIPA_MULTIEXP_SSM: ("G1", "Fr", "G2", None),
        _lib.GIPA_SCALAR_PEDERSEN_G2_G2: ("Fr", "Fr", "G2", "G2"),
        _lib.GIPA_SCALAR_PEDERSEN_G2_G1: ("Fr", "Fr", "G2", "G1"),
        _lib.GIPA_SCALAR_SSM: ("Fr", "Fr", "G2", None),
        _lib.GIPA_SCALAR_SSM_G1: ("Fr", "Fr", "G1", None),
    }[kind]

    def gen(t, tag):
        if t == "G1":
            return OS.g1_points(tag, n, seed)
        if t == "G2":
            return OS.g2_points(tag, n, seed)
        if t == "Fr":
            return OS.scalars(tag, n, seed)
        return [None] * n

    return gen(ta, "gipa-a"), gen(tb, "gipa-b"), gen(tv, "gipa-v"), gen(tw, "gipa-w")


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4, 5, 6])
def test_gipa_proof_bytes_match_oracle(ctx, kind):
    a, b, v, w = _inputs(kind, N)
    IP, LMC, RMC, IPC = _oracle_gipa(kind)
    og = O.GIPA(IP, LMC, RMC, IPC)
    want, aux = og.prove_with_aux((a, b), (v, w, [None]))
    proof, transcript, ck_base = GIPA(kind, ctx).prove_with_aux((a, b), (v, w if w[0] is not None else None))
    assert proof == og.ser_proof(want)
    assert transcript == aux["r_transcript"]
    want_ck = LMC.Key.ser(aux["ck_base"][0]) + (RMC.Key.ser(aux["ck_base"][1]) if w[0] is not None else b"")
    assert ck_base == want_ck
    # and the oracle verifier accepts the (identical) proof: gipa.rs:135-160
    com = (LMC.commit(v, a), RMC.commit(w, b), IPC.commit([None], [IP.inner_product(a, b)]))
    assert og.verify((v, w, None), com, want)


def test_gipa_requires_power_of_two(ctx):
    a, b, v, w = _inputs(_lib.GIPA_PAIRING, 6)
    with pytest.raises(InnerProductArgumentError):
        GIPA(_lib.GIPA_PAIRING, ctx).prove_with_aux((a, b), (v, w))
    d = ctx.to_device(C.g1_vec_enc(a))
    with pytest.raises(_lib.RippError) as e:
        ctx.gipa_prove_dev(_lib.GIPA_PAIRING, d, d, d, d, 6)
    assert e.value.status == _lib.RIPP_ERR_NOT_POW2


def _srs(n, seed=0):
    alpha, beta = OS.scalar("srs-alpha", 0, seed), OS.scalar("srs-beta", 0, seed)
    return O.tipa_setup(n, alpha, beta)


@pytest.mark.parametrize("kind", [_lib.GIPA_PAIRING, _lib.GIPA_MULTIEXP_PEDERSEN, _lib.GIPA_SCALAR_PEDERSEN_G2_G1])
def test_tipa_proof_bytes_match_oracle(ctx, kind):
    """tipa/mod.rs:450-526 plus the SRS-shift variant :528-579."""
    srs = _srs(N)
    ck_a, ck_b = srs.get_commitment_keys()
    a, b, _, _ = _inputs(kind, N)
    IP, LMC, RMC, IPC = _oracle_gipa(kind)
    ot = O.TIPA(IP, LMC, RMC, IPC)
    for r_shift in (1, OS.scalar("shift", 0)):
        ck_a_r = [E.g2_mul(k, pow(r_shift, -i, E.R)) for i, k in enumerate(ck_a)]
        want = ot.prove_with_srs_shift(srs, (a, b), (ck_a_r, ck_b, None), r_shift)
        got = TIPA(kind, ctx).prove_with_srs_shift((srs.g_alpha_powers, srs.h_beta_powers), (a, b), (ck_a_r, ck_b), r_shift)
        assert got == ot.ser_proof(want)
        if kind == _lib.GIPA_PAIRING:  # oracle verifier on the SRS-shift statement (tipa/mod.rs:561-578)
            a_r = a  # commitment to `a` under the shifted key
            com = (LMC.commit(ck_a_r, a_r), RMC.commit(ck_b, b), [IP.inner_product(a_r, b)])
            assert ot.verify_with_srs_shift(srs.get_verifier_key(), None, com, want, r_shift)


def test_tipa_ssm_proof_bytes_match_oracle(ctx):
    """structured_scalar_message.rs:360-390."""
    srs = _srs(N)
    ck_a, _ = srs.get_commitment_keys()
    a = OS.g1_points("ssm-a", N)
    s = OS.scalar("ssm-b", 0)
    b = O.structured_scalar_power(N, s)
    ot = O.TIPAWithSSM(O.MultiexponentiationInnerProduct(O.G1T), O.AFGHOCommitmentG1, O.IdentityCommitment(O.G1T))
    want = ot.prove_with_structured_scalar_message(srs, (a, b), (ck_a, None))
    got = TIPA(_lib.GIPA_MULTIEXP_SSM, ctx).prove((srs.g_alpha_powers, srs.h_beta_powers), (a, b), (ck_a, None))
    assert got == ot.ser_proof(want)
    com = (O.AFGHOCommitmentG1.commit(ck_a, a), [E.msm(a, b, E.g1_add, E.g1_mul)])
    assert ot.verify_with_structured_scalar_message(srs.get_verifier_key(), None, com, s, want)


def test_kzg_opening(ctx):
    srs = _srs(N)
    tr = _frs(3)
    z, shift = _frs(2)
    import numpy as np

    tr_enc = C.fr_vec_enc(tr)
    s1, s2 = ctx.to_device(C.g1_vec_enc(srs.g_alpha_powers)), ctx.to_device(C.g2_vec_enc(srs.h_beta_powers))
    got1 = ctx.kzg_open_dev(1, s1, 2 * N - 1, tr_enc, C.fr_enc(shift).copy(), C.fr_enc(z).copy())
    got2 = ctx.kzg_open_dev(2, s2, 2 * N - 1, tr_enc, C.fr_enc(shift).copy(), C.fr_enc(z).copy())
    assert C.g1_dec(got1) == O.prove_commitment_key_kzg_opening(O.G1T, srs.g_alpha_powers, tr, shift, z)
    assert C.g2_dec(got2) == O.prove_commitment_key_kzg_opening(O.G2T, srs.h_beta_powers, tr, shift, z)


@pytest.mark.parametrize("n", [2, 8])
def test_aggregate_proofs_bytes_match_oracle(ctx, n):
    """benches/examples/groth16_aggregation.rs:92-105 on trapdoor-simulated Groth16 proofs."""
    srs = _srs(n)
    vk, proofs, inputs = OS.groth16_instance(n)
    want = O.aggregate_proofs(srs, proofs)
    got = aggregate_proofs((srs.g_alpha_powers, srs.h_beta_powers), proofs, ctx)
    assert got == O.ser_aggregate_proof(want)
    assert O.verify_aggregate_proof(srs.get_verifier_key(), vk, inputs, want)


@pytest.mark.parametrize("n", [2, 8, 32])
def test_sipp_prove_matches_oracle(ctx, n):
    """sipp/src/lib.rs:233-254 (n = 32 there), on BLS12-381 + Blake2s (BASELINE configs[0])."""
    from ripp_b200.sipp import SIPP, product_of_pairings_with_coeffs

    a, b, r = OS.g1_points("sipp-a", n), OS.g2_points("sipp-b", n), OS.scalars("sipp-r", n)
    z = product_of_pairings_with_coeffs(a, b, r, ctx)
    assert z == O.product_of_pairings_with_coeffs(a, b, r)
    got = SIPP.prove(a, b, r, z, ctx)
    want = O.sipp_prove(a, b, r, z)
    assert got == O.ser_sipp_proof(want)
    assert O.sipp_verify(a, b, r, z, want)
