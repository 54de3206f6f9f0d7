"""Verifiers on the GPU through the C ABI: they must accept what the
GPU provers emit (whose bytes equal the oracle's, tests/test_gpu_protocols.py), agree with the oracle verifier's
decision, and reject altered statements / proofs -- the shape of the reference's own round-trip tests
(gipa.rs:470-561, tipa/mod.rs:450-579, structured_scalar_message.rs:360-423, sipp/src/lib.rs:233-254,
benches/examples/groth16_aggregation.rs:92-118)."""
import random

import numpy as np
import pytest

from oracle import bls12_381 as E
from oracle import protocols as O
from oracle import synth as OS
from ripp_b200 import _lib, codec as C
from ripp_b200.ip_proofs import GIPA, TIPA, aggregate_proofs, verify_aggregate_proof
from test_gpu_protocols import N, _inputs, _oracle_gipa, _srs

pytestmark = pytest.mark.gpu
rnd = random.Random(11)


def _flip(b, i, bit=0):
    b = bytearray(b)
    b[i] ^= 1 << bit
    return bytes(b)


def _rejects(fn):
    """A tampered input is either decoded and rejected, or does not decode (RIPP_ERR_ARG) -- never accepted."""
    try:
        return not fn()
    except _lib.RippError as e:
        assert e.status == _lib.RIPP_ERR_ARG
        return True


def test_gt_multiexp(ctx):
    """prod g_i^(s_i) for arbitrary Fq12 elements (not only cyclotomic ones) against the oracle."""
    rf2 = lambda: (rnd.randrange(E.P), rnd.randrange(E.P))
    for n in (1, 2, 7, 11):
        gs = [tuple(rf2() for _ in range(6)) for _ in range(n)]
        ss = [rnd.randrange(E.R) for _ in range(n)]
        ss[0] = 0 if n == 7 else ss[0]
        ss[-1] = E.R - 1 if n == 11 else ss[-1]
        acc = None
        for g, s in zip(gs, ss):
            t = E.f12_pow(g, s)
            acc = t if acc is None else E.f12_mul(acc, t)
        d_g = ctx.to_device(np.stack([C.gt_enc(g) for g in gs]))
        d_s = ctx.to_device(C.fr_vec_enc(ss))
        out = ctx.alloc(576)
        ctx.gt_multiexp_dev(d_g, d_s, n, out)
        ctx.sync()
        assert C.gt_dec(out.download(144)) == acc


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4, 5, 6])
def test_gipa_verify(ctx, kind):
    a, b, v, w = _inputs(kind, N)
    IP, LMC, RMC, IPC = _oracle_gipa(kind)
    ssm = w[0] is None
    g = GIPA(kind, ctx)
    if ssm:  # structured scalar message: b = (1, s, s^2, ...) (structured_scalar_message.rs:392-423)
        s = OS.scalar("ssm-b", 0)
        b = O.structured_scalar_power(N, s)
    proof, _, _ = g.prove_with_aux((a, b), (v, None if ssm else w))
    t = IP.inner_product(a, b)
    com_a = LMC.commit(v, a)
    if ssm:
        assert g.verify((v, None), (com_a, t), proof, scalar_b=s)
        assert not g.verify((v, None), (com_a, t), proof, scalar_b=s + 1)
        bad_t = IP.inner_product(a, b[::-1])
        assert not g.verify((v, None), (com_a, bad_t), proof, scalar_b=s)
    else:
        com_b = RMC.commit(w, b)
        assert g.verify((v, w), (com_a, com_b, t), proof)
        # wrong statement: inner product of a permuted vector
        bad_t = IP.inner_product(a, b[::-1])
        assert not g.verify((v, w), (com_a, com_b, bad_t), proof)
        # wrong keys
        assert not g.verify((v[::-1], w), (com_a, com_b, t), proof)
    # altered proof bytes (r_base lives in the last bytes; Fr / GT low bytes stay decodable)
    for pos in (8, len(proof) // 2, len(proof) - 1):
        ck = (v, None) if ssm else (v, w)
        cm = (com_a, t) if ssm else (com_a, com_b, t)
        assert _rejects(lambda: g.verify(ck, cm, _flip(proof, pos), scalar_b=s if ssm else None))


def test_gipa_verify_errors(ctx):
    a, b, v, w = _inputs(_lib.GIPA_PAIRING, N)
    g = GIPA(_lib.GIPA_PAIRING, ctx)
    proof, _, _ = g.prove_with_aux((a, b), (v, w))
    d = ctx.to_device(C.g2_vec_enc(v))
    with pytest.raises(_lib.RippError) as e:  # gipa.rs:140-146
        ctx.gipa_verify_dev(_lib.GIPA_PAIRING, d, d, 6, b"", proof)
    assert e.value.status == _lib.RIPP_ERR_NOT_POW2
    with pytest.raises(_lib.RippError) as e:  # truncated proof does not decode
        ctx.gipa_verify_dev(_lib.GIPA_PAIRING, d, d, N, b"", proof[:-1])
    assert e.value.status == _lib.RIPP_ERR_ARG


@pytest.mark.parametrize("kind", [_lib.GIPA_PAIRING, _lib.GIPA_MULTIEXP_PEDERSEN, _lib.GIPA_SCALAR_PEDERSEN_G2_G1])
def test_tipa_verify(ctx, kind):
    """tipa/mod.rs:450-579 including the SRS-shift statement (:528-579)."""
    srs = _srs(N)
    vs = srs.get_verifier_key()
    ck_a, ck_b = srs.get_commitment_keys()
    a, b, _, _ = _inputs(kind, N)
    IP, LMC, RMC, IPC = _oracle_gipa(kind)
    t = TIPA(kind, ctx)
    for r_shift in (1, OS.scalar("shift", 0)):
        ck_a_r = [E.g2_mul(k, pow(r_shift, -i, E.R)) for i, k in enumerate(ck_a)]
        proof = t.prove_with_srs_shift((srs.g_alpha_powers, srs.h_beta_powers), (a, b), (ck_a_r, ck_b), r_shift)
        com = (LMC.commit(ck_a_r, a), RMC.commit(ck_b, b), IP.inner_product(a, b))
        assert t.verify_with_srs_shift(vs, com, proof, r_shift)
        assert not t.verify_with_srs_shift(vs, com, proof, r_shift + 1)  # KZG check of the shifted key fails
        assert not t.verify_with_srs_shift(vs, (com[0], com[1], IP.inner_product(a, b[::-1])), proof, r_shift)
        bad_vs = dict(vs, g_beta=E.g1_mul(vs["g_beta"], 2))
        assert not t.verify_with_srs_shift(bad_vs, com, proof, r_shift)
        assert _rejects(lambda: t.verify_with_srs_shift(vs, com, _flip(proof, len(proof) - 1), r_shift))
        assert _rejects(lambda: t.verify_with_srs_shift(vs, com, _flip(proof, 8), r_shift))


def test_tipa_ssm_verify(ctx):
    """structured_scalar_message.rs:360-390."""
    srs = _srs(N)
    vs = srs.get_verifier_key()
    ck_a, _ = srs.get_commitment_keys()
    a = OS.g1_points("ssm-a", N)
    s = OS.scalar("ssm-b", 0)
    b = O.structured_scalar_power(N, s)
    t = TIPA(_lib.GIPA_MULTIEXP_SSM, ctx)
    proof = t.prove((srs.g_alpha_powers, srs.h_beta_powers), (a, b), (ck_a, None))
    com = (O.AFGHOCommitmentG1.commit(ck_a, a), E.msm(a, b, E.g1_add, E.g1_mul))
    assert t.verify_with_structured_scalar_message(vs, com, s, proof)
    assert not t.verify_with_structured_scalar_message(vs, com, s + 1, proof)
    assert not t.verify_with_structured_scalar_message(vs, (com[0], E.g1_mul(com[1], 2)), s, proof)
    assert _rejects(lambda: t.verify_with_structured_scalar_message(vs, com, s, _flip(proof, 8)))


@pytest.mark.parametrize("n", [2, 8])
def test_verify_aggregate_proof(ctx, n):
    """benches/examples/groth16_aggregation.rs:92-118: aggregate, verify; a wrong public input must be rejected."""
    srs = _srs(n)
    vs = srs.get_verifier_key()
    vk, proofs, inputs = OS.groth16_instance(n)
    agg = aggregate_proofs((srs.g_alpha_powers, srs.h_beta_powers), proofs, ctx)
    assert verify_aggregate_proof(vs, vk, inputs, agg, ctx)
    bad = [list(x) for x in inputs]
    bad[n - 1][2] = (bad[n - 1][2] + 1) % E.R
    assert not verify_aggregate_proof(vs, vk, bad, agg, ctx)
    bad_vk = dict(vk, delta_g2=E.g2_mul(vk["delta_g2"], 3))
    assert not verify_aggregate_proof(vs, bad_vk, inputs, agg, ctx)
    # com_a's lowest byte (GT, little-endian): changes r, both TIPA statements and the PPE
    assert _rejects(lambda: verify_aggregate_proof(vs, vk, inputs, _flip(agg, 0), ctx))
    # ip_ab (4th GT)
    assert _rejects(lambda: verify_aggregate_proof(vs, vk, inputs, _flip(agg, 3 * 576), ctx))


@pytest.mark.parametrize("n", [2, 8, 32])
def test_sipp_verify(ctx, n):
    """sipp/src/lib.rs:233-254."""
    from ripp_b200.sipp import SIPP, product_of_pairings_with_coeffs

    a, b, r = OS.g1_points("sipp-a", n), OS.g2_points("sipp-b", n), OS.scalars("sipp-r", n)
    z = product_of_pairings_with_coeffs(a, b, r, ctx)
    proof = SIPP.prove(a, b, r, z, ctx)
    assert SIPP.verify(a, b, r, z, proof, ctx)
    assert not SIPP.verify(a, b, r, E.f12_sqr(z), proof, ctx)  # wrong claimed value
    assert not SIPP.verify(a, b, r[::-1], z, proof, ctx)
    assert _rejects(lambda: SIPP.verify(a, b, r, z, _flip(proof, 0), ctx))


def test_full_size_round_trips(ctx):
    """BASELINE.json's full sizes through the size-independent property the domain offers: prove -> verify accepts,
    verify of an altered statement rejects.  TIPP aggregation of 2^12 proofs (configs[3]) and SIPP over 2^10 pairs
    (configs[0]); inputs generated on the GPU from the synthetic scalar streams."""
    from ripp_b200 import synth

    n = 1 << 12
    inst = synth.tipp_instance_dev(ctx, n)
    proof = ctx.tipp_aggregate_dev(inst["srs_g1"], inst["srs_g2"], inst["a"], inst["b"], inst["c"], n)
    assert len(proof) == 62544
    assert ctx.tipp_verify_aggregate(inst["vsrs"], inst["vk"], inst["inputs"], proof)
    bad = inst["inputs"].copy()
    bad[n - 1, 0] = C.fr_enc((C.fr_dec(bad[n - 1, 0]) + 1) % E.R)
    assert not ctx.tipp_verify_aggregate(inst["vsrs"], inst["vk"], bad, proof)

    m = 1 << 10
    a = synth.g1_points_dev(ctx, "sipp-a", m).download((m, 24))
    b = synth.g2_points_dev(ctx, "sipp-b", m).download((m, 48))
    r = synth.scalars_mont("sipp-r", m)
    z = ctx.sipp_product_with_coeffs(a, b, r)
    sp = ctx.sipp_prove(a, b, r, z)
    assert len(sp) == 10 * 1152
    assert ctx.sipp_verify(a, b, r, z, sp)
    r2 = r.copy()
    r2[5] = r[6]
    assert not ctx.sipp_verify(a, b, r2, z, sp)


def test_verifiers_reject_on_curve_points_outside_the_subgroups(ctx):
    """ark-serialize's deserialize_uncompressed runs Valid::check (curve AND prime-order subgroup, order r for GT) on
    every proof element; the C-ABI verifiers take raw bytes, so they must do the same before the points reach the
    endomorphism-based MSM / fold kernels.  Each substitution below is a well-formed encoding of an element that is
    on the curve (in the cyclotomic subgroup for GT) but outside the prime-order subgroup."""
    from oracle.encoding import ser_g1, ser_g2, ser_gt

    n = 4
    srs = _srs(n)
    vs = srs.get_verifier_key()
    vk, proofs, inputs = OS.groth16_instance(n)
    agg = aggregate_proofs((srs.g_alpha_powers, srs.h_beta_powers), proofs, ctx)
    assert verify_aggregate_proof(vs, vk, inputs, agg, ctx)
    off1, off2, offt = OS.g1_point_off_subgroup(), OS.g2_point_off_subgroup(), OS.gt_cyclotomic_off_subgroup()
    assert E.g1_is_on_curve(off1) and E.g2_is_on_curve(off2)

    def splice(b, at, new):
        return b[:at] + new + b[at + len(new):]

    # AggregateProof = com_a | com_b | com_c | ip_ab (4 GT) | agg_c (G1) | TIPA proof ab | TIPA-SSM proof c;
    # the SSM proof ends with final_ck (G2) | final_ck_proof (G2)
    cases = {
        "G1": splice(agg, 4 * 576, ser_g1(off1)),
        "G2": splice(agg, len(agg) - 192, ser_g2(off2)),
        "GT": splice(agg, 576, ser_gt(offt)),
    }
    for name, bad in cases.items():
        with pytest.raises(_lib.RippError) as e:
            verify_aggregate_proof(vs, vk, inputs, bad, ctx)
        assert e.value.status == _lib.RIPP_ERR_ARG and "subgroup" in str(e.value) and name in str(e.value), name

    # SIPP: the proof's GT elements
    from ripp_b200.sipp import SIPP, product_of_pairings_with_coeffs

    a, b, r = OS.g1_points("sipp-a", n), OS.g2_points("sipp-b", n), OS.scalars("sipp-r", n)
    z = product_of_pairings_with_coeffs(a, b, r, ctx)
    sp = SIPP.prove(a, b, r, z, ctx)
    with pytest.raises(_lib.RippError) as e:
        SIPP.verify(a, b, r, z, splice(sp, 576, ser_gt(offt)), ctx)
    assert "subgroup" in str(e.value)
