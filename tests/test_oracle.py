"""The oracle itself (oracle/*.py), pinned by algebraic identities and by the behaviours the
reference's own tests assert: honest proofs verify, tampered ones do not,
commitments reject wrong messages, the SRS-shift identity holds.  PARITY UNPINNED: the reference has
no golden vectors and cannot be built here; these tests are what anchors the oracle."""
import hashlib
import random

import pytest

from oracle import bls12_381 as E
from oracle import encoding as enc
from oracle import protocols as O
from oracle import synth as OS

rnd = random.Random(1)
N = 4


def test_curve_constants():
    assert E.g1_is_on_curve(E.G1_GEN) and E.g2_is_on_curve(E.G2_GEN)
    assert E.g1_mul(E.G1_GEN, E.R - 1) == E.g1_neg(E.G1_GEN)
    assert E.g2_mul(E.G2_GEN, E.R - 1) == E.g2_neg(E.G2_GEN)
    x = E.X_ABS
    assert E.R == x**4 - x**2 + 1 and E.P == (x + 1) ** 2 * E.R // 3 - x  # x negative: (x-1)^2 -> (|x|+1)^2


def test_pairing_bilinearity_and_order():
    e = E.pairing(E.G1_GEN, E.G2_GEN)
    a, b = rnd.randrange(E.R), rnd.randrange(E.R)
    assert E.pairing(E.g1_mul(E.G1_GEN, a), E.g2_mul(E.G2_GEN, b)) == E.gt_pow(e, a * b)
    assert E.f12_pow(e, E.R) == E.F12_ONE and e != E.F12_ONE
    assert E.pairing(None, E.G2_GEN) == E.F12_ONE


def test_final_exponentiation_is_cube_of_reduced_ate():
    """ark-ec's hard part computes exponent 3 (p^4 - p^2 + 1)/r."""
    f = E.miller_loop(E.g1_mul(E.G1_GEN, 3), E.g2_mul(E.G2_GEN, 5))
    assert E.final_exponentiation(f) == E.final_exponentiation_naive(f)
    assert (E.X - 1) ** 2 * (E.X + E.P) * (E.X**2 + E.P**2 - 1) + 3 == 3 * (E.P**4 - E.P**2 + 1) // E.R


def test_hash_and_rng_known_answers():
    # RFC 7693 appendix A / E ("abc"), RFC 8439 2.3.2 keystream block
    assert enc.blake2b(b"abc").hex().startswith("ba80a53f981c4d0d6a2797b69f12f6e9")
    assert enc.blake2s(b"abc").hex() == "508c5e8c327c14e2e1a72ba34eeb452f37458b209ed63a294d999b4c86675982"
    key = bytes(range(32))
    # RFC 8439 test vector uses a 32-bit counter = 1 and a 96-bit nonce; with nonce words (0x09000000, 0x4a000000)
    # mapped onto our (counter_hi, nonce) layout: counter = 1 | (0x09000000 << 32), nonce words (0x4a000000, 0)
    blk = enc.chacha20_block(key, 1 | (0x09000000 << 32), (0x4A000000, 0))
    assert blk.hex().startswith("10f1e7e4d13b5915500fdd1fa32071c4")
    assert enc.fr_from_random_bytes(b"\xff" * 64) is None
    assert enc.fr_from_random_bytes(b"\x01" + b"\x00" * 63) == 1


def test_serialisation_sizes():
    assert len(enc.ser_g1(E.G1_GEN)) == 96 and len(enc.ser_g2(E.G2_GEN)) == 192
    assert enc.ser_g1(None)[0] == 0x40 and len(enc.ser_gt(E.F12_ONE)) == 576
    assert enc.ser_gt(E.F12_ONE)[:48] == (1).to_bytes(48, "little")


def test_commitments_reject_wrong_message_and_length():
    """afgho16/mod.rs:62-93, pedersen/mod.rs:40-54."""
    ck, msg, wrong = OS.g2_points("ck", N), OS.g1_points("m", N), OS.g1_points("w", N)
    com = O.AFGHOCommitmentG1.commit(ck, msg)
    assert O.AFGHOCommitmentG1.verify(ck, msg, com) and not O.AFGHOCommitmentG1.verify(ck, wrong, com)
    with pytest.raises(O.InnerProductError):
        O.AFGHOCommitmentG1.verify(ck[:-1], msg, com)
    Ped = O.PedersenCommitment(O.G1T)
    k, m = OS.g1_points("pk", N), OS.scalars("pm", N)
    com = Ped.commit(k, m)
    assert Ped.verify(k, m, com) and not Ped.verify(k, OS.scalars("pw", N), com)


def _kinds():
    G1, G2, GT, Fr = O.G1T, O.G2T, O.GTT, O.FrT
    MSM1 = O.MultiexponentiationInnerProduct(G1)
    return {
        "pairing": (O.PairingInnerProduct, O.AFGHOCommitmentG1, O.AFGHOCommitmentG2, O.IdentityCommitment(GT), ("G1", "G2", "G2", "G1")),
        "multiexp": (MSM1, O.AFGHOCommitmentG1, O.PedersenCommitment(G1), O.IdentityCommitment(G1), ("G1", "Fr", "G2", "G1")),
        "scalar": (O.ScalarInnerProduct, O.PedersenCommitment(G2), O.PedersenCommitment(G2), O.IdentityCommitment(Fr), ("Fr", "Fr", "G2", "G2")),
    }


def _gen(t, tag, n):
    return {"G1": OS.g1_points, "G2": OS.g2_points, "Fr": OS.scalars}[t](tag, n)


@pytest.mark.parametrize("name", ["pairing", "multiexp", "scalar"])
def test_gipa_round_trip_and_tamper(name):
    """gipa.rs:470-561."""
    IP, LMC, RMC, IPC, types = _kinds()[name]
    a, b, v, w = (_gen(t, "g" + str(i), N) for i, t in enumerate(types))
    g = O.GIPA(IP, LMC, RMC, IPC)
    t = IP.inner_product(a, b)
    com = (LMC.commit(v, a), RMC.commit(w, b), IPC.commit([None], [t]))
    proof = g.prove((a, b, t), (v, w, None), com)
    assert g.verify((v, w, None), com, proof)
    bad = O.GIPAProof(list(proof.r_commitment_steps), (proof.r_base[1] if name == "scalar" else proof.r_base[0], proof.r_base[1]))
    if name == "scalar":
        bad.r_base = ((proof.r_base[0] + 1) % E.R, proof.r_base[1])
    else:
        bad.r_base = (E.g1_add(proof.r_base[0], E.G1_GEN), proof.r_base[1])
    assert not g.verify((v, w, None), com, bad)
    with pytest.raises(O.InnerProductError):
        g.prove((a[:3], b[:3], t), (v[:3], w[:3], None), com)


def test_tipa_round_trip_with_srs_shift():
    """tipa/mod.rs:528-579."""
    srs = O.tipa_setup(N, OS.scalar("srs-alpha", 0), OS.scalar("srs-beta", 0))
    ck_a, ck_b = srs.get_commitment_keys()
    a, b = OS.g1_points("ta", N), OS.g2_points("tb", N)
    IP, LMC, RMC, IPC, _ = _kinds()["pairing"]
    t = O.TIPA(IP, LMC, RMC, IPC)
    r = OS.scalar("shift", 0)
    r_vec = O.structured_scalar_power(N, r)
    a_r = [E.g1_mul(p, s) for p, s in zip(a, r_vec)]
    ck_a_r = [E.g2_mul(k, pow(s, -1, E.R)) for k, s in zip(ck_a, r_vec)]
    com_a = LMC.commit(ck_a, a)
    assert com_a == IP.inner_product(a_r, ck_a_r)  # tipa/mod.rs:561
    com = (com_a, RMC.commit(ck_b, b), [IP.inner_product(a_r, b)])
    proof = t.prove_with_srs_shift(srs, (a_r, b), (ck_a_r, ck_b, None), r)
    assert t.verify_with_srs_shift(srs.get_verifier_key(), None, com, proof, r)
    assert not t.verify_with_srs_shift(srs.get_verifier_key(), None, com, proof, (r + 1) % E.R)


def test_aggregate_proofs_round_trip():
    n = 2
    srs = O.tipa_setup(n, OS.scalar("srs-alpha", 0), OS.scalar("srs-beta", 0))
    vk, proofs, inputs = OS.groth16_instance(n)
    # the simulated proofs satisfy the Groth16 equation
    A, B, Cc = proofs[0]
    ic = vk["gamma_abc_g1"][0]
    for x, g in zip(inputs[0], vk["gamma_abc_g1"][1:]):
        ic = E.g1_add(ic, E.g1_mul(g, x))
    rhs = E.gt_mul(E.gt_mul(E.pairing(vk["alpha_g1"], vk["beta_g2"]), E.pairing(ic, vk["gamma_g2"])), E.pairing(Cc, vk["delta_g2"]))
    assert E.pairing(A, B) == rhs
    agg = O.aggregate_proofs(srs, proofs)
    assert O.verify_aggregate_proof(srs.get_verifier_key(), vk, inputs, agg)
    bad_inputs = [list(x) for x in inputs]
    bad_inputs[0][0] = (bad_inputs[0][0] + 1) % E.R
    assert not O.verify_aggregate_proof(srs.get_verifier_key(), vk, bad_inputs, agg)


def test_sipp_round_trip():
    """sipp/src/lib.rs:233-254 (on BLS12-381 + Blake2s, BASELINE configs[0])."""
    n = 4
    a, b, r = OS.g1_points("sipp-a", n), OS.g2_points("sipp-b", n), OS.scalars("sipp-r", n)
    z = O.product_of_pairings_with_coeffs(a, b, r)
    proof = O.sipp_prove(a, b, r, z)
    assert O.sipp_verify(a, b, r, z, proof)
    assert not O.sipp_verify(a, b, r, E.gt_mul(z, z), proof)


def test_subgroup_membership_forms_agree():
    """The endomorphism forms of the membership tests (what the CUDA verifiers evaluate) against the definitions
    [r]P = O / f^r = 1, on members and on points of the curves outside the prime-order subgroups."""
    g, h = E.g1_mul(E.G1_GEN, 31337), E.g2_mul(E.G2_GEN, 271828)
    assert E.g1_in_subgroup(g) and E.g1_in_subgroup_fast(g) and E.g1_in_subgroup_fast(None)
    assert E.g2_in_subgroup(h) and E.g2_in_subgroup_fast(h) and E.g2_in_subgroup_fast(None)
    e = E.pairing(g, h)
    assert E.gt_in_subgroup(e) and E.gt_in_subgroup_fast(e) and E.gt_in_subgroup_fast(E.F12_ONE)
    for seed in range(3):
        p, q, c = OS.g1_point_off_subgroup(seed), OS.g2_point_off_subgroup(seed), OS.gt_cyclotomic_off_subgroup(seed)
        assert E.g1_is_on_curve(p) and not E.g1_in_subgroup(p) and not E.g1_in_subgroup_fast(p)
        assert E.g2_is_on_curve(q) and not E.g2_in_subgroup(q) and not E.g2_in_subgroup_fast(q)
        assert not E.gt_in_subgroup(c) and not E.gt_in_subgroup_fast(c)
    # a sum of a member and a non-member is a non-member
    assert not E.g1_in_subgroup_fast(E.g1_add(g, OS.g1_point_off_subgroup()))
    assert (E.X_ABS**2 - 1) ** 2 + (E.X_ABS**2 - 1) + 1 == E.R
