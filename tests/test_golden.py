"""Golden vectors from the real arkworks reference, when somebody has produced them.

The reference cannot be built in the development image (no Rust toolchain, arkworks crates not vendored), so
the oracle is pinned by identities only ("parity unpinned").  `tools/ark_golden/` is a Rust
program that dumps the reference's outputs for the synthetic inputs below into tests/golden/ark_golden.json;
once that file exists this test pins the oracle -- and through tests/test_gpu_*.py the CUDA path -- bit for bit
against arkworks.  Without the file the test is skipped and says why."""
import json
import os

import pytest

from oracle import bls12_381 as E
from oracle import encoding as S
from oracle import protocols as O
from oracle import synth as OS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ark_golden.json")
N = 8


@pytest.fixture(scope="module")
def golden():
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/ark_golden.json absent: run tools/ark_golden with a Rust toolchain (parity unpinned)")
    return {k: bytes.fromhex(v) for k, v in json.load(open(GOLDEN)).items()}


def _srs():
    return O.tipa_setup(N, OS.scalar("srs-alpha", 0), OS.scalar("srs-beta", 0))


def test_inputs(golden):
    assert S.ser_fr(OS.scalar("gipa-b", 0)) == golden["scalar_gipa-b_0"]
    assert S.ser_g1(OS.g1_points("gipa-a", 1)[0]) == golden["g1_gipa-a_0"]
    assert S.ser_g2(OS.g2_points("gipa-b", 1)[0]) == golden["g2_gipa-b_0"]


def test_inner_products_and_commitments(golden):
    a, b = OS.g1_points("gipa-a", N), OS.g2_points("gipa-b", N)
    v, w = OS.g2_points("gipa-v", N), OS.g1_points("gipa-w", N)
    fb = OS.scalars("gipa-b", N)
    assert S.ser_gt(O.PairingInnerProduct.inner_product(a, b)) == golden["pairing_ip_n8"]
    assert S.ser_gt(O.AFGHOCommitmentG1.commit(v, a)) == golden["afgho_g1_commit_n8"]
    assert S.ser_gt(O.AFGHOCommitmentG2.commit(w, b)) == golden["afgho_g2_commit_n8"]
    assert S.ser_g1(O.MultiexponentiationInnerProduct(O.G1T).inner_product(a, fb)) == golden["msm_g1_n8"]
    assert S.ser_fr(O.ScalarInnerProduct.inner_product(OS.scalars("gipa-a", N), fb)) == golden["scalar_ip_n8"]


def test_gipa_proofs(golden):
    a, b = OS.g1_points("gipa-a", N), OS.g2_points("gipa-b", N)
    v, w = OS.g2_points("gipa-v", N), OS.g1_points("gipa-w", N)
    g = O.GIPA(O.PairingInnerProduct, O.AFGHOCommitmentG1, O.AFGHOCommitmentG2, O.IdentityCommitment(O.GTT))
    proof, _ = g.prove_with_aux((a, b), (v, w, [None]))
    assert g.ser_proof(proof) == golden["gipa_pairing_n8_proof"]
    fb = OS.scalars("gipa-b", N)
    g = O.GIPA(O.MultiexponentiationInnerProduct(O.G1T), O.AFGHOCommitmentG1, O.PedersenCommitment(O.G1T),
               O.IdentityCommitment(O.G1T))
    proof, _ = g.prove_with_aux((a, fb), (v, w, [None]))
    assert g.ser_proof(proof) == golden["gipa_multiexp_n8_proof"]


def test_tipa_proofs(golden):
    srs = _srs()
    ck_a, ck_b = srs.get_commitment_keys()
    a, b = OS.g1_points("gipa-a", N), OS.g2_points("gipa-b", N)
    t = O.TIPA(O.PairingInnerProduct, O.AFGHOCommitmentG1, O.AFGHOCommitmentG2, O.IdentityCommitment(O.GTT))
    assert t.ser_proof(t.prove(srs, (a, b), (ck_a, ck_b, None))) == golden["tipa_pairing_n8_proof"]
    assert S.ser_gt(O.AFGHOCommitmentG1.commit(ck_a, a)) == golden["tipa_pairing_n8_com_a"]
    s = OS.scalar("ssm-b", 0)
    ts = O.TIPAWithSSM(O.MultiexponentiationInnerProduct(O.G1T), O.AFGHOCommitmentG1, O.IdentityCommitment(O.G1T))
    proof = ts.prove_with_structured_scalar_message(srs, (OS.g1_points("ssm-a", N), O.structured_scalar_power(N, s)), (ck_a, None))
    assert ts.ser_proof(proof) == golden["tipa_ssm_n8_proof"]


def test_sipp(golden):
    a, b, r = OS.g1_points("sipp-a", N), OS.g2_points("sipp-b", N), OS.scalars("sipp-r", N)
    z = O.product_of_pairings_with_coeffs(a, b, r)
    assert S.ser_gt(z) == golden["sipp_n8_value"]
    if "sipp_n8_proof" in golden:
        assert O.ser_sipp_proof(O.sipp_prove(a, b, r, z)) == golden["sipp_n8_proof"]


def test_sipp_bls12_377(golden):
    """The reference's own SIPP curve.  The goldens carry arkworks' two generators; inputs are scalar(tag, i) * generator."""
    from oracle import bls12_377 as E7
    from oracle import sipp_377 as S7

    if "bls12_377_sipp_n8_value" not in golden:
        pytest.skip("golden file predates the BLS12-377 section of tools/ark_golden")

    def fq(b):
        return int.from_bytes(b, "little")

    g1b, g2b = golden["bls12_377_g1_generator"], golden["bls12_377_g2_generator"]
    g1 = (fq(g1b[:48]), fq(g1b[48:95] + bytes([g1b[95] & 0x3F])))
    g2 = ((fq(g2b[:48]), fq(g2b[48:96])), (fq(g2b[96:144]), fq(g2b[144:191] + bytes([g2b[191] & 0x3F]))))
    assert E7.g1_is_on_curve(g1) and E7.g2_is_on_curve(g2)
    assert E7.ser_g1(g1) == g1b and E7.ser_g2(g2) == g2b  # the flag convention of ark-ec's default serialisation
    assert g1 == E7.G1_GEN  # the derived generator is arkworks' generator
    sc = lambda tag: [OS.scalar(tag, i) % E7.R for i in range(N)]
    a = [E7.g1_mul(g1, s) for s in sc("s377-a")]
    b = [E7.g2_mul(g2, s) for s in sc("s377-b")]
    r = sc("s377-r")
    assert E7.ser_g1(a[0]) == golden["bls12_377_g1_s377-a_0"] and E7.ser_g2(b[0]) == golden["bls12_377_g2_s377-b_0"]
    z = S7.product_of_pairings_with_coeffs(a, b, r)
    assert E7.ser_gt(z) == golden["bls12_377_sipp_n8_value"]
    if "bls12_377_sipp_n8_proof" in golden:
        assert S7.ser_proof(S7.sipp_prove(a, b, r, z)) == golden["bls12_377_sipp_n8_proof"]
