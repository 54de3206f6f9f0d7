"""oracle/cpu (the compiled C++ restatement: full-size checker and CPU baseline) pinned bit-for-bit to the
pure-Python oracle (oracle/*.py): every leaf op of the backend interface on edge cases, and the whole
aggregate_proofs / SIPP transcripts at n = 8.  DESIGN.md §2 relies on this equality when the GPU path is
compared with oracle/cpu at BASELINE sizes (tests/test_gpu_fullsize.py)."""
import random

import pytest

from oracle import bls12_381 as E
from oracle import protocols as O
from oracle import synth as OS
from oracle.cpu import binding as B

rnd = random.Random(11)


@pytest.fixture(scope="module")
def be():
    return B.CppBackend()


@pytest.fixture(scope="module")
def py():
    return O.PyBackend()


def _pts(n):
    g1, g2 = OS.g1_points("cpu-a", n), OS.g2_points("cpu-b", n)
    # edge cases of SURVEY.md §8d: identity, repeated point, P and -P
    if n >= 6:
        g1[1], g2[2] = None, None
        g1[3] = g1[0]
        g2[4] = g2[0]
        g1[5] = E.g1_neg(g1[0])
    return g1, g2


@pytest.mark.parametrize("n", [0, 1, 2, 7, 33])
def test_pairing_product(be, py, n):
    g1, g2 = _pts(n)
    assert be.pairing_product(g1, g2) == py.pairing_product(g1, g2)


@pytest.mark.parametrize("n", [1, 2, 7, 40])
def test_msm_scale_fold(be, py, n):
    g1, g2 = _pts(n)
    sc = [rnd.randrange(E.R) for _ in range(n)]
    if n >= 3:
        sc[0], sc[1], sc[2] = 0, 1, E.R - 1
    assert be.msm_g1(g1, sc) == py.msm_g1(g1, sc)
    assert be.msm_g2(g2, sc) == py.msm_g2(g2, sc)
    assert list(be.mul_vec_g1(g1, sc)) == list(py.mul_vec_g1(g1, sc))
    assert list(be.mul_vec_g2(g2, sc)) == list(py.mul_vec_g2(g2, sc))
    h1, h2 = OS.g1_points("cpu-h", n), OS.g2_points("cpu-h", n)
    for c in (rnd.randrange(E.R), rnd.randrange(1 << 128), 1, E.R - 1):
        assert list(be.fold_g1(h1, g1, c)) == list(py.fold_g1(h1, g1, c))
        assert list(be.fold_g2(h2, g2, c)) == list(py.fold_g2(h2, g2, c))
    s2 = [rnd.randrange(E.R) for _ in range(n)]
    c = rnd.randrange(E.R)
    assert list(be.fold_fr(sc, s2, c)) == [(a * c + b) % E.R for a, b in zip(sc, s2)]


def test_aggregate_proofs_n8_bytes(be):
    """groth16_aggregation.rs:77-160 through both backends: identical AggregateProof bytes, and each verifier
    accepts the other's proof."""
    n = 8
    srs_py = O.tipa_setup(n, OS.scalar("srs-alpha", 0), OS.scalar("srs-beta", 0))
    srs_cpp = O.tipa_setup(n, OS.scalar("srs-alpha", 0), OS.scalar("srs-beta", 0), be)
    vk, proofs, inputs = OS.groth16_instance(n)
    want = O.aggregate_proofs(srs_py, proofs)
    got = O.aggregate_proofs(srs_cpp, proofs, be=be)
    assert O.ser_aggregate_proof(got) == O.ser_aggregate_proof(want)
    assert O.verify_aggregate_proof(srs_py.get_verifier_key(), vk, inputs, got)
    assert O.verify_aggregate_proof(srs_cpp.get_verifier_key(), vk, inputs, want, be=be)


def test_cpu_baseline_workload_is_the_same_statement(be):
    """oracle/cpu_baseline.TippWorkload (bench.py's CPU leg and the full-size parity test) builds the instance of
    oracle/synth.groth16_instance and proves it to the same bytes."""
    from oracle import cpu_baseline

    n = 4
    work = cpu_baseline.TippWorkload(n)
    work.run()
    srs = O.tipa_setup(n, OS.scalar("srs-alpha", 0), OS.scalar("srs-beta", 0))
    _, proofs, _ = OS.groth16_instance(n)
    assert [p[0] for p in proofs] == list(work.a) and [p[1] for p in proofs] == list(work.b)
    assert O.ser_aggregate_proof(work.proof) == O.ser_aggregate_proof(O.aggregate_proofs(srs, proofs))


def test_sipp_n8_bytes(be):
    n = 8
    a, b, r = OS.g1_points("sipp-a", n), OS.g2_points("sipp-b", n), OS.scalars("sipp-r", n)
    z = O.product_of_pairings_with_coeffs(a, b, r)
    assert O.product_of_pairings_with_coeffs(a, b, r, be=be) == z
    want = O.sipp_prove(a, b, r, z)
    got = O.sipp_prove(a, b, r, z, be=be)
    assert O.ser_sipp_proof(got) == O.ser_sipp_proof(want)
    assert O.sipp_verify(a, b, r, z, got, be=be)


def test_gipa_multiexp_n8_bytes(be):
    """The configs[2] instantiation (benches/benches/gipa.rs:86-94) through both backends."""
    n = 8
    a, v, w = OS.g1_points("gipa-a", n), OS.g2_points("gipa-v", n), OS.g1_points("gipa-w", n)
    b = OS.scalars("gipa-b", n)
    args = (O.MultiexponentiationInnerProduct(O.G1T), O.AFGHOCommitmentG1, O.PedersenCommitment(O.G1T), O.IdentityCommitment(O.G1T))
    p1, aux1 = O.GIPA(*args).prove_with_aux((a, b), (v, w, [None]))
    p2, aux2 = O.GIPA(*args, be=be).prove_with_aux((a, b), (v, w, [None]))
    assert O.GIPA(*args).ser_proof(p1) == O.GIPA(*args).ser_proof(p2)
    assert aux1["r_transcript"] == aux2["r_transcript"] and aux1["ck_base"] == aux2["ck_base"]
