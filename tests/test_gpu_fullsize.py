"""Bit-exact parity at BASELINE.json's FULL sizes: the CUDA path against the compiled CPU restatement
(oracle/cpu, itself pinned to the pure-Python oracle by tests/test_oracle_cpu.py) on the same seeded inputs.

These sizes reach the kernel variants the small-n tests never select: `k_miller6<4,4>` (nseg*n >= 23 680),
the MSM plan with 13-bit windows and fat buckets, the one-thread-per-element `k_fold_endo` (n > 512), the
2^12-element GIPA rounds of the aggregation.  Inputs are generated twice -- by the CPU oracle from the synthetic
scalar streams and by the product's own GPU generator -- and compared before use, so a wrong input cannot hide a
wrong output.

  configs[0]  SIPP prove over 2^10 pairs                          sipp/src/lib.rs:42-106
  configs[1]  PairingInnerProduct / AFGHO commitment, 2^16 pairs   inner_products/src/lib.rs:56-116
  configs[2]  G1 MSM 2^18 points; GIPA multiexp prove 2^18 (slow)  inner_products/src/lib.rs:123-142, gipa.rs:162-312
  configs[3]  TIPP aggregate_proofs of 2^12 proofs + oracle verify groth16_aggregation.rs:77-231
"""
import numpy as np
import pytest

from oracle import bls12_381 as E
from oracle import protocols as O
from oracle import synth as OS
from oracle.encoding import ser_g1, ser_g2
from oracle.cpu import binding as B
from ripp_b200 import _lib, codec as C, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    return B.CppBackend()


def _g1(be, tag, n, seed=0):
    return be.mul_vec_g1(be.vec_g1([E.G1_GEN] * n), be.vec_fr(OS.scalars(tag, n, seed)))


def _g2(be, tag, n, seed=0):
    return be.mul_vec_g2(be.vec_g2([E.G2_GEN] * n), be.vec_fr(OS.scalars(tag, n, seed)))


def _u32(vec):
    """oracle/cpu packed vector -> the (n, words) uint32 limb array of the C ABI (same bytes)."""
    a = np.ascontiguousarray(vec.a)
    return a.view(np.uint32).reshape(a.shape[0], -1)


def _same(dev, vec, words):
    got = dev.download((len(vec), words))
    return np.array_equal(got, _u32(vec))


def test_pairing_product_2p16_matches_cpu_oracle(ctx, be):
    """configs[1]: 2^16 pairs select k_miller6<4,4> (four pairs per lane group share one accumulator)."""
    n = 1 << 16
    a, b = _g1(be, "cfg2-m", n), _g2(be, "cfg2-k", n)
    da, db = synth.g1_points_dev(ctx, "cfg2-m", n), synth.g2_points_dev(ctx, "cfg2-k", n)
    assert _same(da, a, 24) and _same(db, b, 48), "GPU-generated inputs (k_scale at 2^16) differ from the oracle's"
    out = ctx.alloc(576)
    ctx.pairing_ip_dev(da, db, n, out)
    ctx.sync()
    assert C.gt_dec(out.download(144)) == be.pairing_product(a, b)
    # the host-pointer entry the trait impl binds (Jacobian in, z = 1) gives the same value
    ja = np.concatenate([_u32(a), np.tile(C.fq_enc(1), (n, 1))], axis=1)
    jb = np.concatenate([_u32(b), np.tile(np.concatenate([C.fq_enc(1), C.fq_enc(0)]), (n, 1))], axis=1)
    assert C.gt_dec(ctx.pairing_ip(np.ascontiguousarray(ja), np.ascontiguousarray(jb))) == be.pairing_product(a, b)


@pytest.mark.parametrize("n", [1500, 3000, 6000, 13000])
def test_pairing_product_engine_shapes_match_cpu_oracle(ctx, be, n):
    """The sizes at which ripp_pairing_batch_l6 switches shape: eighteen-lane warps with 2 and 4 pairs per accumulator
    (1185..4736 pairs), the six-lane kernel with 1 pair per group (..11 839) and with 2 (..23 679)."""
    a, b = _g1(be, "shape-a", n), _g2(be, "shape-b", n)
    da, db = ctx.to_device(_u32(a)), ctx.to_device(_u32(b))
    out = ctx.alloc(576)
    ctx.pairing_ip_dev(da, db, n, out)
    ctx.sync()
    assert C.gt_dec(out.download(144)) == be.pairing_product(a, b)


@pytest.mark.parametrize("logn", [13, 18])
def test_msm_g1_matches_cpu_oracle(ctx, be, logn):
    """configs[2] leaf: 2^18 points (window plan c = 13, fat top-window buckets); 8191 = the KZG opening size of 2^12 proofs."""
    n = (1 << logn) - (1 if logn == 13 else 0)
    pts = _g1(be, "cfg3-a", n)
    sc = OS.scalars("cfg3-b", n)
    d = ctx.to_device(_u32(pts))
    s = ctx.to_device(C.fr_vec_enc(sc))
    out = ctx.alloc(96)
    ctx.msm_g1_dev(d, s, n, out)
    ctx.sync()
    assert C.g1_dec(out.download(24)) == be.msm_g1(pts, be.vec_fr(sc))


def test_msm_g2_8191_matches_cpu_oracle(ctx, be):
    """The G2 KZG opening MSM of a 2^12-proof aggregation (tipa/mod.rs:333-334)."""
    n = (1 << 13) - 1
    pts = _g2(be, "kzg-g2", n)
    sc = OS.scalars("kzg-q", n)
    d = ctx.to_device(_u32(pts))
    s = ctx.to_device(C.fr_vec_enc(sc))
    out = ctx.alloc(192)
    ctx.msm_g2_dev(d, s, n, out)
    ctx.sync()
    assert C.g2_dec(out.download(48)) == be.msm_g2(pts, be.vec_fr(sc))


@pytest.mark.parametrize("n", [2048, 4096 + 37])
def test_folds_above_team_threshold_match_cpu_oracle(ctx, be, n):
    """n > 512 selects the one-thread-per-element endomorphism kernels k_fold_endo<G1/G2> (mul_helper, gipa.rs:261-291)."""
    hi1, lo1 = _g1(be, "fold-hi", n), _g1(be, "fold-lo", n)
    hi2, lo2 = _g2(be, "fold-hi", n), _g2(be, "fold-lo", n)
    for c in (OS.scalar("fold-c", 0), OS.scalar("fold-c", 1) >> 120, E.R - 1):
        cw = C.fr_enc(c).copy()
        d_hi, d_lo = ctx.to_device(_u32(hi1)), ctx.to_device(_u32(lo1))
        out = ctx.alloc(n * 96)
        ctx.g1_fold_dev(d_hi, d_lo, cw, n, out)
        ctx.sync()
        assert _same(out, be.fold_g1(hi1, lo1, c), 24)
        d_hi, d_lo = ctx.to_device(_u32(hi2)), ctx.to_device(_u32(lo2))
        out = ctx.alloc(n * 192)
        ctx.g2_fold_dev(d_hi, d_lo, cw, n, out)
        ctx.sync()
        assert _same(out, be.fold_g2(hi2, lo2, c), 48)


def test_tipp_2p12_proof_bytes_match_cpu_oracle_and_oracle_verifier_accepts(ctx, be):
    """configs[3], the headline: AggregateProof bytes of 2^12 proofs equal the oracle's, and the ORACLE's
    verify_aggregate_proof accepts them."""
    from oracle import cpu_baseline

    n = 1 << 12
    work = cpu_baseline.TippWorkload(n)
    work.run()
    want = O.ser_aggregate_proof(work.proof)
    inst = synth.tipp_instance_dev(ctx, n)
    assert _same(inst["a"], work.a, 24) and _same(inst["b"], work.b, 48) and _same(inst["c"], work.c, 24)
    got = ctx.tipp_aggregate_dev(inst["srs_g1"], inst["srs_g2"], inst["a"], inst["b"], inst["c"], n)
    assert len(got) == 62544
    assert got == want, "GPU AggregateProof differs from the oracle's at 2^12 proofs"
    # the host-pointer entry point too (what the Rust shim calls)
    a_h, b_h, c_h = (np.ascontiguousarray(_u32(v)) for v in (work.a, work.b, work.c))
    assert ctx.tipp_aggregate(inst["srs_g1"], inst["srs_g2"], a_h, b_h, c_h) == want
    # oracle verifier on the oracle-side statement
    sc, _, inputs = OS.groth16_instance_scalars(n)
    vk = {
        "alpha_g1": E.g1_mul(E.G1_GEN, sc["vk"]["alpha"]), "beta_g2": E.g2_mul(E.G2_GEN, sc["vk"]["beta"]),
        "gamma_g2": E.g2_mul(E.G2_GEN, sc["vk"]["gamma"]), "delta_g2": E.g2_mul(E.G2_GEN, sc["vk"]["delta"]),
        "gamma_abc_g1": [E.g1_mul(E.G1_GEN, s) for s in sc["ic"]],
    }
    assert O.verify_aggregate_proof(work.srs.get_verifier_key(), vk, inputs, work.proof, be=be)
    # and the GPU verifier agrees on the same bytes
    assert ctx.tipp_verify_aggregate(inst["vsrs"], inst["vk"], inst["inputs"], got)


def test_sipp_2p10_matches_cpu_oracle(ctx, be):
    """configs[0]: SIPP prove over 2^10 pairs with independent r_i; value and proof bytes equal the oracle's and the
    oracle's verifier accepts."""
    n = 1 << 10
    a, b, r = _g1(be, "sipp-a", n), _g2(be, "sipp-b", n), OS.scalars("sipp-r", n)
    a_h, b_h, r_h = np.ascontiguousarray(_u32(a)), np.ascontiguousarray(_u32(b)), C.fr_vec_enc(r)
    z = ctx.sipp_product_with_coeffs(a_h, b_h, r_h)
    zo = O.product_of_pairings_with_coeffs(a, b, be.vec_fr(r), be=be)
    assert C.gt_dec(z) == zo
    got = ctx.sipp_prove(a_h, b_h, r_h, z)
    want = O.sipp_prove(list(a), list(b), r, zo, be=be)
    assert got == O.ser_sipp_proof(want)
    assert ctx.sipp_verify(a_h, b_h, r_h, z, got)


@pytest.mark.slow
def test_gipa_multiexp_2p18_proof_bytes_match_cpu_oracle(ctx, be):
    """configs[2] (ii): GIPA<MultiexpIP<G1>, AFGHO-G1, Pedersen<G1>, Identity<G1>, Blake2b> over 2^18 elements
    (benches/benches/gipa.rs:86-94): proof bytes, transcript and base keys equal the oracle's."""
    n = 1 << 18
    a, v, w = _g1(be, "cfg3-a", n), _g2(be, "cfg3-v", n), _g1(be, "cfg3-w", n)
    b = OS.scalars("cfg3-b", n)
    og = O.GIPA(O.MultiexponentiationInnerProduct(O.G1T), O.AFGHOCommitmentG1, O.PedersenCommitment(O.G1T),
                O.IdentityCommitment(O.G1T), be=be)
    want, aux = og.prove_with_aux((a, be.vec_fr(b)), (v, w, [None]))
    da, dv, dw = ctx.to_device(_u32(a)), ctx.to_device(_u32(v)), ctx.to_device(_u32(w))
    db = ctx.to_device(C.fr_vec_enc(b))
    proof, tr, ck = ctx.gipa_prove_dev(_lib.GIPA_MULTIEXP_PEDERSEN, da, db, dv, dw, n)
    assert proof == og.ser_proof(want)
    assert [C.fr_dec(t) for t in tr] == aux["r_transcript"]
    assert ck == ser_g2(aux["ck_base"][0]) + ser_g1(aux["ck_base"][1])
