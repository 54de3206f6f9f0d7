"""Setup entry points and the checked GIPA::prove on the GPU against the oracle.
Mirrors: structured_generators_scalar_power (tipa/mod.rs:372-391), TIPA::setup (tipa/mod.rs:150-164), random_generators
behind AFGHO16 / Pedersen `setup` (dh_commitments/src/lib.rs:59-61), GIPA::prove's pre-checks (gipa.rs:108-133)."""
import random

import numpy as np
import pytest

from oracle import bls12_381 as E
from oracle import protocols as O
from oracle import synth as OS
from ripp_b200 import _lib, codec as C
from ripp_b200.dh_commitments import AFGHOCommitmentG1, PedersenCommitmentG1, generators_from_scalars
from ripp_b200.ip_proofs import GIPA, TIPA, InnerProductArgumentError, structured_generators_scalar_power
from test_gpu_protocols import _inputs, _oracle_gipa

pytestmark = pytest.mark.gpu
rnd = random.Random(17)


@pytest.fixture(scope="module")
def be():
    from oracle.cpu import binding as B

    return B.CppBackend()


def _u32(vec):
    a = np.ascontiguousarray(vec.a)
    return a.view(np.uint32).reshape(a.shape[0], -1)


@pytest.mark.parametrize("num", [1, 5, 70])
def test_structured_generators_match_oracle(ctx, num):
    """Short vectors (the per-element scaling path) with the generator and with a caller-supplied base."""
    s = rnd.randrange(E.R)
    pw = O.structured_scalar_power(num, s)
    assert structured_generators_scalar_power(num, None, s, 1, ctx) == [E.g1_mul(E.G1_GEN, k) for k in pw]
    assert structured_generators_scalar_power(num, None, s, 2, ctx) == [E.g2_mul(E.G2_GEN, k) for k in pw]
    b1, b2 = OS.g1_points("sg-base", 1)[0], OS.g2_points("sg-base", 1)[0]
    assert structured_generators_scalar_power(num, b1, s, 1, ctx) == [E.g1_mul(b1, k) for k in pw]
    assert structured_generators_scalar_power(num, b2, s, 2, ctx) == [E.g2_mul(b2, k) for k in pw]


def test_fixed_base_table_path_matches_variable_base(ctx, be):
    """2^12 outputs take the windowed table (FB_C = 8, 32 windows); every output is compared with the CPU oracle's
    scalar multiplication, scalars 0 / 1 / r - 1 / 2^248 included (digit 0 in every window / only the top window)."""
    n = 4096
    s = [rnd.randrange(E.R) for _ in range(n)]
    s[0], s[1], s[2], s[3] = 0, 1, E.R - 1, 1 << 248
    ds = ctx.to_device(C.fr_vec_enc(s))
    out = ctx.alloc(n * 192)
    ctx.fixed_base_msm_dev(1, None, ds, n, out)
    assert np.array_equal(out.download((n, 24)), _u32(be.mul_vec_g1(be.vec_g1([E.G1_GEN] * n), be.vec_fr(s))))
    ctx.fixed_base_msm_dev(2, None, ds, n, out)
    assert np.array_equal(out.download((n, 48)), _u32(be.mul_vec_g2(be.vec_g2([E.G2_GEN] * n), be.vec_fr(s))))


def test_tipa_setup_matches_oracle(ctx):
    """TIPA::setup for given trapdoors: both towers of 2 size - 1 powers (table path at size 2048), g^beta, h^alpha; the
    verifier key is what get_verifier_key returns (tipa/mod.rs:120-127)."""
    alpha, beta = OS.scalar("srs-alpha", 0), OS.scalar("srs-beta", 0)
    for size in (4, 2048):
        srs, v_srs = TIPA.setup(alpha, beta, size, ctx)
        m = 2 * size - 1
        g1 = C.g1_vec_dec(srs["g_alpha_powers"].download((m, 24)))
        g2 = C.g2_vec_dec(srs["h_beta_powers"].download((m, 48)))
        idx = range(m) if size == 4 else [0, 1, 2, 255, 256, m // 2, m - 2, m - 1]
        for i in idx:
            assert g1[i] == E.g1_mul(E.G1_GEN, pow(alpha, i, E.R))
            assert g2[i] == E.g2_mul(E.G2_GEN, pow(beta, i, E.R))
        assert srs["g_beta"] == E.g1_mul(E.G1_GEN, beta) and srs["h_alpha"] == E.g2_mul(E.G2_GEN, alpha)
        assert v_srs == {"g": E.G1_GEN, "h": E.G2_GEN, "g_beta": srs["g_beta"], "h_alpha": srs["h_alpha"]}
    want = O.tipa_setup(4, alpha, beta)
    srs, _ = TIPA.setup(alpha, beta, 4, ctx)
    assert C.g1_vec_dec(srs["g_alpha_powers"].download((7, 24))) == list(want.g_alpha_powers)
    assert C.g2_vec_dec(srs["h_beta_powers"].download((7, 48))) == list(want.h_beta_powers)


def test_commitment_setup_keys(ctx):
    """`setup` of the commitments = random_generators with the exponents supplied by the caller; the keys commit and
    verify as the reference's tests do (afgho16/mod.rs:62-87, pedersen/mod.rs:40-54)."""
    n = 6
    ex = OS.scalars("ck-setup", n)
    k2 = AFGHOCommitmentG1.setup(ex, ctx)
    k1 = PedersenCommitmentG1.setup(ex, ctx)
    assert k2 == [E.g2_mul(E.G2_GEN, e) for e in ex] and k1 == [E.g1_mul(E.G1_GEN, e) for e in ex]
    assert generators_from_scalars([], 1, ctx) == []
    m = OS.g1_points("ck-msg", n)
    com = AFGHOCommitmentG1.commit(k2, m, ctx)
    assert AFGHOCommitmentG1.verify(k2, m, com, ctx)
    assert not AFGHOCommitmentG1.verify(k2, OS.g1_points("ck-other", n), com, ctx)


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 6])
def test_gipa_checked_prove(ctx, kind):
    """GIPA::prove: same bytes as prove_with_aux for a true statement; InnerProductInvalid for a wrong inner product,
    a wrong left commitment, a wrong right commitment; the inner-product check comes before the length check."""
    a, b, v, w = _inputs(kind, 8)
    IP, LMC, RMC, IPC = _oracle_gipa(kind)
    og = O.GIPA(IP, LMC, RMC, IPC)
    t = IP.inner_product(a, b)
    ssm = w[0] is None
    com_a, com_b = LMC.commit(v, a), (None if ssm else RMC.commit(w, b))
    com = (com_a, t) if ssm else (com_a, com_b, t)
    g = GIPA(kind, ctx)
    ck = (v, None if ssm else w, None)
    want, _ = og.prove_with_aux((a, b), (v, w, [None]))
    assert g.prove((a, b, t), ck, com) == og.ser_proof(want)
    # the oracle's checked entry agrees on the true statement
    assert og.ser_proof(og.prove((a, b, t), (v, w, None), (com_a, RMC.commit(w, b), IPC.commit([None], [t])))) == og.ser_proof(want)
    a2, b2, v2, w2 = _inputs(kind, 8, seed=1)
    t_bad = IP.inner_product(a2, b2)
    with pytest.raises(InnerProductArgumentError, match="InnerProductInvalid"):
        g.prove((a, b, t_bad), ck, (com_a, t_bad) if ssm else (com_a, com_b, t_bad))
    bad_a = LMC.commit(v, a2)
    with pytest.raises(InnerProductArgumentError, match="InnerProductInvalid"):
        g.prove((a, b, t), ck, (bad_a, t) if ssm else (bad_a, com_b, t))
    if not ssm:
        with pytest.raises(InnerProductArgumentError, match="InnerProductInvalid"):
            g.prove((a, b, t), ck, (com_a, RMC.commit(w, b2), t))
    # gipa.rs:113-122: a length-6 instance with a wrong value fails on the value, with the right value on the length
    a6, b6, v6, w6 = _inputs(kind, 6)
    t6 = IP.inner_product(a6, b6)
    c6 = LMC.commit(v6, a6)
    ck6 = (v6, None if ssm else w6, None)
    with pytest.raises(InnerProductArgumentError, match="InnerProductInvalid"):
        g.prove((a6, b6, t_bad), ck6, (c6, t_bad) if ssm else (c6, RMC.commit(w6, b6), t_bad))
    with pytest.raises(InnerProductArgumentError, match="left length"):
        g.prove((a6, b6, t6), ck6, (c6, t6) if ssm else (c6, RMC.commit(w6, b6), t6))
