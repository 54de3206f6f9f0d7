"""MultiexponentiationInnerProduct / Pedersen / ScalarInnerProduct and the fold kernels on the GPU
against the oracle.  Mirrors dh_commitments/src/pedersen/mod.rs:40-54 (correct verifies, wrong does
not, wrong length errors)."""
import random

import numpy as np
import pytest

from oracle import bls12_381 as E
from oracle import synth as OS
from ripp_b200 import _lib, codec as C, synth

pytestmark = pytest.mark.gpu
rnd = random.Random(99)


@pytest.mark.parametrize("n", [0, 1, 2, 5, 33, 100, 300])
def test_msm_g1_matches_oracle(ctx, n):
    from ripp_b200.inner_products import MultiexponentiationInnerProductG1 as IP

    pts = OS.g1_points("m-a", n, seed=n)
    sc = [rnd.randrange(E.R) for _ in range(n)]
    if n >= 5:
        sc[0], sc[1], sc[2] = 0, 1, E.R - 1
        pts[3] = None
    if n >= 33:  # repeated points and P, -P in the same MSM
        pts[10] = pts[11]
        pts[12] = E.g1_neg(pts[13])
        sc[12] = sc[13]
    assert IP.inner_product(pts, sc, ctx) == E.msm(pts, sc, E.g1_add, E.g1_mul)


@pytest.mark.parametrize("n", [1, 7, 64, 130])
def test_msm_g2_matches_oracle(ctx, n):
    from ripp_b200.inner_products import MultiexponentiationInnerProductG2 as IP

    pts = OS.g2_points("m-b", n, seed=n)
    sc = [rnd.randrange(E.R) for _ in range(n)]
    if n >= 7:  # scalars 0 / 1 / r - 1 and the identity among the bases
        sc[0], sc[1], sc[2] = 0, 1, E.R - 1
        pts[3] = None
    if n >= 64:  # repeated points (a doubling inside a bucket) and P, -P with equal scalars (a bucket that cancels)
        pts[10] = pts[11]
        sc[10] = sc[11]
        pts[12] = E.g2_neg(pts[13])
        sc[12] = sc[13]
        sc[20] = (1 << 128) - 1  # one endomorphism digit only
    assert IP.inner_product(pts, sc, ctx) == E.msm(pts, sc, E.g2_add, E.g2_mul)


@pytest.mark.parametrize("group", [1, 2])
def test_msm_signed_digit_corner_scalars(ctx, group):
    """Scalars chosen for the signed window recoding of msm.cu: every window digit at the carry boundary (2^(c-1),
    2^(c-1) + 1, 2^c - 1 repeated), the all-ones scalar, r - 1, and many points sharing ONE scalar (a single fat bucket per
    window, the size-sorted path with one giant class)."""
    n = 600
    pts = (OS.g1_points if group == 1 else OS.g2_points)("m-corner", n)
    pats = []
    for c in (7, 8, 13, 15, 16):
        for d in ((1 << (c - 1)), (1 << (c - 1)) + 1, (1 << c) - 1):
            pats.append(sum(d << (c * w) for w in range(255 // c)) % E.R)
    sc = [pats[i % len(pats)] for i in range(n)]
    sc[0], sc[1] = (1 << 254) - 1, E.R - 1
    for i in range(300, 600):
        sc[i] = sc[300]
    d_p = ctx.to_device((C.g1_vec_enc if group == 1 else C.g2_vec_enc)(pts))
    d_s = ctx.to_device(C.fr_vec_enc(sc))
    out = ctx.alloc(192)
    (ctx.msm_g1_dev if group == 1 else ctx.msm_g2_dev)(d_p, d_s, n, out)
    ctx.sync()
    add, mul = (E.g1_add, E.g1_mul) if group == 1 else (E.g2_add, E.g2_mul)
    got = (C.g1_dec(out.download(24)) if group == 1 else C.g2_dec(out.download(48)))
    assert got == E.msm(pts, sc, add, mul)


def test_msm_linearity_at_scale(ctx):
    """sum s_i (t_i G) = (sum s_i t_i) G at 2^14 points; oracle does one scalar multiplication."""
    n = 1 << 14
    t = synth.scalars("lin-t", n)
    s = synth.scalars("lin-s", n)
    bases = synth.g1_points_dev(ctx, "lin-t", n)
    sc = ctx.to_device(C.fr_vec_enc(s))
    out = ctx.alloc(96)
    ctx.msm_g1_dev(bases, sc, n, out)
    e = sum(a * b for a, b in zip(s, t)) % E.R
    assert C.g1_dec(out.download(24)) == E.g1_mul(E.G1_GEN, e)
    bases2 = synth.g2_points_dev(ctx, "lin-t", n)
    out2 = ctx.alloc(192)
    ctx.msm_g2_dev(bases2, sc, n, out2)
    assert C.g2_dec(out2.download(48)) == E.g2_mul(E.G2_GEN, e)


def test_pedersen_commitment(ctx):
    from ripp_b200.dh_commitments import PedersenCommitmentG1 as Ped

    n = 8
    ck = OS.g1_points("ped-ck", n)
    msg = [rnd.randrange(E.R) for _ in range(n)]
    wrong = [rnd.randrange(E.R) for _ in range(n)]
    com = Ped.commit(ck, msg, ctx)
    assert com == E.msm(ck, msg, E.g1_add, E.g1_mul)
    assert Ped.verify(ck, msg, com, ctx)
    assert not Ped.verify(ck, wrong, com, ctx)
    with pytest.raises(_lib.LengthMismatch):
        Ped.verify(ck[:-1], msg, com, ctx)


def test_scalar_inner_product(ctx):
    from ripp_b200.inner_products import ScalarInnerProduct as IP

    for n in (0, 1, 17, 1000):
        a = [rnd.randrange(E.R) for _ in range(n)]
        b = [rnd.randrange(E.R) for _ in range(n)]
        assert IP.inner_product(a, b, ctx) == sum(x * y for x, y in zip(a, b)) % E.R
    with pytest.raises(_lib.LengthMismatch):
        IP.inner_product([1, 2], [3], ctx)


@pytest.mark.parametrize("cbits", [255, 128, 1])
def test_folds(ctx, cbits):
    """out[i] = hi[i] * c + lo[i] (gipa.rs:261-291) for G1, G2 and Fr."""
    n = 37
    c = rnd.randrange(1 << (cbits - 1), min(E.R, 1 << cbits)) if cbits > 1 else 1
    ch = C.fr_enc(c).copy()
    hi1, lo1 = OS.g1_points("f-hi", n), OS.g1_points("f-lo", n)
    hi2, lo2 = OS.g2_points("f-hi", n), OS.g2_points("f-lo", n)
    lo1[0], hi1[1] = None, None
    lo1[2] = E.g1_neg(E.g1_mul(hi1[2], c))  # result is the identity
    d_hi, d_lo = ctx.to_device(C.g1_vec_enc(hi1)), ctx.to_device(C.g1_vec_enc(lo1))
    ctx.g1_fold_dev(d_hi, d_lo, ch, n, d_lo)  # in place over lo
    assert C.g1_vec_dec(d_lo.download((n, 24))) == [E.g1_add(E.g1_mul(h, c), l) for h, l in zip(hi1, lo1)]
    d_hi, d_lo = ctx.to_device(C.g2_vec_enc(hi2)), ctx.to_device(C.g2_vec_enc(lo2))
    out = ctx.alloc(n * 192)
    ctx.g2_fold_dev(d_hi, d_lo, ch, n, out)
    assert C.g2_vec_dec(out.download((n, 48))) == [E.g2_add(E.g2_mul(h, c), l) for h, l in zip(hi2, lo2)]
    a = [rnd.randrange(E.R) for _ in range(n)]
    b = [rnd.randrange(E.R) for _ in range(n)]
    d_a, d_b = ctx.to_device(C.fr_vec_enc(a)), ctx.to_device(C.fr_vec_enc(b))
    ctx.fr_fold_dev(d_a, d_b, ch, n, d_b)
    assert C.fr_vec_dec(d_b.download((n, 8))) == [(x * c + y) % E.R for x, y in zip(a, b)]


@pytest.mark.parametrize("n", [1, 3, 37])
def test_scalings_match_oracle(ctx, n):
    """out[i] = s[i] * P[i] with one scalar per element (groth16_aggregation.rs:118-131 a_r / ck_1_r; sipp/src/lib.rs:61-66):
    the lane-team kernel (k_scale_xt), scalars 0 / 1 / r-1 / 128-bit, identity inputs, and the generator form (pts = NULL)."""
    s = [rnd.randrange(E.R) for _ in range(n)]
    for j, v in enumerate([0, 1, E.R - 1, rnd.randrange(1 << 128), 2]):
        if j < n:
            s[j] = v
    p1, p2 = OS.g1_points("sc-p", n), OS.g2_points("sc-p", n)
    if n > 5:
        p1[5], p2[6] = None, None
    ds = ctx.to_device(C.fr_vec_enc(s))
    out = ctx.alloc(n * 192)
    ctx.g1_scale_dev(ctx.to_device(C.g1_vec_enc(p1)), ds, n, out)
    assert C.g1_vec_dec(out.download((n, 24))) == [E.g1_mul(p, k) for p, k in zip(p1, s)]
    ctx.g2_scale_dev(ctx.to_device(C.g2_vec_enc(p2)), ds, n, out)
    assert C.g2_vec_dec(out.download((n, 48))) == [E.g2_mul(p, k) for p, k in zip(p2, s)]
    ctx.g1_scale_dev(None, ds, n, out)
    assert C.g1_vec_dec(out.download((n, 24))) == [E.g1_mul(E.G1_GEN, k) for k in s]
    ctx.g2_scale_dev(None, ds, n, out)
    assert C.g2_vec_dec(out.download((n, 48))) == [E.g2_mul(E.G2_GEN, k) for k in s]


def test_msm_random_small_instances(ctx):
    """Randomised differential test of the MSM plan (signed digits, size-sorted buckets, fat path, lane-team Horner):
    a dozen G1 instances of random length with duplicated points, P / -P pairs, zero / one / r - 1 / short scalars and
    identities sprinkled in, each compared with the oracle's sum of scalar multiplications."""
    r2 = random.Random(20261017)
    base = OS.g1_points("m-rand", 40)
    for it in range(12):
        n = r2.choice([1, 2, 3, 7, 31, 32, 33, 64, 100, 157, 200])
        pts, sc = [], []
        for i in range(n):
            p = base[r2.randrange(len(base))]
            kind = r2.randrange(10)
            if kind == 0:
                p = None
            elif kind == 1:
                p = E.g1_neg(p)
            s = r2.choice([0, 1, E.R - 1, r2.randrange(1 << 64), r2.randrange(1 << 128), r2.randrange(E.R), r2.randrange(E.R)])
            pts.append(p)
            sc.append(s)
        d_p, d_s = ctx.to_device(C.g1_vec_enc(pts)), ctx.to_device(C.fr_vec_enc(sc))
        out = ctx.alloc(96)
        ctx.msm_g1_dev(d_p, d_s, n, out)
        ctx.sync()
        assert C.g1_dec(out.download(24)) == E.msm(pts, sc, E.g1_add, E.g1_mul), "instance %d (n = %d)" % (it, n)
