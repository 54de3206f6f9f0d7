"""GPU parity of the PTX limb layer and everything built on it: each primitive op runs on the
device through the C ABI (ripp_test_elementwise) and is compared bit-for-bit with the Python
big-int oracle on the same seeded inputs."""
import random

import numpy as np
import pytest

from oracle import bls12_381 as E
from ripp_b200 import codec as C

pytestmark = pytest.mark.gpu
rnd = random.Random(7)
rf = lambda: rnd.randrange(E.P)
rf2 = lambda: (rf(), rf())
rf12 = lambda: tuple(rf2() for _ in range(6))


@pytest.mark.parametrize("pre,p,n,radix", [("FQ", E.P, 12, 1 << 384), ("FR", E.R, 8, 1 << 256)])
def test_prime_field(ctx, pre, p, n, radix):
    rinv = pow(radix, -1, p)
    cases = [(0, 0), (p - 1, p - 1), (1, p - 1), (0, 5), (p - 1, 1)] + [
        (rnd.randrange(p), rnd.randrange(p)) for _ in range(2000)
    ]
    a = np.stack([C._words(x, n) for x, _ in cases])
    b = np.stack([C._words(y, n) for _, y in cases])
    mul = ctx.test_elementwise(pre + "_MUL", a, b, n)
    add = ctx.test_elementwise(pre + "_ADD", a, b, n)
    sub = ctx.test_elementwise(pre + "_SUB", a, b, n)
    for i, (x, y) in enumerate(cases):
        assert C._int(mul[i]) == x * y * rinv % p
        assert C._int(add[i]) == (x + y) % p
        assert C._int(sub[i]) == (x - y) % p
    if pre == "FQ":
        half = ctx.test_elementwise("FQ_HALF", a, b, n)
        for i, (x, _) in enumerate(cases):
            assert C._int(half[i]) == x * pow(2, -1, p) % p
    vals = [rnd.randrange(1, p) for _ in range(64)]
    a = np.stack([C._words(v * radix % p, n) for v in vals])
    inv = ctx.test_elementwise(pre + "_INV", a, a, n)
    for i, v in enumerate(vals):
        assert C._int(inv[i]) == pow(v, -1, p) * radix % p


def test_fq2(ctx):
    xs, ys = [rf2() for _ in range(200)], [rf2() for _ in range(200)]
    a, b = np.stack([C.fq2_enc(x) for x in xs]), np.stack([C.fq2_enc(y) for y in ys])
    mul, sqr, inv = (ctx.test_elementwise(op, a, b, 24) for op in ("FQ2_MUL", "FQ2_SQR", "FQ2_INV"))
    for i in range(200):
        assert C.fq2_dec(mul[i]) == E.f2_mul(xs[i], ys[i])
        assert C.fq2_dec(sqr[i]) == E.f2_sqr(xs[i])
        assert C.fq2_dec(inv[i]) == E.f2_inv(xs[i])


def test_fq12(ctx):
    n = 40
    xs, ys = [rf12() for _ in range(n)], [rf12() for _ in range(n)]
    a, b = np.stack([C.gt_enc(x) for x in xs]), np.stack([C.gt_enc(y) for y in ys])
    mul, sqr, inv, fr1 = (ctx.test_elementwise(op, a, b, 144) for op in ("FQ12_MUL", "FQ12_SQR", "FQ12_INV", "FQ12_FROB1"))
    for i in range(n):
        assert C.gt_dec(mul[i]) == E.f12_mul(xs[i], ys[i])
        assert C.gt_dec(sqr[i]) == E.f12_sqr(xs[i])
        assert C.gt_dec(inv[i]) == E.f12_inv(xs[i])
        assert C.gt_dec(fr1[i]) == E.f12_frob(xs[i], 1)
    cyc = []
    for x in xs[:8]:
        c = E.f12_mul(E.f12_conj(x), E.f12_inv(x))
        cyc.append(E.f12_mul(E.f12_frob(c, 2), c))
    a = np.stack([C.gt_enc(c) for c in cyc])
    got = ctx.test_elementwise("FQ12_CYC_SQR", a, a, 144)
    for i, c in enumerate(cyc):
        assert C.gt_dec(got[i]) == E.f12_sqr(c)


@pytest.mark.parametrize("grp", ["G1", "G2"])
def test_group_law(ctx, grp):
    if grp == "G1":
        enc, dec, w, gen, add, mul, neg = C.g1_enc, C.g1_dec, 24, E.G1_GEN, E.g1_add, E.g1_mul, E.g1_neg
    else:
        enc, dec, w, gen, add, mul, neg = C.g2_enc, C.g2_dec, 48, E.G2_GEN, E.g2_add, E.g2_mul, E.g2_neg
    pts = [mul(gen, rnd.randrange(E.R)) for _ in range(12)]
    pairs = [(pts[i], pts[i + 1]) for i in range(0, 12, 2)]
    pairs += [(pts[0], pts[0]), (pts[1], neg(pts[1])), (pts[2], None), (None, pts[3]), (None, None)]
    a, b = np.stack([enc(x) for x, _ in pairs]), np.stack([enc(y) for _, y in pairs])
    got_add = ctx.test_elementwise(grp + "_ADD", a, b, w)
    got_dbl = ctx.test_elementwise(grp + "_DBL", a, b, w)
    for i, (x, y) in enumerate(pairs):
        assert dec(got_add[i]) == add(x, y)
        assert dec(got_dbl[i]) == add(x, x)


def test_miller_and_final_exp(ctx):
    ss = [rnd.randrange(1, E.R) for _ in range(6)]
    ps = [E.g1_mul(E.G1_GEN, s) for s in ss] + [None]
    qs = [E.g2_mul(E.G2_GEN, s + 1) for s in ss] + [E.G2_GEN]
    m = ctx.test_elementwise("MILLER", C.g1_vec_enc(ps), C.g2_vec_enc(qs), 144)
    fe = ctx.test_elementwise("FINAL_EXP", m, m, 144)
    for i in range(len(ps)):
        assert C.gt_dec(fe[i]) == E.pairing(ps[i], qs[i])
    # the oracle's own Miller values through the device final exponentiation
    fo = [E.miller_loop(ps[i], qs[i]) for i in range(3)]
    a = np.stack([C.gt_enc(f) for f in fo])
    fe = ctx.test_elementwise("FINAL_EXP", a, a, 144)
    for i in range(3):
        assert C.gt_dec(fe[i]) == E.final_exponentiation(fo[i])
