"""The device arithmetic headers (ripp_b200/csrc/*.cuh), compiled for the host with an emulated
carry flag, against the independent Python big-int oracle.  Pins the limb sequences, tower,
group law, Miller loop and final exponentiation without a GPU; tests/test_gpu_*.py repeat the
comparison on the real PTX path."""
import random

import pytest

from oracle import bls12_381 as E
from ripp_b200 import codec as C

rnd = random.Random(2024)
rf = lambda: rnd.randrange(E.P)
rf2 = lambda: (rf(), rf())
rf12 = lambda: tuple(rf2() for _ in range(6))


@pytest.mark.parametrize("pre,p,n,radix", [("fq", E.P, 12, 1 << 384), ("fr", E.R, 8, 1 << 256)])
def test_prime_field(hostsim, pre, p, n, radix):
    rinv = pow(radix, -1, p)
    w = lambda v: C._words(v, n)
    cases = [(0, 0), (p - 1, p - 1), (1, p - 1), (0, 5)] + [(rnd.randrange(p), rnd.randrange(p)) for _ in range(500)]
    for a, b in cases:
        assert C._int(hostsim.call("hs_%s_mul" % pre, w(a), w(b), out=n)) == a * b * rinv % p
        assert C._int(hostsim.call("hs_%s_add" % pre, w(a), w(b), out=n)) == (a + b) % p
        assert C._int(hostsim.call("hs_%s_sub" % pre, w(a), w(b), out=n)) == (a - b) % p
        assert C._int(hostsim.call("hs_%s_half" % pre, w(a), w(b), out=n)) == a * pow(2, -1, p) % p
    # inversion (safegcd divsteps, modinv.cuh) on Montgomery values, incl. the ends of the range and 0 -> 0
    edge = [1, 2, 3, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, 1 << 30, (1 << 30) - 1, 1 << (32 * n - 4) if (1 << (32 * n - 4)) < p else 5,
            pow(2, -1, p), rinv, radix % p]
    for a in edge + [rnd.randrange(1, p) for _ in range(400)]:
        assert C._int(hostsim.call("hs_%s_inv" % pre, w(a * radix % p), w(0), out=n)) == pow(a, -1, p) * radix % p, a
    assert C._int(hostsim.call("hs_%s_inv" % pre, w(0), w(0), out=n)) == 0


def test_fq2(hostsim):
    for _ in range(50):
        a, b = rf2(), rf2()
        assert C.fq2_dec(hostsim.call("hs_fq2_mul", C.fq2_enc(a), C.fq2_enc(b), out=24)) == E.f2_mul(a, b)
        assert C.fq2_dec(hostsim.call("hs_fq2_sqr", C.fq2_enc(a), C.fq2_enc(b), out=24)) == E.f2_sqr(a)
        assert C.fq2_dec(hostsim.call("hs_fq2_inv", C.fq2_enc(a), C.fq2_enc(b), out=24)) == E.f2_inv(a)


def _f6(a):  # oracle has no Fq6 type: embed (c0, c1, c2) as c0 + c1 w^2 + c2 w^4
    return (a[0], (0, 0), a[1], (0, 0), a[2], (0, 0))


def test_fq6(hostsim):
    import numpy as np

    enc = lambda a: np.concatenate([C.fq2_enc(c) for c in a])
    dec = lambda w: tuple(C.fq2_dec(w[24 * i : 24 * i + 24]) for i in range(3))
    for _ in range(20):
        a, b = (rf2(), rf2(), rf2()), (rf2(), rf2(), rf2())
        got = dec(hostsim.call("hs_fq6_mul", enc(a), enc(b), out=72))
        assert _f6(got) == E.f12_mul(_f6(a), _f6(b))
        got = dec(hostsim.call("hs_fq6_sqr", enc(a), enc(b), out=72))
        assert _f6(got) == E.f12_sqr(_f6(a))
        got = dec(hostsim.call("hs_fq6_inv", enc(a), enc(b), out=72))
        assert _f6(got) == E.f12_inv(_f6(a))


def test_fq12(hostsim):
    for _ in range(10):
        a, b = rf12(), rf12()
        ea, eb = C.gt_enc(a), C.gt_enc(b)
        assert C.gt_dec(hostsim.call("hs_fq12_mul", ea, eb, out=144)) == E.f12_mul(a, b)
        assert C.gt_dec(hostsim.call("hs_fq12_sqr", ea, eb, out=144)) == E.f12_sqr(a)
        assert C.gt_dec(hostsim.call("hs_fq12_inv", ea, eb, out=144)) == E.f12_inv(a)
        for n in (1, 2, 3):
            assert C.gt_dec(hostsim.call("hs_fq12_frob%d" % n, ea, eb, out=144)) == E.f12_frob(a, n)
        d0, d1, d4 = rf2(), rf2(), rf2()
        sparse = (d0, (0, 0), d1, d4, (0, 0), (0, 0))  # d0 + d1 v + d4 v w = d0 + d1 w^2 + d4 w^3
        got = hostsim.call("hs_fq12_mul_by_014", ea, C.fq2_enc(d0), C.fq2_enc(d1), C.fq2_enc(d4), out=144)
        assert C.gt_dec(got) == E.f12_mul(a, sparse)


def test_cyclotomic(hostsim):
    a = rf12()
    c = E.f12_mul(E.f12_conj(a), E.f12_inv(a))
    c = E.f12_mul(E.f12_frob(c, 2), c)  # in the cyclotomic subgroup
    ec = C.gt_enc(c)
    assert C.gt_dec(hostsim.call("hs_fq12_cyc_sqr", ec, ec, out=144)) == E.f12_sqr(c)
    assert C.gt_dec(hostsim.call("hs_fq12_exp_by_x", ec, ec, out=144)) == E.f12_cyc_pow(c, E.X)


@pytest.mark.parametrize("grp", ["g1", "g2"])
def test_group_law(hostsim, grp):
    if grp == "g1":
        enc, dec, n, gen, add, mul, neg = C.g1_enc, C.g1_dec, 24, E.G1_GEN, E.g1_add, E.g1_mul, E.g1_neg
    else:
        enc, dec, n, gen, add, mul, neg = C.g2_enc, C.g2_dec, 48, E.G2_GEN, E.g2_add, E.g2_mul, E.g2_neg
    assert dec(hostsim.call("hs_%s_gen" % grp, out=n)) == gen
    for _ in range(5):
        a, b = mul(gen, rnd.randrange(E.R)), mul(gen, rnd.randrange(E.R))
        for x, y in ((a, b), (a, a), (a, neg(a)), (a, None), (None, b), (None, None)):
            assert dec(hostsim.call("hs_%s_add" % grp, enc(x), enc(y), out=n)) == add(x, y)
            assert dec(hostsim.call("hs_%s_add_mixed" % grp, enc(x), enc(y), out=n)) == add(x, y)
        assert dec(hostsim.call("hs_%s_dbl" % grp, enc(a), enc(a), out=n)) == add(a, a)
        for k in (0, 1, E.R - 1, rnd.randrange(E.R)):
            assert dec(hostsim.call("hs_%s_mul" % grp, enc(a), C.scalar_words(k), 255, out=n)) == mul(a, k)


def test_pairing(hostsim):
    p, q = E.g1_mul(E.G1_GEN, 123), E.g2_mul(E.G2_GEN, 456)
    f = E.miller_loop(p, q)
    ef = C.gt_enc(f)
    assert C.gt_dec(hostsim.call("hs_final_exp", ef, ef, out=144)) == E.final_exponentiation(f)
    # Miller values differ by subfield factors between formulas; they agree after the final exponentiation
    m = hostsim.call("hs_miller", C.g1_enc(p), C.g2_enc(q), out=144)
    assert C.gt_dec(hostsim.call("hs_final_exp", m, m, out=144)) == E.pairing(p, q)
    ps = [E.g1_mul(E.G1_GEN, i + 2) for i in range(4)] + [None, E.G1_GEN]
    qs = [E.g2_mul(E.G2_GEN, 3 * i + 1) for i in range(4)] + [E.G2_GEN, None]
    got = hostsim.call("hs_pairing_product", len(ps), C.g1_vec_enc(ps), C.g2_vec_enc(qs), out=144)
    assert C.gt_dec(got) == E.multi_pairing(ps, qs)


def test_op_counts(hostsim):
    """Algorithmic Fq-product counts behind the roofline model (bench.py)."""
    p, q = E.g1_mul(E.G1_GEN, 7), E.g2_mul(E.G2_GEN, 9)
    hostsim.lib.hs_mul_count_reset()
    m = hostsim.call("hs_miller", C.g1_enc(p), C.g2_enc(q), out=144)
    miller = hostsim.lib.hs_mul_count(1)
    hostsim.lib.hs_mul_count_reset()
    hostsim.call("hs_final_exp", m, m, out=144)
    fexp = hostsim.lib.hs_mul_count(1)
    import bench

    assert miller == bench.FQ_MUL_PER_MILLER_PAIR
    assert fexp == bench.FQ_MUL_PER_FINAL_EXP


@pytest.mark.parametrize("eng", ["l6", "l18"])
def test_l6_six_lane_fq12(hostsim, eng):
    """The Fq12 engine of l6.cuh against the oracle, each lane a host thread: six lanes (W = 1, one lane per Fq2
    coefficient) and eighteen (W = 3, three Karatsuba roles per coefficient exchanging through the bus)."""
    a, b = rf12(), rf12()
    ea, eb = C.gt_enc(a), C.gt_enc(b)
    op = lambda code, x, y: C.gt_dec(hostsim.call("hs_%s_op" % eng, code, x, y, out=144))
    assert op(0, ea, eb) == E.f12_mul(a, b)
    assert op(1, ea, eb) == E.f12_sqr(a)
    assert op(2, ea, eb) == E.f12_conj(a)
    assert op(3, ea, eb) == E.f12_frob(a, 1)
    assert op(4, ea, eb) == E.f12_frob(a, 2)
    assert op(5, ea, eb) == E.f12_inv(a)
    d0, d1, d4 = rf2(), rf2(), rf2()
    got = hostsim.call("hs_%s_mul_line" % eng, ea, C.fq2_enc(d0), C.fq2_enc(d1), C.fq2_enc(d4), out=144)
    assert C.gt_dec(got) == E.f12_mul(a, (d0, (0, 0), d1, d4, (0, 0), (0, 0)))
    c = E.f12_mul(E.f12_conj(a), E.f12_inv(a))
    c = E.f12_mul(E.f12_frob(c, 2), c)
    ec = C.gt_enc(c)
    assert op(8, ec, ec) == E.f12_sqr(c)  # Granger-Scott squaring in the flat basis
    assert op(6, ec, ec) == E.f12_cyc_pow(c, E.X)
    # pow_fr: generic a^e with a 256-bit canonical exponent (the verifiers' GT scalar multiplication)
    import numpy as np

    for e in (0, 1, 2, 3, rnd.randrange(E.R), E.R - 1):
        ew = np.zeros(144, dtype=np.uint32)
        ew[:8] = C.scalar_words(e)
        assert op(9, ea, ew) == E.f12_pow(a, e)


def test_l18_lazy_sums(hostsim):
    """l6.cuh's address-driven lazy sums behind the eighteen-lane Granger-Scott squaring: the single-quotient reduction
    of a 13-limb value below 2^386 (exhaustive over the 10 bits the estimate reads, both ends of every bucket), and the
    squaring against the six-lane body on inputs that drive every unreduced sum to its bound."""
    import numpy as np

    def words13(t):
        return np.array([(t >> (32 * i)) & 0xFFFFFFFF for i in range(13)], dtype=np.uint32)

    radix = 1 << 384
    h = 0
    while h < 1024:
        lo, hi = h << 376, ((h + 1) << 376) - 1
        q_est = (h * 2520) >> 16
        for t in (lo, hi):
            assert 0 <= t // E.P - q_est <= 1
        h += 1
    cases = [0, 1, E.P - 1, E.P, E.P + 1, 2 * E.P - 1, 2 * E.P, 31 * E.P, 32 * E.P - 1, (1 << 386) - 1]
    cases += [k * E.P + d for k in range(1, 40) for d in (-1, 0, 1) if k * E.P + d < (1 << 386)]
    cases += [rnd.randrange(1 << 386) for _ in range(300)]
    for t in cases:
        got = C._int(hostsim.call("hs_lz_reduce13", words13(t), out=12))
        assert got == t % E.P, hex(t)
    # one lane's share of the doubling step's passes: (sum of P slots - sum of M slots) [/ 2] [* 12] mod p on the word
    # patterns that drive the unreduced sum to its bounds (every slot p - 1, or the added ones p - 1 and the subtracted 0, ...)
    def fq_words(x):
        return np.array([(x >> (32 * i)) & 0xFFFFFFFF for i in range(12)], dtype=np.uint32)

    inv2 = pow(2, -1, E.P)
    hi = E.P - 1
    slot_sets = [[hi] * 7, [hi] * 4 + [0] * 3, [0] * 4 + [hi] * 3, [0] * 7, [1, 0, 0, 0, 0, 0, 0], [hi, hi, 0, 0, hi, 0, 0]]
    slot_sets += [[rnd.randrange(E.P) for _ in range(7)] for _ in range(40)]
    for sl in slot_sets:
        words = np.concatenate([fq_words(x) for x in sl] + [fq_words(0)])
        p4, p3, p2, m3, m1 = sum(sl[:4]), sum(sl[:3]), sum(sl[:2]), sum(sl[4:7]), sl[4]
        assert C._int(hostsim.call("hs_lz_pass", 0, 0, words, out=12)) == (p4 - m3) % E.P
        assert C._int(hostsim.call("hs_lz_pass", 0, 1, words, out=12)) == (p4 - m3) * inv2 % E.P
        assert C._int(hostsim.call("hs_lz_pass", 1, 0, words, out=12)) == (p3 - m3) % E.P
        assert C._int(hostsim.call("hs_lz_pass", 2, 0, words, out=12)) == 12 * (p2 - m1) % E.P
    # the squaring as a polynomial map (any Fq12 input, not only cyclotomic ones): extreme coefficient patterns
    big, mont = E.P - 1, lambda x: x * pow(radix, -1, E.P) % E.P  # mont(x): the value whose Montgomery form is the word pattern x
    pats = [tuple((mont(big), mont(big)) for _ in range(6)), tuple((0, 0) for _ in range(6)),
            tuple((mont(big), 0) if k < 3 else (0, mont(big)) for k in range(6)),
            tuple((mont(big - k), mont(k)) for k in range(6))] + [rf12() for _ in range(6)]
    for a in pats:
        ea = C.gt_enc(a)
        new = hostsim.call("hs_l18_op", 8, ea, ea, out=144)
        ref = hostsim.call("hs_l6_op", 8, ea, ea, out=144)
        assert (new == ref).all()
    # the eighteen-lane Fq12 product and sparse line product (operands formed by address, xi applied as an unreduced
    # role operand) against the six-lane bodies on the same patterns
    for a in pats[:5]:
        for b in (pats[0], pats[3], pats[4]):
            ea, eb = C.gt_enc(a), C.gt_enc(b)
            assert (hostsim.call("hs_l18_op", 0, ea, eb, out=144) == hostsim.call("hs_l6_op", 0, ea, eb, out=144)).all()
        assert (hostsim.call("hs_l18_op", 1, C.gt_enc(a), C.gt_enc(a), out=144) == hostsim.call("hs_l6_op", 1, C.gt_enc(a), C.gt_enc(a), out=144)).all()
        d = [C.fq2_enc(x) for x in (pats[0][0], pats[3][1], pats[4][2])]
        ea = C.gt_enc(a)
        assert (hostsim.call("hs_l18_mul_line", ea, d[0], d[1], d[2], out=144) == hostsim.call("hs_l6_mul_line", ea, d[0], d[1], d[2], out=144)).all()


@pytest.mark.parametrize("eng", ["l6", "l18"])
def test_l6_miller_and_final_exp(hostsim, eng):
    p, q = E.g1_mul(E.G1_GEN, 123), E.g2_mul(E.G2_GEN, 456)
    f = E.miller_loop(p, q)
    ef = C.gt_enc(f)
    assert C.gt_dec(hostsim.call("hs_%s_op" % eng, 7, ef, ef, out=144)) == E.final_exponentiation(f)
    import numpy as np

    one = np.array([1], dtype=np.int32)
    got = hostsim.call("hs_%s_miller" % eng, C.g1_enc(p), C.g2_enc(q), one.view(np.uint32), 1, 1, out=144)
    assert C.gt_dec(got) == E.pairing(p, q)
    # three pairs sharing one accumulator, the middle one masked (contributes 1)
    ps = [E.g1_mul(E.G1_GEN, s) for s in (5, 6, 7)]
    qs = [E.g2_mul(E.G2_GEN, s) for s in (8, 9, 10)]
    valid = np.array([1, 0, 1], dtype=np.int32)
    got = hostsim.call("hs_%s_miller" % eng, C.g1_vec_enc(ps), C.g2_vec_enc(qs), valid.view(np.uint32), 3, 1, out=144)
    assert C.gt_dec(got) == E.multi_pairing([ps[0], ps[2]], [qs[0], qs[2]])
    # four pairs with full-size random scalars, the third masked
    ps = [E.g1_mul(E.G1_GEN, rnd.randrange(E.R)) for _ in range(4)]
    qs = [E.g2_mul(E.G2_GEN, rnd.randrange(E.R)) for _ in range(4)]
    valid = np.array([1, 1, 0, 1], dtype=np.int32)
    got = hostsim.call("hs_%s_miller" % eng, C.g1_vec_enc(ps), C.g2_vec_enc(qs), valid.view(np.uint32), 4, 1, out=144)
    assert C.gt_dec(got) == E.multi_pairing([ps[0], ps[1], ps[3]], [qs[0], qs[1], qs[3]])


def test_lazy_sum_of_products(hostsim):
    """wide_mul / redc_wide (fp.cuh): REDC(sum_i a_i b_i) for up to 12 products incl. the extreme operands."""
    import ctypes

    import numpy as np

    rinv = pow(1 << 384, -1, E.P)
    for k in (1, 2, 6, 12):
        for it in range(40):
            a = [E.P - 1 if it == 0 else rnd.randrange(E.P) for _ in range(k)]
            b = [E.P - 1 if it == 0 else rnd.randrange(E.P) for _ in range(k)]
            A = np.concatenate([C._words(x, 12) for x in a])
            B = np.concatenate([C._words(x, 12) for x in b])
            got = hostsim.call("hs_fq_dot_redc", A, B, k, 3, out=12)
            assert C._int(got) == sum(x * y for x, y in zip(a, b)) * rinv % E.P


def test_x3_three_warp_team_point_arithmetic(hostsim):
    """x3.cuh (a team of three warps per vector of curve points; here each role is a host thread and the exchange a
    barrier-protected bus): k * P + L against the oracle, incl. the exceptional cases of the mixed addition."""
    import numpy as np

    def naf(k):
        pos = neg = 0
        i = 0
        while k:
            if k & 1:
                if k & 3 == 1:
                    pos |= 1 << i
                    k -= 1
                else:
                    neg |= 1 << i
                    k += 1
            k >>= 1
            i += 1
        return pos, neg, i

    def fold(group, p, lo, k):
        pos, neg, nd = naf(k)
        P = np.frombuffer(pos.to_bytes(36, "little"), dtype=np.uint32).copy()
        N = np.frombuffer(neg.to_bytes(36, "little"), dtype=np.uint32).copy()
        enc, dec, w = (C.g1_enc, C.g1_dec, 24) if group == 1 else (C.g2_enc, C.g2_dec, 48)
        return dec(hostsim.call("hs_x3_fold", group, enc(p), enc(lo), P, N, nd, out=w))

    k = rnd.randrange(E.R)
    p1, l1 = E.g1_mul(E.G1_GEN, 11), E.g1_mul(E.G1_GEN, 13)
    p2, l2 = E.g2_mul(E.G2_GEN, 11), E.g2_mul(E.G2_GEN, 13)
    assert fold(1, p1, l1, k) == E.g1_add(E.g1_mul(p1, k), l1)
    assert fold(2, p2, l2, k) == E.g2_add(E.g2_mul(p2, k), l2)
    k = 12345
    assert fold(1, p1, None, k) == E.g1_mul(p1, k) and fold(1, None, p1, k) == p1
    assert fold(1, p1, E.g1_neg(E.g1_mul(p1, k)), k) is None
    assert fold(2, p2, E.g2_mul(p2, k), k) == E.g2_mul(p2, 2 * k)
    assert fold(2, p2, E.g2_neg(E.g2_mul(p2, k)), k) is None and fold(2, None, p2, k) == p2 and fold(2, p2, None, k) == E.g2_mul(p2, k)

    # the shape k_fold_w3 runs: endomorphism digits + team formulas + barrier-safe additions + branch-free to_affine
    def efold(group, p, lo, k):
        enc, dec, w = (C.g1_enc, C.g1_dec, 24) if group == 1 else (C.g2_enc, C.g2_dec, 48)
        return dec(hostsim.call("hs_x3_endo_fold", group, enc(p), enc(lo), C.scalar_words(k), out=w))

    _check_efold(efold, p1, l1, p2, l2)


def _check_efold(efold, p1, l1, p2, l2):
    for k in (rnd.randrange(E.R), rnd.randrange(1 << 128), 1, 0, E.R - 1):
        assert efold(1, p1, l1, k) == E.g1_add(E.g1_mul(p1, k), l1), k
        assert efold(2, p2, l2, k) == E.g2_add(E.g2_mul(p2, k), l2), k
    k = rnd.randrange(1 << 128)
    assert efold(1, p1, E.g1_neg(E.g1_mul(p1, k)), k) is None and efold(2, p2, E.g2_neg(E.g2_mul(p2, k)), k) is None
    assert efold(1, p1, E.g1_mul(p1, k), k) == E.g1_mul(p1, 2 * k) and efold(2, p2, E.g2_mul(p2, k), k) == E.g2_mul(p2, 2 * k)
    assert efold(1, None, l1, k) == l1 and efold(2, None, l2, k) == l2
    assert efold(1, p1, None, k) == E.g1_mul(p1, k) and efold(2, p2, None, k) == E.g2_mul(p2, k)


def test_xt_lane_team_folds(hostsim):
    """xt.cuh: the fold out = k p + lo on a team of 3 lanes (G1) / 9 lanes (G2: 3 product units x 3 Karatsuba roles),
    level-parallel dbl-2009-l / madd-2007-bl, exceptional cases patched by the complete formulas, team normalisation."""
    p1, l1 = E.g1_mul(E.G1_GEN, 5), E.g1_mul(E.G1_GEN, 7)
    p2, l2 = E.g2_mul(E.G2_GEN, 11), E.g2_mul(E.G2_GEN, 13)

    def efold(group, p, lo, k):
        enc, dec, w = (C.g1_enc, C.g1_dec, 24) if group == 1 else (C.g2_enc, C.g2_dec, 48)
        return dec(hostsim.call("hs_xt_endo_fold", group, enc(p), enc(lo), C.scalar_words(k), out=w))

    _check_efold(efold, p1, l1, p2, l2)

    # part-parallel shape (k_fold4_xp): one team per endomorphism part, the parts added afterwards
    def pfold(group, p, lo, k):
        enc, dec, w = (C.g1_enc, C.g1_dec, 24) if group == 1 else (C.g2_enc, C.g2_dec, 48)
        return dec(hostsim.call("hs_xp_endo_fold", group, enc(p), enc(lo), C.scalar_words(k), out=w))

    _check_efold(pfold, p1, l1, p2, l2)
    for k in (E.X_ABS, E.X_ABS**2 - 1, E.X_ABS**3 + 5, 2):  # scalars that leave some parts empty
        assert pfold(1, p1, l1, k) == E.g1_add(E.g1_mul(p1, k), l1) and pfold(2, p2, l2, k) == E.g2_add(E.g2_mul(p2, k), l2)


def test_endomorphisms(hostsim):
    """phi = [x^2 - 1] on G1 and -psi = [|x|] on G2 (curve.cuh endo_map, constants from tools/gen_constants.py)."""
    p = E.g1_mul(E.G1_GEN, 4242)
    q = E.g2_mul(E.G2_GEN, 4242)
    assert C.g1_dec(hostsim.call("hs_g1_endo", C.g1_enc(p), C.g1_enc(p), out=24)) == E.g1_mul(p, E.X_ABS**2 - 1)
    assert C.g2_dec(hostsim.call("hs_g2_endo", C.g2_enc(q), C.g2_enc(q), out=48)) == E.g2_mul(q, E.X_ABS)


def test_endo_scalar_multiplication(hostsim):
    """GLV / GLS decomposition + joint double-and-add (endo.cuh) against the oracle, incl. edge scalars."""
    p = E.g1_mul(E.G1_GEN, 99)
    q = E.g2_mul(E.G2_GEN, 99)
    lam = E.X_ABS**2 - 1
    for k in (0, 1, 2, lam - 1, lam, lam + 1, E.X_ABS, E.X_ABS**2, E.X_ABS**3 + 5, E.R - 1, rnd.randrange(E.R), rnd.randrange(1 << 128)):
        kw = C.scalar_words(k)
        assert C.g1_dec(hostsim.call("hs_g1_endo_mul", C.g1_enc(p), kw, out=24)) == E.g1_mul(p, k), k
        assert C.g2_dec(hostsim.call("hs_g2_endo_mul", C.g2_enc(q), kw, out=48)) == E.g2_mul(q, k), k


@pytest.mark.parametrize("flags", [["-DRIPP_FP_UNSATURATED"], ["-DRIPP_L6_CYC_EAGER"]], ids=lambda f: f[0][2:])
def test_compile_time_variants(flags):
    """The retained compile-time A/B arms stay bit-exact: the carry-free (28-bit limb) Montgomery product -- which now also
    receives the unreduced operands of the lazy sums -- and the eighteen-lane squaring on the generic body."""
    from conftest import _build_hostsim

    hs = _build_hostsim(flags)
    a, b = rf(), rf()
    rinv = pow(1 << 384, -1, E.P)
    import numpy as np

    w = lambda v: np.array([(v >> (32 * i)) & 0xFFFFFFFF for i in range(12)], dtype=np.uint32)
    assert C._int(hs.call("hs_fq_mul", w(a), w(b), out=12)) == a * b * rinv % E.P
    x = rf12()
    c = E.f12_mul(E.f12_conj(x), E.f12_inv(x))
    c = E.f12_mul(E.f12_frob(c, 2), c)
    ec = C.gt_enc(c)
    for eng in ("l6", "l18"):
        assert C.gt_dec(hs.call("hs_%s_op" % eng, 8, ec, ec, out=144)) == E.f12_sqr(c)
        assert C.gt_dec(hs.call("hs_%s_op" % eng, 0, ec, C.gt_enc(x), out=144)) == E.f12_mul(c, x)
    p, q = E.g1_mul(E.G1_GEN, 321), E.g2_mul(E.G2_GEN, 654)
    one = np.array([1], dtype=np.int32)
    got = hs.call("hs_l18_miller", C.g1_enc(p), C.g2_enc(q), one.view(np.uint32), 1, 1, out=144)
    assert C.gt_dec(got) == E.pairing(p, q)
