"""ORACLE (test infrastructure, NOT product code) -- BLS12-381 big-int arithmetic.

PARITY UNPINNED: the reference (arkworks-rs/ripp) ships no golden vectors and its
arithmetic lives in un-vendored crates (ark-ff / ark-ec / ark-bls12-381 "0.4", not in
/root/reference, no Cargo.lock).  This file restates the *published* algorithms those
crates implement, written independently of the CUDA code path (affine textbook group
law, textbook Miller loop over the untwisted curve, single-extension style Fq12), and
is pinned by algebraic identities (tests/test_oracle_*.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Conventions restated:
  * Fq2 = Fq[u]/(u^2+1); Fq12 = Fq2[w]/(w^6 - (1+u)); arkworks' tower c_i.c_j is the
    coefficient of w^(2j+i)                                   (A-2)
  * pairing = (optimal ate, loop |x|, conjugate because x<0) ^ (3*(p^12-1)/r)  (A-6, A-7)
    -- the cube comes from the Hayashida-Hayasaka-Teruya hard part that
    ark-ec 0.4 `Bls12::final_exponentiation` uses; call sites:
    inner_products/src/lib.rs:115, sipp/src/lib.rs:216.
"""

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
X_ABS = 0xD201000000010000  # |x|, x is negative
X = -X_ABS

G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
G2_GEN = (
    (
        0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
        0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E,
    ),
    (
        0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
        0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE,
    ),
)

# ----------------------------------------------------------------------------- Fq2
F2_ZERO = (0, 0)
F2_ONE = (1, 0)
XI = (1, 1)  # 1 + u, the Fq6/Fq12 non-residue


def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return (-a[0] % P, -a[1] % P)


def f2_mul(a, b):
    a0, a1 = a
    b0, b1 = b
    return ((a0 * b0 - a1 * b1) % P, (a0 * b1 + a1 * b0) % P)


def f2_sqr(a):
    a0, a1 = a
    return ((a0 + a1) * (a0 - a1) % P, 2 * a0 * a1 % P)


def f2_muls(a, s):
    return (a[0] * s % P, a[1] * s % P)


def f2_conj(a):
    return (a[0], -a[1] % P)


def f2_inv(a):
    a0, a1 = a
    d = pow(a0 * a0 + a1 * a1, -1, P)
    return (a0 * d % P, -a1 * d % P)


def f2_pow(a, e):
    r = F2_ONE
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


# ----------------------------------------------------------------------------- Fq12
# element = 6-tuple of Fq2, coefficient k multiplies w^k, w^6 = XI.
F12_ONE = (F2_ONE,) + (F2_ZERO,) * 5


def f12_mul(a, b):
    # schoolbook over Fq2 with lazy reduction
    acc0 = [0] * 11
    acc1 = [0] * 11
    for i in range(6):
        ai0, ai1 = a[i]
        if ai0 == 0 and ai1 == 0:
            continue
        for j in range(6):
            bj0, bj1 = b[j]
            acc0[i + j] += ai0 * bj0 - ai1 * bj1
            acc1[i + j] += ai0 * bj1 + ai1 * bj0
    out = []
    for k in range(6):
        c0, c1 = acc0[k], acc1[k]
        if k < 5:
            h0, h1 = acc0[k + 6], acc1[k + 6]
            c0 += h0 - h1  # times (1+u)
            c1 += h0 + h1
        out.append((c0 % P, c1 % P))
    return tuple(out)


def f12_sqr(a):
    return f12_mul(a, a)


def f12_conj(a):
    """w -> -w: the p^6 Frobenius (inverse on the cyclotomic subgroup)."""
    return (a[0], f2_neg(a[1]), a[2], f2_neg(a[3]), a[4], f2_neg(a[5]))


# Frobenius constants gamma[n][k] = XI^(k*(p^n-1)/6)
def _gammas(n):
    e = (P**n - 1) // 6
    g = f2_pow(XI, e)
    out = [F2_ONE]
    for _ in range(5):
        out.append(f2_mul(out[-1], g))
    return out


_GAMMA1 = _gammas(1)
_GAMMA2 = _gammas(2)


def f12_frob(a, n=1):
    n %= 12
    for _ in range(n % 2):
        a = tuple(f2_mul(f2_conj(a[k]), _GAMMA1[k]) for k in range(6))
    for _ in range(n // 2):
        a = tuple(f2_mul(a[k], _GAMMA2[k]) for k in range(6))
    return a


def f12_inv(a):
    # N = a * a^(p^6) lies in Fq6 (odd coefficients vanish)
    ac = f12_conj(a)
    n = f12_mul(a, ac)
    n2 = f12_frob(n, 2)
    n4 = f12_frob(n, 4)
    t = f12_mul(n2, n4)
    d = f12_mul(n, t)  # in Fq2
    assert all(c == F2_ZERO for c in d[1:])
    dinv = f2_inv(d[0])
    ninv = tuple(f2_mul(c, dinv) for c in t)
    return f12_mul(ac, ninv)


def f12_pow(a, e):
    if e < 0:
        return f12_pow(f12_inv(a), -e)
    r = F12_ONE
    for bit in bin(e)[2:]:
        r = f12_sqr(r)
        if bit == "1":
            r = f12_mul(r, a)
    return r


def f12_cyc_pow(a, e):
    """Exponentiation in the cyclotomic subgroup (inverse = conjugate)."""
    if e < 0:
        return f12_conj(f12_pow(a, -e))
    return f12_pow(a, e)


# ----------------------------------------------------------------------------- G1 (affine, None = infinity)
def g1_is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - 4) % P == 0


def g1_neg(pt):
    return None if pt is None else (pt[0], -pt[1] % P)


def g1_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def g1_mul(pt, k):
    k %= R
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g1_add(acc, acc)
        if bit == "1":
            acc = g1_add(acc, pt)
    return acc


# ----------------------------------------------------------------------------- G2 (affine over Fq2)
B2 = (4, 4)  # 4*(1+u)


def g2_is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return f2_sub(f2_sqr(y), f2_add(f2_mul(f2_sqr(x), x), B2)) == F2_ZERO


def g2_neg(pt):
    return None if pt is None else (pt[0], f2_neg(pt[1]))


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if f2_add(y1, y2) == F2_ZERO:
            return None
        lam = f2_mul(f2_muls(f2_sqr(x1), 3), f2_inv(f2_muls(y1, 2)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


def g2_mul(pt, k):
    k %= R
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, pt)
    return acc


# ----------------------------------------------------------------------------- pairing
def _line(lam, tx, ty, px, py):
    """Line through T (twist coords) with twist slope lam, evaluated at P, times w^3:
    (lam*x' - y') + (-lam*xP) w^2 + yP w^3."""
    c0 = f2_sub(f2_mul(lam, tx), ty)
    c2 = f2_muls(lam, -px % P)
    return (c0, F2_ZERO, c2, (py, 0), F2_ZERO, F2_ZERO)


def miller_loop(p1, q2):
    """f_{|x|,Q}(P) conjugated (x<0).  Pairs containing an identity give 1 (ark-ec A-6)."""
    if p1 is None or q2 is None:
        return F12_ONE
    px, py = p1
    tx, ty = q2
    qx, qy = q2
    f = F12_ONE
    for bit in bin(X_ABS)[3:]:
        lam = f2_mul(f2_muls(f2_sqr(tx), 3), f2_inv(f2_muls(ty, 2)))
        f = f12_mul(f12_sqr(f), _line(lam, tx, ty, px, py))
        nx = f2_sub(f2_sqr(lam), f2_muls(tx, 2))
        ty = f2_sub(f2_mul(lam, f2_sub(tx, nx)), ty)
        tx = nx
        if bit == "1":
            lam = f2_mul(f2_sub(ty, qy), f2_inv(f2_sub(tx, qx)))
            f = f12_mul(f, _line(lam, tx, ty, px, py))
            nx = f2_sub(f2_sub(f2_sqr(lam), tx), qx)
            ty = f2_sub(f2_mul(lam, f2_sub(tx, nx)), ty)
            tx = nx
    return f12_conj(f)


def _exp_by_x(a):
    return f12_cyc_pow(a, X)


def final_exponentiation(f):
    """ark-ec 0.4 Bls12::final_exponentiation (A-7): exponent 3*(p^12-1)/r."""
    f1 = f12_conj(f)
    f2 = f12_inv(f)
    r = f12_mul(f1, f2)
    f2 = r
    r = f12_mul(f12_frob(r, 2), f2)
    # hard part, exponent (x-1)^2 (x+p)(x^2+p^2-1) + 3
    y0 = f12_sqr(r)
    y1 = _exp_by_x(r)
    y2 = f12_conj(r)
    y1 = f12_mul(y1, y2)
    y2 = _exp_by_x(y1)
    y1 = f12_conj(y1)
    y1 = f12_mul(y1, y2)
    y2 = _exp_by_x(y1)
    y1 = f12_frob(y1, 1)
    y1 = f12_mul(y1, y2)
    r = f12_mul(r, y0)
    y0 = _exp_by_x(y1)
    y2 = _exp_by_x(y0)
    y0 = f12_frob(y1, 2)
    y1 = f12_conj(y1)
    y1 = f12_mul(y1, y2)
    y1 = f12_mul(y1, y0)
    return f12_mul(r, y1)


def final_exponentiation_naive(f):
    """Plain f^(3*(p^12-1)/r) -- pins the chain above."""
    return f12_pow(f, 3 * (P**12 - 1) // R)


def multi_pairing(g1s, g2s):
    """cfg_multi_pairing (inner_products/src/lib.rs:77-116): product of Miller loops, one final exp."""
    f = F12_ONE
    for a, b in zip(g1s, g2s):
        f = f12_mul(f, miller_loop(a, b))
    return final_exponentiation(f)


def pairing(a, b):
    return multi_pairing([a], [b])


# GT in arkworks is written additively: add = Fq12 mul, scalar mul = cyclotomic exp.
GT_ONE = F12_ONE


def gt_mul(a, b):
    return f12_mul(a, b)


def gt_pow(a, k):
    return f12_cyc_pow(a, k % R)


# ----------------------------------------------------------------------------- Fr helpers
def fr_inv(a):
    return pow(a, -1, R)


def msm(points, scalars, add, mul):
    acc = None
    for pt, s in zip(points, scalars):
        acc = add(acc, mul(pt, s))
    return acc


# ----------------------------------------------------------------------------- subgroup membership
# What ark-serialize's `Valid::check` enforces when the reference deserialises proof elements
# (ark-ec short_weierstrass Affine: on curve and in the prime-order subgroup; PairingOutput: order r).
# The definition ([r]P = O, f^r = 1) and the endomorphism forms the CUDA verifier evaluates
# (M. Scott, "A note on group membership tests for G1, G2 and GT on BLS pairing-friendly curves", ePrint
# 2021/1130); tests/test_oracle.py checks that the two agree on points outside the subgroups.
def _mul_raw(pt, k, add):
    """[k]pt without reducing k modulo r (pt may lie outside the prime-order subgroup)."""
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = add(acc, acc)
        if bit == "1":
            acc = add(acc, pt)
    return acc


def g1_in_subgroup(pt):
    return pt is None or (g1_is_on_curve(pt) and _mul_raw(pt, R, g1_add) is None)


def g2_in_subgroup(pt):
    return pt is None or (g2_is_on_curve(pt) and _mul_raw(pt, R, g2_add) is None)


def gt_in_subgroup(f):
    return f12_pow(f, R) == F12_ONE


ENDO_BETA = 0x1A0111EA397FE699EC02408663D4DE85AA0D857D89759AD4897D29650FB85F9B409427EB4F49FFFD8BFD00000000AAAC
_PSI_CX = f2_inv(f2_pow(XI, (P - 1) // 3))
_PSI_CY = f2_inv(f2_pow(XI, (P - 1) // 2))


def g1_phi(pt):
    """(x, y) -> (beta x, y): acts on G1 as multiplication by lambda = x^2 - 1; lambda^2 + lambda + 1 = r."""
    return None if pt is None else (pt[0] * ENDO_BETA % P, pt[1])


def g2_psi(pt):
    """Untwist-Frobenius-twist: acts on G2 as multiplication by x (= p mod r)."""
    return None if pt is None else (f2_mul(f2_conj(pt[0]), _PSI_CX), f2_mul(f2_conj(pt[1]), _PSI_CY))


def g1_in_subgroup_fast(pt):
    """phi(P) == [x^2 - 1] P; since phi^2 + phi + 1 = 0 on the whole curve this forces [r] P = O."""
    return pt is None or (g1_is_on_curve(pt) and g1_phi(pt) == _mul_raw(pt, X_ABS * X_ABS - 1, g1_add))


def g2_in_subgroup_fast(pt):
    """psi(P) == [x] P, i.e. -psi(P) == [|x|] P."""
    if pt is None:
        return True
    return g2_is_on_curve(pt) and g2_neg(g2_psi(pt)) == _mul_raw(pt, X_ABS, g2_add)


def gt_in_subgroup_fast(f):
    """f in the cyclotomic subgroup (f^(p^4) f == f^(p^2)) and f^p == f^x."""
    if f12_mul(f12_frob(f, 4), f) != f12_frob(f, 2):
        return False
    return f12_frob(f, 1) == f12_cyc_pow(f, X)
