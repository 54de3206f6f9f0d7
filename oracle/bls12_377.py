"""ORACLE (test infrastructure, NOT product code) -- BLS12-377 big-int arithmetic: the reference's own SIPP curve
(sipp/src/lib.rs:228-254 `SIPP<Bls12_377, Blake2s>`, sipp/examples/scaling-ipp.rs:10).

PARITY UNPINNED, like oracle/bls12_381.py: ark-bls12-377 0.4 is an un-vendored dependency.  The parameters are DERIVED
here from the BLS12 family polynomials at x = 0x8508c00000000001 and checked (primality, curve orders, twist, subgroup
generators); the conventions restated are SURVEY.md App. A-2 / A-4 / A-6 / A-7:
  * Fq2 = Fq[u]/(u^2 + 5); Fq12 = Fq2[w]/(w^6 - u) (Fq6 non-residue u); arkworks' tower c_i.c_j = coefficient of w^(2j+i)
  * G1: y^2 = x^3 + 1;  G2 on the D-type twist y^2 = x^3 + 1/u; untwist (x', y') -> (x' w^2, y' w^3)
  * line evaluation placed at w^0, w^1, w^3 (ark-ec `mul_by_034`), x > 0: no conjugation after the Miller loop
  * final exponentiation: the same Hayashida-Hayasaka-Teruya chain as BLS12-381, exponent 3 (p^12 - 1) / r
Same function names as oracle/bls12_381.py so protocol code can take either module.

Only tests/ may import this.
"""

X = 0x8508C00000000001
X_ABS = X
R = X**4 - X**2 + 1
P = (X - 1) ** 2 * R // 3 + X
assert ((X - 1) ** 2 * R) % 3 == 0
assert P == 0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001
assert R == 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
BETA = 5  # u^2 = -5
H1 = (X - 1) ** 2 // 3
H2 = (X**8 - 4 * X**7 + 5 * X**6 - 4 * X**4 + 6 * X**3 - 4 * X**2 - 4 * X + 13) // 9

# ----------------------------------------------------------------------------- Fq2
F2_ZERO = (0, 0)
F2_ONE = (1, 0)
XI = (0, 1)  # u, the Fq6 / Fq12 non-residue


def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return (-a[0] % P, -a[1] % P)


def f2_mul(a, b):
    a0, a1 = a
    b0, b1 = b
    return ((a0 * b0 - BETA * a1 * b1) % P, (a0 * b1 + a1 * b0) % P)


def f2_sqr(a):
    a0, a1 = a
    return ((a0 * a0 - BETA * a1 * a1) % P, 2 * a0 * a1 % P)


def f2_muls(a, s):
    return (a[0] * s % P, a[1] * s % P)


def f2_conj(a):
    return (a[0], -a[1] % P)


def f2_inv(a):
    a0, a1 = a
    d = pow(a0 * a0 + BETA * a1 * a1, -1, P)
    return (a0 * d % P, -a1 * d % P)


def f2_pow(a, e):
    r = F2_ONE
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


def f2_mul_xi(a):
    """a * u = -5 a1 + a0 u."""
    return (-BETA * a[1] % P, a[0])


# ----------------------------------------------------------------------------- Fq12: 6 x Fq2, coefficient k multiplies w^k, w^6 = u
F12_ONE = (F2_ONE,) + (F2_ZERO,) * 5


def f12_mul(a, b):
    acc = [F2_ZERO] * 11
    for i in range(6):
        if a[i] == F2_ZERO:
            continue
        for j in range(6):
            acc[i + j] = f2_add(acc[i + j], f2_mul(a[i], b[j]))
    return tuple(f2_add(acc[k], f2_mul_xi(acc[k + 6])) if k < 5 else acc[k] for k in range(6))


def f12_sqr(a):
    return f12_mul(a, a)


def f12_conj(a):
    return (a[0], f2_neg(a[1]), a[2], f2_neg(a[3]), a[4], f2_neg(a[5]))


def _gammas(n):
    g = f2_pow(XI, (P**n - 1) // 6)
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
    ac = f12_conj(a)
    n = f12_mul(a, ac)
    n2 = f12_frob(n, 2)
    n4 = f12_frob(n, 4)
    t = f12_mul(n2, n4)
    d = f12_mul(n, t)
    assert all(c == F2_ZERO for c in d[1:])
    dinv = f2_inv(d[0])
    return f12_mul(ac, tuple(f2_mul(c, dinv) for c in t))


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
    if e < 0:
        return f12_conj(f12_pow(a, -e))
    return f12_pow(a, e)


# ----------------------------------------------------------------------------- square roots (p = 1 mod 2^46: Tonelli-Shanks)
def fq_sqrt(a):
    a %= P
    if a == 0:
        return 0
    if pow(a, (P - 1) // 2, P) != 1:
        return None
    q, s = P - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (P - 1) // 2, P) != P - 1:
        z += 1
    m, c, t, r = s, pow(z, q, P), pow(a, q, P), pow(a, (q + 1) // 2, P)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % P
            i += 1
        b = pow(c, 1 << (m - i - 1), P)
        m, c = i, b * b % P
        t, r = t * c % P, r * b % P
    return r


def f2_sqrt(a):
    """Square root in Fq[u]/(u^2 + 5) by the norm method (None if a is not a square)."""
    a0, a1 = a
    if a1 == 0:
        s = fq_sqrt(a0)
        if s is not None:
            return (s, 0)
        s = fq_sqrt(-a0 * pow(BETA, -1, P) % P)  # (s u)^2 = -5 s^2
        return None if s is None else (0, s)
    n = fq_sqrt((a0 * a0 + BETA * a1 * a1) % P)  # norm
    if n is None:
        return None
    inv2 = pow(2, -1, P)
    for sign in (1, -1):
        x0sq = (a0 + sign * n) * inv2 % P
        x0 = fq_sqrt(x0sq)
        if x0 is None or x0 == 0:
            continue
        x1 = a1 * pow(2 * x0, -1, P) % P
        if f2_sqr((x0, x1)) == (a0 % P, a1 % P):
            return (x0, x1)
    return None


# ----------------------------------------------------------------------------- G1: y^2 = x^3 + 1
B1 = 1


def g1_is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B1) % P == 0


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


def _mul_raw(pt, k, add):
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = add(acc, acc)
        if bit == "1":
            acc = add(acc, pt)
    return acc


def g1_mul(pt, k):
    return _mul_raw(pt, k % R, g1_add)


# ----------------------------------------------------------------------------- G2: y^2 = x^3 + 1/u over Fq2 (D-type twist)
B2 = f2_inv(XI)


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
    return _mul_raw(pt, k % R, g2_add)


def _find_generators():
    """Subgroup generators derived deterministically: the smallest x (G1) / x = (k, 1) (G2) giving a curve point, times the
    cofactor.  (SIPP's own inputs are random subgroup points, so WHICH generator is used does not enter the protocol.)"""
    x = 1
    while True:
        y = fq_sqrt(x * x * x + B1)
        if y is not None:
            g1 = _mul_raw((x, min(y, P - y)), H1, g1_add)
            if g1 is not None:
                break
        x += 1
    k = 1
    while True:
        xx = (k, 1)
        y = f2_sqrt(f2_add(f2_mul(f2_sqr(xx), xx), B2))
        if y is not None:
            g2 = _mul_raw((xx, y), H2, g2_add)
            if g2 is not None:
                break
        k += 1
    return g1, g2


G1_GEN, G2_GEN = _find_generators()


# ----------------------------------------------------------------------------- pairing
def _line(lam, tx, ty, px, py):
    """Line through T = (x' w^2, y' w^3) (untwisted) with twist slope lam, evaluated at P:
    yP - lam xP w + (lam x' - y') w^3."""
    return ((py % P, 0), f2_muls(lam, -px % P), F2_ZERO, f2_sub(f2_mul(lam, tx), ty), F2_ZERO, F2_ZERO)


def miller_loop(p1, q2):
    """f_{x,Q}(P) (x > 0: no conjugation).  Pairs containing an identity give 1."""
    if p1 is None or q2 is None:
        return F12_ONE
    px, py = p1
    tx, ty = q2
    qx, qy = q2
    f = F12_ONE
    for bit in bin(X)[3:]:
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
    return f


def _exp_by_x(a):
    return f12_cyc_pow(a, X)


def final_exponentiation(f):
    """ark-ec 0.4 Bls12::final_exponentiation (generic over the curve config): exponent 3 (p^12 - 1) / r."""
    f1 = f12_conj(f)
    f2 = f12_inv(f)
    r = f12_mul(f1, f2)
    f2 = r
    r = f12_mul(f12_frob(r, 2), f2)
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
    return f12_pow(f, 3 * (P**12 - 1) // R)


def multi_pairing(g1s, g2s):
    f = F12_ONE
    for a, b in zip(g1s, g2s):
        f = f12_mul(f, miller_loop(a, b))
    return final_exponentiation(f)


def pairing(a, b):
    return multi_pairing([a], [b])


GT_ONE = F12_ONE


def gt_mul(a, b):
    return f12_mul(a, b)


def gt_pow(a, k):
    return f12_cyc_pow(a, k % R)


def fr_inv(a):
    return pow(a, -1, R)


def msm(points, scalars, add, mul):
    acc = None
    for pt, s in zip(points, scalars):
        acc = add(acc, mul(pt, s))
    return acc


# ----------------------------------------------------------------------------- ark-serialize 0.4 default formats (App. A-4)
def ser_fq(v):
    return (v % P).to_bytes(48, "little")


def ser_fr(v):
    return (v % R).to_bytes(32, "little")


def ser_gt(f):
    """Fq12 = c0 || c1, Fq6 = c0 || c1 || c2, Fq2 = c0 || c1: arkworks' c_i.c_j is flat coefficient 2 j + i."""
    out = b""
    for i in range(2):
        for j in range(3):
            c = f[2 * j + i]
            out += ser_fq(c[0]) + ser_fq(c[1])
    return out


def _flagged(b, flags):
    return b[:-1] + bytes([b[-1] | flags])


def ser_g1(pt):
    """short_weierstrass default, uncompressed: x (LE) || y (LE), flags in the top bits of the LAST byte:
    bit 7 = y is the 'negative' root (y > -y), bit 6 = infinity (x = y = 0)."""
    if pt is None:
        return ser_fq(0) + _flagged(ser_fq(0), 0x40)
    x, y = pt
    return ser_fq(x) + _flagged(ser_fq(y), 0x80 if y > (P - y) % P else 0)


def _f2_gt(a, b):
    """Ordering of Fq2 in ark-ff: lexicographic on (c1, c0)."""
    return (a[1], a[0]) > (b[1], b[0])


def ser_g2(pt):
    if pt is None:
        return ser_fq(0) * 3 + _flagged(ser_fq(0), 0x40)
    x, y = pt
    body = ser_fq(x[0]) + ser_fq(x[1]) + ser_fq(y[0]) + ser_fq(y[1])
    return _flagged(body, 0x80 if _f2_gt(y, f2_neg(y)) else 0)
