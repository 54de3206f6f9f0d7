"""Conversions between Python integers / tuples and the packed little-endian 32-bit-limb arrays of
the C ABI (include/ripp_b200.h).  Field elements cross the ABI in Montgomery form (R = 2^384 for
Fq, 2^256 for Fr), bit-identical to arkworks' in-memory `Fp` limbs.

Affine points: G1 = x || y (24 words), G2 = x.c0 || x.c1 || y.c0 || y.c1 (48 words); the identity
is all-zero.  GT / Fq12 = 144 words in arkworks struct order c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2.
Python-side tuples follow the layout the tests use: Fq2 = (c0, c1); Fq12 = 6 Fq2 coefficients of
w^0..w^5 (slot c_i.c_j is the w^(2j+i) coefficient); points are affine tuples or None.
"""
import numpy as np

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
RQ = 1 << 384
RR = 1 << 256
RQ_INV = pow(RQ, -1, P)
RR_INV = pow(RR, -1, R)

_TOWER_ORDER = (0, 2, 4, 1, 3, 5)  # memory slot -> w-power


def _words(v, n):
    return np.frombuffer(int(v).to_bytes(4 * n, "little"), dtype=np.uint32)


def _int(a):
    return int.from_bytes(np.ascontiguousarray(a, dtype=np.uint32).tobytes(), "little")


def fq_enc(v):
    return _words(v % P * RQ % P, 12)


def fq_dec(a):
    return _int(a) * RQ_INV % P


def fr_enc(v):
    return _words(v % R * RR % R, 8)


def fr_dec(a):
    return _int(a) * RR_INV % R


def fr_vec_enc(vals):
    out = np.empty((len(vals), 8), dtype=np.uint32)
    for i, v in enumerate(vals):
        out[i] = fr_enc(v)
    return out


def fr_vec_dec(arr):
    return [fr_dec(row) for row in np.asarray(arr).reshape(-1, 8)]


def fq2_enc(v):
    return np.concatenate([fq_enc(v[0]), fq_enc(v[1])])


def fq2_dec(a):
    return (fq_dec(a[:12]), fq_dec(a[12:24]))


def g1_enc(pt):
    if pt is None:
        return np.zeros(24, dtype=np.uint32)
    return np.concatenate([fq_enc(pt[0]), fq_enc(pt[1])])


def g1_dec(a):
    a = np.asarray(a)
    if not a.any():
        return None
    return (fq_dec(a[:12]), fq_dec(a[12:24]))


def g2_enc(pt):
    if pt is None:
        return np.zeros(48, dtype=np.uint32)
    return np.concatenate([fq2_enc(pt[0]), fq2_enc(pt[1])])


def g2_dec(a):
    a = np.asarray(a)
    if not a.any():
        return None
    return (fq2_dec(a[:24]), fq2_dec(a[24:48]))


def g1_vec_enc(pts):
    out = np.empty((len(pts), 24), dtype=np.uint32)
    for i, p in enumerate(pts):
        out[i] = g1_enc(p)
    return out


def g2_vec_enc(pts):
    out = np.empty((len(pts), 48), dtype=np.uint32)
    for i, p in enumerate(pts):
        out[i] = g2_enc(p)
    return out


def g1_vec_dec(arr):
    return [g1_dec(r) for r in np.asarray(arr).reshape(-1, 24)]


def g2_vec_dec(arr):
    return [g2_dec(r) for r in np.asarray(arr).reshape(-1, 48)]


def g1_jac_enc(pt, z=1):
    """Jacobian (X, Y, Z) words for an affine point scaled by z (arkworks `Projective` layout)."""
    if pt is None:
        return np.concatenate([fq_enc(1), fq_enc(1), fq_enc(0)])
    return np.concatenate([fq_enc(pt[0] * z * z), fq_enc(pt[1] * z * z * z), fq_enc(z)])


def g1_jac_dec(a):
    a = np.asarray(a)
    z = fq_dec(a[24:36])
    if z == 0:
        return None
    zi = pow(z, -1, P)
    return (fq_dec(a[:12]) * zi * zi % P, fq_dec(a[12:24]) * zi * zi * zi % P)


def g2_jac_enc(pt):
    if pt is None:
        return np.concatenate([fq2_enc((1, 0)), fq2_enc((1, 0)), fq2_enc((0, 0))])
    return np.concatenate([fq2_enc(pt[0]), fq2_enc(pt[1]), fq2_enc((1, 0))])


def g2_jac_dec(a):
    """Only Z in {0, 1} (what the library returns)."""
    a = np.asarray(a)
    z = fq2_dec(a[48:72])
    if z == (0, 0):
        return None
    assert z == (1, 0)
    return (fq2_dec(a[:24]), fq2_dec(a[24:48]))


def gt_enc(f):
    return np.concatenate([fq2_enc(f[k]) for k in _TOWER_ORDER])


def gt_dec(a):
    a = np.asarray(a).reshape(6, 24)
    out = [None] * 6
    for slot, k in enumerate(_TOWER_ORDER):
        out[k] = fq2_dec(a[slot])
    return tuple(out)


def scalar_words(s, n=8):
    """Canonical (non-Montgomery) little-endian words of an integer scalar."""
    return _words(s, n)


# ---- arkworks `serialize_uncompressed` of statement values (what the verifier entry points take) ----
def ser_fr(v):
    return (v % R).to_bytes(32, "little")


def ser_gt(f):
    """Fq12 in struct order c0.c0 .. c1.c2, each Fq little-endian canonical (PairingOutput)."""
    return b"".join((c % P).to_bytes(48, "little") for k in _TOWER_ORDER for c in f[k])


def ser_g1(pt):
    """ark-bls12-381 uncompressed G1: big-endian x || y; identity = 0x40 then zeros."""
    if pt is None:
        return b"\x40" + bytes(95)
    return pt[0].to_bytes(48, "big") + pt[1].to_bytes(48, "big")


def ser_g2(pt):
    if pt is None:
        return b"\x40" + bytes(191)
    (x0, x1), (y0, y1) = pt
    return b"".join(v.to_bytes(48, "big") for v in (x1, x0, y1, y0))


def ser_identity_output(item_bytes):
    """IdentityOutput<T>(vec![t]) (dh_commitments/src/identity/mod.rs:33-62): u64 LE length, then the item."""
    return (1).to_bytes(8, "little") + item_bytes


def ser_value(v):
    """Serialise a commitment / inner-product value by its Python shape: int = Fr, 6-tuple = GT,
    None or a pair of ints = G1, a pair of pairs = G2."""
    if isinstance(v, int):
        return ser_fr(v)
    if v is not None and len(v) == 6:
        return ser_gt(v)
    if v is None:
        raise ValueError("ambiguous identity point: serialise with ser_g1 / ser_g2")
    return ser_g1(v) if isinstance(v[0], int) else ser_g2(v)
