"""Host-side mirror of `sipp/src/lib.rs` for the reference's own instantiation `SIPP<Bls12_377, Blake2s>`
(sipp/src/lib.rs:228-254) over the C ABI (ripp377_*).  Values are Python ints / affine tuples; limbs are Montgomery
(R = 2^384 for the 377-bit Fq, 2^256 for the 253-bit Fr), 12 / 8 little-endian 32-bit words."""
import numpy as np

from .inner_products import default_context

X = 0x8508C00000000001
R = X**4 - X**2 + 1
P = (X - 1) ** 2 * R // 3 + X
_RQ, _RR = 1 << 384, 1 << 256
_RQI = pow(_RQ, -1, P)
_TOWER_ORDER = (0, 2, 4, 1, 3, 5)  # arkworks' c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2 = flat coefficients 0, 2, 4, 1, 3, 5


def _w(v, n):
    return np.frombuffer(int(v).to_bytes(4 * n, "little"), dtype=np.uint32)


def fq_enc(v):
    return _w(v % P * _RQ % P, 12)


def fq_dec(a):
    return int.from_bytes(np.ascontiguousarray(a, dtype=np.uint32).tobytes(), "little") * _RQI % P


def fr_vec_enc(vals):
    return np.stack([_w(v % R * _RR % R, 8) for v in vals]) if len(vals) else np.zeros((0, 8), dtype=np.uint32)


def g1_vec_enc(pts):
    if not len(pts):
        return np.zeros((0, 24), dtype=np.uint32)
    return np.stack([np.zeros(24, dtype=np.uint32) if p is None else np.concatenate([fq_enc(p[0]), fq_enc(p[1])]) for p in pts])


def g2_vec_enc(pts):
    if not len(pts):
        return np.zeros((0, 48), dtype=np.uint32)
    return np.stack([np.zeros(48, dtype=np.uint32) if p is None else
                     np.concatenate([fq_enc(p[0][0]), fq_enc(p[0][1]), fq_enc(p[1][0]), fq_enc(p[1][1])]) for p in pts])


def gt_enc(f):
    return np.ascontiguousarray(np.concatenate([np.concatenate([fq_enc(f[k][0]), fq_enc(f[k][1])]) for k in _TOWER_ORDER]))


def gt_dec(a):
    a = np.asarray(a).reshape(6, 24)
    out = [None] * 6
    for slot, k in enumerate(_TOWER_ORDER):
        out[k] = (fq_dec(a[slot][:12]), fq_dec(a[slot][12:]))
    return tuple(out)


def _enc(a, b, r):
    return np.ascontiguousarray(g1_vec_enc(a)), np.ascontiguousarray(g2_vec_enc(b)), np.ascontiguousarray(fr_vec_enc(r))


def pairing_inner_product(a, b, ctx=None):
    """PairingInnerProduct<Bls12_377>::inner_product (inner_products/src/lib.rs:52-74) on affine inputs."""
    ctx = ctx or default_context()
    return gt_dec(ctx.pairing_ip_affine_377(np.ascontiguousarray(g1_vec_enc(a)), np.ascontiguousarray(g2_vec_enc(b))))


def product_of_pairings_with_coeffs(a, b, r, ctx=None):
    """sipp/src/lib.rs:184-217."""
    ctx = ctx or default_context()
    return gt_dec(ctx.sipp_product_with_coeffs_377(*_enc(a, b, r)))


class SIPP377:
    @staticmethod
    def prove(a, b, r, value, ctx=None):
        """sipp/src/lib.rs:42-106 -> serialised Proof::gt_elems."""
        assert len(a) == len(b) and bin(len(a)).count("1") == 1
        ctx = ctx or default_context()
        return ctx.sipp_prove_377(*_enc(a, b, r), gt_enc(value))

    @staticmethod
    def verify(a, b, r, claimed_value, proof, ctx=None):
        """sipp/src/lib.rs:109-180 -> bool."""
        assert len(a) == len(b) and len(a) >= 2 and bin(len(a)).count("1") == 1
        ctx = ctx or default_context()
        return ctx.sipp_verify_377(*_enc(a, b, r), gt_enc(claimed_value), proof)
