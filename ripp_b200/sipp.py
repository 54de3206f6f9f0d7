"""Host-side mirror of `sipp/src/lib.rs` (SIPP<E, D> with E = BLS12-381, D = Blake2s) over the C ABI."""
import numpy as np

from . import codec
from .inner_products import default_context


def _enc(a, b, r):
    n = len(a)
    return (codec.g1_vec_enc(a).reshape(n, 24), codec.g2_vec_enc(b).reshape(n, 48), codec.fr_vec_enc(r).reshape(n, 8))


def product_of_pairings_with_coeffs(a, b, r, ctx=None):
    """sipp/src/lib.rs:184-217."""
    ctx = ctx or default_context()
    return codec.gt_dec(ctx.sipp_product_with_coeffs(*_enc(a, b, r)))


class SIPP:
    @staticmethod
    def prove(a, b, r, value, ctx=None):
        """sipp/src/lib.rs:42-106 -> serialised Proof::gt_elems (log2 n pairs of GT)."""
        assert len(a) == len(b) and bin(len(a)).count("1") == 1  # lib.rs:47-53
        ctx = ctx or default_context()
        ea, eb, er = _enc(a, b, r)
        return ctx.sipp_prove(ea, eb, er, np.ascontiguousarray(codec.gt_enc(value)))

    @staticmethod
    def verify(a, b, r, claimed_value, proof, ctx=None):
        """sipp/src/lib.rs:109-180 -> bool."""
        assert len(a) == len(b) and len(a) >= 2 and bin(len(a)).count("1") == 1  # lib.rs:117-120
        ctx = ctx or default_context()
        ea, eb, er = _enc(a, b, r)
        return ctx.sipp_verify(ea, eb, er, np.ascontiguousarray(codec.gt_enc(claimed_value)), proof)
