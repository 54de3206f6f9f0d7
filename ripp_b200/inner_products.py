"""Host-side mirror of `inner_products/src/lib.rs` (trait InnerProduct, lib.rs:40-49) over the
C ABI.  Same names and error behaviour as the reference; values are Python ints / affine tuples
(tests' convention, see codec.py); all arithmetic runs on the GPU."""
import numpy as np

from . import _lib, codec

_ctx = None


def default_context():
    """Process-wide context on LOCAL_RANK's GPU (the Rust shim's `lazy_static` context)."""
    global _ctx
    if _ctx is None:
        import os

        _ctx = _lib.Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _ctx


class PairingInnerProduct:
    """inner_products/src/lib.rs:51-74."""

    @staticmethod
    def inner_product(left, right, ctx=None):
        ctx = ctx or default_context()
        g1 = codec.g1_vec_enc(left).reshape(len(left), 24)
        g2 = codec.g2_vec_enc(right).reshape(len(right), 48)
        return codec.gt_dec(ctx.pairing_ip_affine(g1, g2))


class MultiexponentiationInnerProductG1:
    """inner_products/src/lib.rs:118-142 with G = G1: sum_i right[i] * left[i]."""

    @staticmethod
    def inner_product(left, right, ctx=None):
        ctx = ctx or default_context()
        bases = np.stack([codec.g1_jac_enc(p) for p in left]) if len(left) else np.zeros((0, 36), dtype=np.uint32)
        return codec.g1_jac_dec(ctx.msm_g1(bases, codec.fr_vec_enc(right)))


class MultiexponentiationInnerProductG2:
    """inner_products/src/lib.rs:118-142 with G = G2."""

    @staticmethod
    def inner_product(left, right, ctx=None):
        ctx = ctx or default_context()
        bases = np.stack([codec.g2_jac_enc(p) for p in left]) if len(left) else np.zeros((0, 72), dtype=np.uint32)
        return codec.g2_jac_dec(ctx.msm_g2(bases, codec.fr_vec_enc(right)))


class ScalarInnerProduct:
    """inner_products/src/lib.rs:144-166."""

    @staticmethod
    def inner_product(left, right, ctx=None):
        ctx = ctx or default_context()
        return codec.fr_dec(ctx.scalar_ip(codec.fr_vec_enc(left), codec.fr_vec_enc(right)))
