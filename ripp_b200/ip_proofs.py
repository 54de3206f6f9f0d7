"""Host-side mirror of `ip_proofs/src` (GIPA gipa.rs:16-22,97-178; TIPA tipa/mod.rs:32-39,150-231;
TIPAWithSSM structured_scalar_message.rs:130-268; aggregate_proofs groth16_aggregation.rs:77-160)
over the C ABI.  Values are Python ints / affine tuples; proofs are returned as the arkworks
`serialize_uncompressed` bytes the library emits."""
import numpy as np

from . import _lib, codec
from .inner_products import default_context


class InnerProductArgumentError(Exception):
    """ip_proofs/src/lib.rs:22-43."""


_ENC = {"G1": codec.g1_vec_enc, "G2": codec.g2_vec_enc, "Fr": codec.fr_vec_enc, None: None}
_KIND_TYPES = {
    _lib.GIPA_PAIRING: ("G1", "G2", "G2", "G1"),
    _lib.GIPA_MULTIEXP_PEDERSEN: ("G1", "Fr", "G2", "G1"),
    _lib.GIPA_MULTIEXP_SSM: ("G1", "Fr", "G2", None),
    _lib.GIPA_SCALAR_PEDERSEN_G2_G2: ("Fr", "Fr", "G2", "G2"),
    _lib.GIPA_SCALAR_PEDERSEN_G2_G1: ("Fr", "Fr", "G2", "G1"),
    _lib.GIPA_SCALAR_SSM: ("Fr", "Fr", "G2", None),
}


def _upload(ctx, kind, vecs):
    out = []
    for t, v in zip(_KIND_TYPES[kind], vecs):
        out.append(None if t is None else ctx.to_device(_ENC[t](v)))
    return out


class GIPA:
    """GIPA<IP, LMC, RMC, IPC, Blake2b> for one of the instantiations in include/ripp_b200.h."""

    def __init__(self, kind, ctx=None):
        self.kind, self.ctx = kind, ctx or default_context()

    def prove_with_aux(self, values, ck):
        """gipa.rs:162-178.  values = (m_a, m_b); ck = (ck_a, ck_b or None).  -> (proof bytes, transcript ints, ck_base bytes)"""
        m_a, m_b = values
        n = len(m_a)
        if n == 0 or n & (n - 1) or len(m_b) != n:
            raise InnerProductArgumentError("left length, right length: %d, %d" % (len(m_a), len(m_b)))
        a, b, v, w = _upload(self.ctx, self.kind, (m_a, m_b, ck[0], ck[1]))
        proof, tr, ckb = self.ctx.gipa_prove_dev(self.kind, a, b, v, w, n)
        return proof, codec.fr_vec_dec(tr), ckb


class TIPA:
    """TIPA<IP, LMC, RMC, IPC, Bls12_381, Blake2b> / TIPAWithSSM (for the *_SSM kinds)."""

    def __init__(self, kind, ctx=None):
        self.kind, self.ctx = kind, ctx or default_context()

    def prove_with_srs_shift(self, srs, values, ck, r_shift=1):
        """tipa/mod.rs:176-231.  srs = (g_alpha_powers, h_beta_powers)."""
        m_a, m_b = values
        n = len(m_a)
        if n == 0 or n & (n - 1) or len(m_b) != n:
            raise InnerProductArgumentError("left length, right length: %d, %d" % (len(m_a), len(m_b)))
        a, b, v, w = _upload(self.ctx, self.kind, (m_a, m_b, ck[0], ck[1]))
        s1 = self.ctx.to_device(codec.g1_vec_enc(srs[0]))
        s2 = self.ctx.to_device(codec.g2_vec_enc(srs[1]))
        return self.ctx.tipa_prove_dev(self.kind, s1, s2, a, b, v, w, n, codec.fr_enc(r_shift).copy())

    def prove(self, srs, values, ck):
        return self.prove_with_srs_shift(srs, values, ck, 1)


def aggregate_proofs(srs, proofs, ctx=None):
    """groth16_aggregation.rs:77-160.  srs = (g_alpha_powers, h_beta_powers); proofs = [(A, B, C)]."""
    ctx = ctx or default_context()
    n = len(proofs)
    a = ctx.to_device(codec.g1_vec_enc([p[0] for p in proofs]))
    b = ctx.to_device(codec.g2_vec_enc([p[1] for p in proofs]))
    c = ctx.to_device(codec.g1_vec_enc([p[2] for p in proofs]))
    s1 = ctx.to_device(codec.g1_vec_enc(srs[0]))
    s2 = ctx.to_device(codec.g2_vec_enc(srs[1]))
    return ctx.tipp_aggregate_dev(s1, s2, a, b, c, n)
