"""Host-side mirror of `dh_commitments/src` (trait DoublyHomomorphicCommitment, lib.rs:20-55)."""
from .inner_products import (
    MultiexponentiationInnerProductG1,
    MultiexponentiationInnerProductG2,
    PairingInnerProduct,
)


def generators_from_scalars(scalars, group, ctx=None):
    """The key vector [s_i * g] behind `setup` (dh_commitments/src/lib.rs:59-61 random_generators draws G::rand(rng);
    here the caller draws the exponents -- SURVEY.md §8d synthetic keys -- and the fixed-base table does the rest)."""
    from . import codec
    from .inner_products import default_context

    ctx = ctx or default_context()
    n = len(scalars)
    out = ctx.alloc(max(n, 1) * (96 if group == 1 else 192))
    ctx.fixed_base_msm_dev(group, None, ctx.to_device(codec.fr_vec_enc(scalars)) if n else None, n, out)
    ctx.sync()
    return (codec.g1_vec_dec if group == 1 else codec.g2_vec_dec)(out.download((n, 24 if group == 1 else 48))) if n else []


class _Commitment:
    KEY_GROUP = None

    @classmethod
    def setup(cls, scalars, ctx=None):
        """trait fn setup (dh_commitments/src/lib.rs:48): `size` keys; the random exponents come from the caller."""
        return generators_from_scalars(scalars, cls.KEY_GROUP, ctx)

    @classmethod
    def verify(cls, k, m, com, ctx=None):
        """dh_commitments/src/lib.rs:52-54: default verify = (commit == com)."""
        return cls.commit(k, m, ctx) == com


class AFGHOCommitmentG1(_Commitment):
    """afgho16/mod.rs:20-33: message in G1, key in G2, commit = IP(m, k)."""
    KEY_GROUP = 2

    @staticmethod
    def commit(k, m, ctx=None):
        return PairingInnerProduct.inner_product(m, k, ctx)


class AFGHOCommitmentG2(_Commitment):
    """afgho16/mod.rs:35-48: message in G2, key in G1, commit = IP(k, m)."""
    KEY_GROUP = 1

    @staticmethod
    def commit(k, m, ctx=None):
        return PairingInnerProduct.inner_product(k, m, ctx)


class PedersenCommitmentG1(_Commitment):
    """pedersen/mod.rs:14-27 with G = G1: commit = MSM(keys, msgs)."""
    KEY_GROUP = 1

    @staticmethod
    def commit(k, m, ctx=None):
        return MultiexponentiationInnerProductG1.inner_product(k, m, ctx)


class PedersenCommitmentG2(_Commitment):
    KEY_GROUP = 2
    @staticmethod
    def commit(k, m, ctx=None):
        return MultiexponentiationInnerProductG2.inner_product(k, m, ctx)
