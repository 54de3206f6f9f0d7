"""Host-side mirror of `dh_commitments/src` (trait DoublyHomomorphicCommitment, lib.rs:20-55)."""
from .inner_products import (
    MultiexponentiationInnerProductG1,
    MultiexponentiationInnerProductG2,
    PairingInnerProduct,
)


class _Commitment:
    @classmethod
    def verify(cls, k, m, com, ctx=None):
        """dh_commitments/src/lib.rs:52-54: default verify = (commit == com)."""
        return cls.commit(k, m, ctx) == com


class AFGHOCommitmentG1(_Commitment):
    """afgho16/mod.rs:20-33: message in G1, key in G2, commit = IP(m, k)."""

    @staticmethod
    def commit(k, m, ctx=None):
        return PairingInnerProduct.inner_product(m, k, ctx)


class AFGHOCommitmentG2(_Commitment):
    """afgho16/mod.rs:35-48: message in G2, key in G1, commit = IP(k, m)."""

    @staticmethod
    def commit(k, m, ctx=None):
        return PairingInnerProduct.inner_product(k, m, ctx)


class PedersenCommitmentG1(_Commitment):
    """pedersen/mod.rs:14-27 with G = G1: commit = MSM(keys, msgs)."""

    @staticmethod
    def commit(k, m, ctx=None):
        return MultiexponentiationInnerProductG1.inner_product(k, m, ctx)


class PedersenCommitmentG2(_Commitment):
    @staticmethod
    def commit(k, m, ctx=None):
        return MultiexponentiationInnerProductG2.inner_product(k, m, ctx)
