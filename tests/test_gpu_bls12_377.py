"""BLS12-377 on the GPU (csrc/bls377.cuh, csrc/sipp377.cu) against oracle/bls12_377.py: the reference's own SIPP
instantiation `SIPP<Bls12_377, Blake2s>` (sipp/src/lib.rs:228-254) -- pairing products, product_of_pairings_with_coeffs,
proof BYTES, verifier accept / reject, error behaviour."""
import random

import pytest

from oracle import bls12_377 as E
from oracle import sipp_377 as S
from ripp_b200 import _lib
from ripp_b200 import sipp_377 as G

pytestmark = pytest.mark.gpu
rnd = random.Random(3770)


@pytest.mark.parametrize("n", [0, 1, 2, 5, 33])
def test_pairing_product_matches_oracle(ctx, n):
    a, b = S.points("g377-a", n, 1), S.points("g377-b", n, 2)
    if n >= 5:
        a[1], b[3] = None, None  # identities contribute 1 (ark-ec multi_miller_loop skips them)
    assert G.pairing_inner_product(a, b, ctx) == E.multi_pairing(a, b)


def test_pairing_length_mismatch(ctx):
    with pytest.raises(_lib.LengthMismatch):
        G.pairing_inner_product(S.points("g377-a", 3, 1), S.points("g377-b", 2, 2), ctx)


@pytest.mark.parametrize("n", [2, 8, 32])
def test_sipp_prove_bytes_and_verify(ctx, n):
    """sipp/src/lib.rs:233-254 (n = 32 there): same proof bytes as the oracle, both verifiers accept, tampering rejects."""
    a, b = S.points("s377-a", n, 1), S.points("s377-b", n, 2)
    r = [rnd.randrange(E.R) for _ in range(n)]
    if n >= 8:
        r[0], r[1] = 0, E.R - 1
    z = G.product_of_pairings_with_coeffs(a, b, r, ctx)
    assert z == S.product_of_pairings_with_coeffs(a, b, r)
    proof = G.SIPP377.prove(a, b, r, z, ctx)
    want = S.sipp_prove(a, b, r, z)
    assert proof == S.ser_proof(want)
    assert S.sipp_verify(a, b, r, z, want)
    assert G.SIPP377.verify(a, b, r, z, proof, ctx)
    assert not G.SIPP377.verify(a, b, r, E.gt_mul(z, z), proof, ctx)
    bad = bytearray(proof)
    bad[600] ^= 1
    try:
        ok = G.SIPP377.verify(a, b, r, z, bytes(bad), ctx)
    except _lib.RippError:
        ok = False  # non-canonical after the flip: what ark-serialize reports before verify runs
    assert not ok
