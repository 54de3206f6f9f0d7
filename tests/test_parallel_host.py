"""Host-side logic of the sharded provers and of the Python mirror (no GPU): the product's own serialisers and
Fiat-Shamir challenges (ripp_b200/codec.py, parallel.py) against the oracle's independent restatement, and the cyclic
partition property the sharded GIPA relies on."""
import random

from oracle import bls12_381 as E
from oracle import encoding as S
from oracle import protocols as O
from oracle import synth as OS
from ripp_b200 import codec as C
from ripp_b200 import parallel as P

rnd = random.Random(7)


def test_serialisers_match_oracle():
    g1, g2 = OS.g1_points("ph-a", 2), OS.g2_points("ph-b", 2)
    gt = tuple((rnd.randrange(E.P), rnd.randrange(E.P)) for _ in range(6))
    s = rnd.randrange(E.R)
    assert C.ser_fr(s) == S.ser_fr(s)
    assert C.ser_gt(gt) == S.ser_gt(gt)
    assert [C.ser_g1(p) for p in g1 + [None]] == [S.ser_g1(p) for p in g1 + [None]]
    assert [C.ser_g2(p) for p in g2 + [None]] == [S.ser_g2(p) for p in g2 + [None]]
    assert C.ser_identity_output(C.ser_gt(gt)) == S.ser_vec([gt], S.ser_gt)
    assert C.ser_value(s) == S.ser_fr(s) and C.ser_value(gt) == S.ser_gt(gt)
    assert C.ser_value(g1[0]) == S.ser_g1(g1[0]) and C.ser_value(g2[0]) == S.ser_g2(g2[0])
    # codec round trips through the ABI's Montgomery words
    assert C.gt_dec(C.gt_enc(gt)) == gt and C.g2_dec(C.g2_enc(g2[1])) == g2[1] and C.fr_dec(C.fr_enc(s)) == s


def test_gipa_challenge_matches_oracle():
    gt = lambda: tuple((rnd.randrange(E.P), rnd.randrange(E.P)) for _ in range(6))
    g = O.GIPA(O.PairingInnerProduct, O.AFGHOCommitmentG1, O.AFGHOCommitmentG2, O.IdentityCommitment(O.GTT))
    for prev in (0, rnd.randrange(E.R)):
        com_1, com_2 = (gt(), gt(), [gt()]), (gt(), gt(), [gt()])
        b = b""
        for com in (com_1, com_2):
            b += C.ser_gt(com[0]) + C.ser_gt(com[1]) + C.ser_identity_output(C.ser_gt(com[2][0]))
        assert P.gipa_challenge(prev, b) == g._challenge(prev, com_1, com_2)
    parts = bytes(rnd.randrange(256) for _ in range(300))
    assert P.challenge_from_random_bytes(parts) == O._kzg_challenge(S.blake2b, parts)


def test_out_types():
    assert P.ip_out_type("G1", "G2") == "GT" and P.ip_out_type("G2", "G1") == "GT"
    assert P.ip_out_type("G1", "Fr") == "G1" and P.ip_out_type("Fr", "G2") == "G2"
    assert P.ip_out_type("Fr", "Fr") == "Fr" and P.ip_out_type(None, "Fr") == "Fr"


def test_cyclic_partition_keeps_rounds_local():
    """While g divides n', index i and its partner i + n' live on the same rank, at local indices j and j + n'/g, and
    the folded element i stays on rank i mod g (gipa.rs:209-217, 261-291)."""
    for g in (1, 2, 4, 8):
        for n in (8, 64):
            glob = list(range(n))
            shares = [P.cyclic_share(glob, k, g) for k in range(g)]
            assert sorted(sum(shares, [])) == glob
            length = n
            while length // 2 >= g:
                half = length // 2
                for k in range(g):
                    loc = shares[k][: length // g]
                    lo, hi = loc[: half // g], loc[half // g : 2 * (half // g)]
                    assert all(h == l + half for l, h in zip(lo, hi))       # partners are co-located
                    assert all(l % g == k for l in lo)                       # and the fold result stays on rank k
                length = half
