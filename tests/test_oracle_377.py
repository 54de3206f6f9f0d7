"""Pins of the BLS12-377 oracle (oracle/bls12_377.py, oracle/sipp_377.py) by derivation and identities: the curve is
DERIVED from the BLS12 family polynomials, so primality, curve orders, the twist and the pairing's bilinearity are checks,
not recollections.  SIPP on the reference's own instantiation (sipp/src/lib.rs:228-254) round-trips and rejects tampering."""
import random

from oracle import bls12_377 as E
from oracle import sipp_377 as S

rnd = random.Random(377)


def _probable_prime(n):
    return all(pow(a, n - 1, n) == 1 for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37))


def test_parameters_follow_from_x():
    assert E.P.bit_length() == 377 and E.R.bit_length() == 253
    assert _probable_prime(E.P) and _probable_prime(E.R)
    assert (E.P + 1 - (E.X + 1)) % E.R == 0  # #E(Fp) = p + 1 - t, t = x + 1, divisible by r
    assert (E.P**12 - 1) % E.R == 0 and all((E.P**k - 1) % E.R for k in (1, 2, 3, 4, 6))  # embedding degree 12
    assert pow(E.P - E.BETA, (E.P - 1) // 2, E.P) == E.P - 1  # -5 is a non-residue: Fq2 is a field
    assert (E.X**2 - 1) ** 2 % E.R == (-(E.X**2 - 1) - 1) % E.R  # lambda = x^2 - 1 is a cube root of unity mod r


def test_generators_and_twist():
    assert E.g1_is_on_curve(E.G1_GEN) and E._mul_raw(E.G1_GEN, E.R, E.g1_add) is None
    assert E.g2_is_on_curve(E.G2_GEN) and E._mul_raw(E.G2_GEN, E.R, E.g2_add) is None
    assert E.f2_mul(E.B2, E.XI) == E.F2_ONE  # D-type twist: b' = b / xi
    # the derived G1 generator is the one arkworks uses for BLS12-377 (x coordinate recalled from ark-bls12-377)
    assert E.G1_GEN[0] == 0x008848DEFE740A67C8FC6225BF87FF5485951E2CAA9D41BB188282C8BD37CB5CD5481512FFCD394EEAB9B16EB21BE9EF
    assert E.G1_GEN[1] == 0x01914A69C5102EFF1F674F5D30AFEEC4BD7FB348CA3E52D96D182AD44FB82305C2FE3D3634A9591AFD82DE55559C8EA6


def test_pairing_bilinear_nondegenerate_order_r():
    a, b = rnd.randrange(1, E.R), rnd.randrange(1, E.R)
    e0 = E.pairing(E.G1_GEN, E.G2_GEN)
    assert e0 != E.F12_ONE and E.f12_pow(e0, E.R) == E.F12_ONE
    assert E.pairing(E.g1_mul(E.G1_GEN, a), E.g2_mul(E.G2_GEN, b)) == E.gt_pow(e0, a * b % E.R)
    f = E.miller_loop(E.G1_GEN, E.G2_GEN)
    assert E.final_exponentiation(f) == E.final_exponentiation_naive(f)
    assert E.pairing(None, E.G2_GEN) == E.F12_ONE


def test_serialisation_flags():
    p = E.g1_mul(E.G1_GEN, 5)
    s, sn = E.ser_g1(p), E.ser_g1(E.g1_neg(p))
    assert len(s) == 96 and len(sn) == 96
    assert (s[95] & 0x80) != (sn[95] & 0x80) and s[:48] == sn[:48]  # the sign flag separates y from -y
    assert E.ser_g1(None)[95] == 0x40 and not any(E.ser_g1(None)[:95])
    q = E.g2_mul(E.G2_GEN, 7)
    t, tn = E.ser_g2(q), E.ser_g2(E.g2_neg(q))
    assert len(t) == 192 and (t[191] & 0x80) != (tn[191] & 0x80)
    assert len(E.ser_gt(E.pairing(p, q))) == 576


def test_sipp_round_trip_and_tamper():
    """sipp/src/lib.rs:233-254 at n = 8 (the reference's test uses 32 with the same structure)."""
    n = 8
    a, b = S.points("s377-a", n, 1), S.points("s377-b", n, 2)
    r = [rnd.randrange(E.R) for _ in range(n)]
    z = S.product_of_pairings_with_coeffs(a, b, r)
    proof = S.sipp_prove(a, b, r, z)
    assert len(proof) == 3 and S.sipp_verify(a, b, r, z, proof)
    assert not S.sipp_verify(a, b, r, E.gt_mul(z, z), proof)
    bad = list(proof)
    bad[1] = (bad[1][1], bad[1][0])
    assert not S.sipp_verify(a, b, r, z, bad)


# ---- the BLS12-377 device headers compiled for the host (tests/hostsim/hostsim377.cpp) against this oracle ----------
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hs377():
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim377.cpp")
    so = os.path.join(ROOT, "tests", "hostsim", "_hostsim377.so")
    csrc = os.path.join(ROOT, "ripp_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    return ctypes.CDLL(so)


def _call(lib, fn, *args, out):
    o = np.zeros(out, dtype=np.uint32)
    keep = [np.ascontiguousarray(a, dtype=np.uint32) if isinstance(a, np.ndarray) else a for a in args]
    cargs = [ctypes.c_void_p(a.ctypes.data) if isinstance(a, np.ndarray) else a for a in keep]
    getattr(lib, fn)(*cargs, ctypes.c_void_p(o.ctypes.data))
    return o


def test_device_headers_on_host_match_oracle(hs377):
    """Miller loop on the D-type twist + final exponentiation, the Granger-Scott squaring, both scalar multiplications and
    the Fr inversion of bls377.cuh, executed on the CPU with the emulated carry flag, equal the big-int oracle."""
    from ripp_b200 import sipp_377 as G

    p, q = E.g1_mul(E.G1_GEN, 12345), E.g2_mul(E.G2_GEN, 6789)
    ep, eq = G.g1_vec_enc([p])[0], G.g2_vec_enc([q])[0]
    assert G.gt_dec(_call(hs377, "hs377_pairing", ep, eq, 1, out=144)) == E.pairing(p, q)
    m = G.gt_dec(_call(hs377, "hs377_pairing", ep, eq, 0, out=144))
    assert E.final_exponentiation(m) == E.pairing(p, q)  # Miller values may differ by subfield factors; the pairing may not
    s = rnd.randrange(E.R)
    sw = np.frombuffer(s.to_bytes(32, "little"), dtype=np.uint32)
    got = _call(hs377, "hs377_g1_mul", ep, sw, out=24)
    assert (G.fq_dec(got[:12]), G.fq_dec(got[12:])) == E.g1_mul(p, s)
    got = _call(hs377, "hs377_g2_mul", eq, sw, out=48)
    assert ((G.fq_dec(got[:12]), G.fq_dec(got[12:24])), (G.fq_dec(got[24:36]), G.fq_dec(got[36:]))) == E.g2_mul(q, s)
    z = E.pairing(p, q)  # in the cyclotomic subgroup
    both = _call(hs377, "hs377_cyc_sqr", G.gt_enc(z), out=288)
    assert G.gt_dec(both[:144]) == G.gt_dec(both[144:]) == E.f12_sqr(z)
    a = rnd.randrange(1, E.R)
    inv = _call(hs377, "hs377_fr_inv", np.frombuffer((a * (1 << 256) % E.R).to_bytes(32, "little"), dtype=np.uint32), out=8)
    assert int.from_bytes(inv.tobytes(), "little") * pow(1 << 256, -1, E.R) % E.R == pow(a, -1, E.R)
