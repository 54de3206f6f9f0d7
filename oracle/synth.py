"""ORACLE (test infrastructure, NOT product code) -- deterministic synthetic inputs.

scalar(tag, i, seed) = low 31 bytes (LE) of Blake2b-512("ripp-b200/" || tag || LE64(seed) || LE64(i)),
always < r; points are scalar * standard generator.  The product generates the same inputs on
the GPU (ripp_b200.synth); this copy is the checker's.
"""
import struct

from . import bls12_381 as E
from .encoding import blake2b


def scalar(tag, i, seed=0):
    h = blake2b(b"ripp-b200/" + tag.encode() + struct.pack("<QQ", seed, i))
    return int.from_bytes(h[:31], "little")


def scalars(tag, n, seed=0):
    return [scalar(tag, i, seed) for i in range(n)]


def g1_points(tag, n, seed=0):
    return [E.g1_mul(E.G1_GEN, s) for s in scalars(tag, n, seed)]


def g2_points(tag, n, seed=0):
    return [E.g2_mul(E.G2_GEN, s) for s in scalars(tag, n, seed)]


def groth16_instance_scalars(n, num_inputs=5, seed=0):
    """Trapdoor-simulated Groth16 proofs as exponents.

    Returns (vk_scalars, proof_scalars[(a, b, c)], public_inputs) such that
    e(A,B) = e(alpha,beta) e(sum_j x_j IC_j, gamma) e(C, delta)
    (the PPE /root/reference/ip_proofs/src/applications/groth16_aggregation.rs:208-228 aggregates).
    """
    R = E.R
    vk = {k: scalar("vk-" + k, 0, seed) for k in ("alpha", "beta", "gamma", "delta")}
    ic = scalars("vk-ic", num_inputs + 1, seed)
    proofs, inputs = [], []
    dinv = pow(vk["delta"], -1, R)
    for i in range(n):
        x = [scalar("pi-%d" % j, i, seed) for j in range(num_inputs - 1)]
        x.append((scalar("pi-w", i, seed) + sum(x)) % R)
        a, b = scalar("proof-a", i, seed), scalar("proof-b", i, seed)
        icx = (ic[0] + sum(xj * icj for xj, icj in zip(x, ic[1:]))) % R
        c = (a * b - vk["alpha"] * vk["beta"] - vk["gamma"] * icx) * dinv % R
        proofs.append((a, b, c))
        inputs.append(x)
    return {"vk": vk, "ic": ic}, proofs, inputs


def groth16_instance(n, num_inputs=5, seed=0):
    sc, proofs, inputs = groth16_instance_scalars(n, num_inputs, seed)
    vk = {
        "alpha_g1": E.g1_mul(E.G1_GEN, sc["vk"]["alpha"]),
        "beta_g2": E.g2_mul(E.G2_GEN, sc["vk"]["beta"]),
        "gamma_g2": E.g2_mul(E.G2_GEN, sc["vk"]["gamma"]),
        "delta_g2": E.g2_mul(E.G2_GEN, sc["vk"]["delta"]),
        "gamma_abc_g1": [E.g1_mul(E.G1_GEN, s) for s in sc["ic"]],
    }
    pts = [(E.g1_mul(E.G1_GEN, a), E.g2_mul(E.G2_GEN, b), E.g1_mul(E.G1_GEN, c)) for a, b, c in proofs]
    return vk, pts, inputs


# ---- adversarial inputs for the verifiers: ON the curve but OUTSIDE the prime-order subgroup ---------------------
def _fp_sqrt(a):
    r = pow(a, (E.P + 1) // 4, E.P)
    return r if r * r % E.P == a % E.P else None


def _fp2_sqrt(a):
    a0, a1 = a
    if a1 == 0:
        s = _fp_sqrt(a0)
        if s is not None:
            return (s, 0)
        s = _fp_sqrt(-a0 % E.P)
        return None if s is None else (0, s)
    alpha = _fp_sqrt((a0 * a0 + a1 * a1) % E.P)
    if alpha is None:
        return None
    inv2 = pow(2, -1, E.P)
    for delta in ((a0 + alpha) * inv2 % E.P, (a0 - alpha) * inv2 % E.P):
        x0 = _fp_sqrt(delta)
        if x0:
            r = (x0, a1 * pow(2 * x0, -1, E.P) % E.P)
            if E.f2_sqr(r) == (a0 % E.P, a1 % E.P):
                return r
    return None


def g1_point_off_subgroup(seed=0):
    """A point of E(Fp): y^2 = x^3 + 4 that is not in G1 (the cofactor is ~2^126, so a random point almost never is)."""
    i = 0
    while True:
        x = scalar("off-g1", i, seed) * scalar("off-g1'", i, seed) % E.P
        y = _fp_sqrt((x**3 + 4) % E.P)
        if y is not None and not E.g1_in_subgroup((x, y)):
            return (x, y)
        i += 1


def g2_point_off_subgroup(seed=0):
    i = 0
    while True:
        x = (scalar("off-g2", i, seed) * scalar("off-g2'", i, seed) % E.P, scalar("off-g2''", i, seed))
        y = _fp2_sqrt(E.f2_add(E.f2_mul(E.f2_sqr(x), x), E.B2))
        if y is not None and not E.g2_in_subgroup((x, y)):
            return (x, y)
        i += 1


def gt_cyclotomic_off_subgroup(seed=0):
    """An element of the cyclotomic subgroup of Fq12* (passes the cheap half of the GT test) whose order is not r."""
    f = tuple((scalar("off-gt", 2 * k, seed) ** 2 % E.P, scalar("off-gt", 2 * k + 1, seed) ** 2 % E.P) for k in range(6))
    c = E.f12_mul(E.f12_conj(f), E.f12_inv(f))
    c = E.f12_mul(E.f12_frob(c, 2), c)
    assert not E.gt_in_subgroup(c)
    return c
