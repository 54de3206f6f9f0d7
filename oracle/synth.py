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
