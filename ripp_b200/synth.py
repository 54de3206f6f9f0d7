"""Deterministic synthetic inputs, product side: scalars from Blake2b on the host,
points = scalar * generator computed ON THE GPU (ripp_g{1,2}_scale_dev with points == NULL).

scalar(tag, i, seed) = low 31 bytes (LE) of Blake2b-512("ripp-b200/" || tag || LE64(seed) || LE64(i)).
"""
import hashlib
import struct

import numpy as np

from . import codec


def scalar(tag, i, seed=0):
    h = hashlib.blake2b(b"ripp-b200/" + tag.encode() + struct.pack("<QQ", seed, i), digest_size=64).digest()
    return int.from_bytes(h[:31], "little")


def scalars(tag, n, seed=0):
    return [scalar(tag, i, seed) for i in range(n)]


def scalars_mont(tag, n, seed=0):
    """(n, 8) uint32 Montgomery-form Fr words."""
    return codec.fr_vec_enc(scalars(tag, n, seed))


def g1_points_dev(ctx, tag, n, seed=0):
    sc = ctx.to_device(scalars_mont(tag, n, seed))
    out = ctx.alloc(n * 96)
    ctx.g1_scale_dev(None, sc, n, out)
    ctx.sync()
    sc.free()
    return out


def g2_points_dev(ctx, tag, n, seed=0):
    sc = ctx.to_device(scalars_mont(tag, n, seed))
    out = ctx.alloc(n * 192)
    ctx.g2_scale_dev(None, sc, n, out)
    ctx.sync()
    sc.free()
    return out


R = codec.R


def groth16_scalars(n, num_inputs=5, seed=0):
    """Trapdoor-simulated Groth16 instance as exponents: returns
    (vk scalars, ic scalars, [(a, b, c)], public inputs) with A = a g1, B = b g2, C = c g1 satisfying
    e(A,B) = e(alpha g1, beta g2) e(sum_j x_j IC_j, gamma g2) e(C, delta g2)."""
    vk = {k: scalar("vk-" + k, 0, seed) for k in ("alpha", "beta", "gamma", "delta")}
    ic = scalars("vk-ic", num_inputs + 1, seed)
    dinv = pow(vk["delta"], -1, R)
    proofs, inputs = [], []
    for i in range(n):
        x = [scalar("pi-%d" % j, i, seed) for j in range(num_inputs - 1)]
        x.append((scalar("pi-w", i, seed) + sum(x)) % R)
        a, b = scalar("proof-a", i, seed), scalar("proof-b", i, seed)
        icx = (ic[0] + sum(xj * icj for xj, icj in zip(x, ic[1:]))) % R
        proofs.append((a, b, (a * b - vk["alpha"] * vk["beta"] - vk["gamma"] * icx) * dinv % R))
        inputs.append(x)
    return vk, ic, proofs, inputs


def _gen_dev(ctx, group, exps):
    sc = ctx.to_device(codec.fr_vec_enc(exps))
    out = ctx.alloc(len(exps) * (96 if group == 1 else 192))
    (ctx.g1_scale_dev if group == 1 else ctx.g2_scale_dev)(None, sc, len(exps), out)
    ctx.sync()
    sc.free()
    return out


def tipp_instance_dev(ctx, n, seed=0):
    """SRS (tipa/mod.rs:150-164 with alpha, beta = scalar("srs-alpha"/"srs-beta")) and n simulated
    Groth16 proofs, all generated on the GPU.  -> dict(srs_g1, srs_g2, a, b, c) of device buffers."""
    alpha, beta = scalar("srs-alpha", 0, seed), scalar("srs-beta", 0, seed)
    m = 2 * n - 1
    pa, pb = [1], [1]
    for _ in range(m - 1):
        pa.append(pa[-1] * alpha % R)
        pb.append(pb[-1] * beta % R)
    vk, ic, proofs, inputs = groth16_scalars(n, seed=seed)
    # verifier material (tipa/mod.rs:120-127 VerifierSRS; ark-groth16 VerifyingKey), packed as include/ripp_b200.h
    # lays it out: g | h | g_beta | h_alpha and alpha_g1 | beta_g2 | gamma_g2 | delta_g2 | gamma_abc_g1[]
    g1s = _gen_dev(ctx, 1, [1, beta, vk["alpha"]] + ic)
    g2s = _gen_dev(ctx, 2, [1, alpha, vk["beta"], vk["gamma"], vk["delta"]])
    h1 = g1s.download((3 + len(ic), 24))
    h2 = g2s.download((5, 48))
    g1s.free()
    g2s.free()
    vsrs = np.concatenate([h1[0], h2[0], h1[1], h2[1]])
    vk_words = np.concatenate([h1[2], h2[2], h2[3], h2[4]] + [h1[3 + j] for j in range(len(ic))])
    return {
        "srs_g1": _gen_dev(ctx, 1, pa), "srs_g2": _gen_dev(ctx, 2, pb),
        "a": _gen_dev(ctx, 1, [p[0] for p in proofs]), "b": _gen_dev(ctx, 2, [p[1] for p in proofs]),
        "c": _gen_dev(ctx, 1, [p[2] for p in proofs]),
        "vsrs": np.ascontiguousarray(vsrs), "vk": np.ascontiguousarray(vk_words),
        "inputs": np.ascontiguousarray(np.stack([codec.fr_vec_enc(x) for x in inputs])),
    }
