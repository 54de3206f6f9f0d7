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
