"""ORACLE (test infrastructure, NOT product code) -- byte encodings and hashes.

PARITY UNPINNED (see oracle/bls12_381.py).  Restates ark-serialize 0.4
`serialize_uncompressed` for the types that feed RIPP's Fiat-Shamir hashes
(SURVEY.md App. A-4/A-5), `Fp::from_random_bytes`, and the SIPP hash-reseeded ChaCha20 RNG
(/root/reference/sipp/src/rng.rs:12-73).
"""
import hashlib
import struct

from . import bls12_381 as E


def ser_fr(a):
    return (a % E.R).to_bytes(32, "little")


def ser_fq(a):
    return (a % E.P).to_bytes(48, "little")


def ser_fq2(a):
    return ser_fq(a[0]) + ser_fq(a[1])


def ser_gt(a):
    """Fq12 = c0 || c1, c_i = c_i0 || c_i1 || c_i2; c_i.c_j is the w^(2j+i) coefficient."""
    return b"".join(ser_fq2(a[2 * j + i]) for i in range(2) for j in range(3))


def ser_g1(pt):
    """ark-bls12-381 0.4 uncompressed G1: big-endian x || y, flags in the top bits of byte 0."""
    if pt is None:
        return b"\x40" + b"\x00" * 95
    return pt[0].to_bytes(48, "big") + pt[1].to_bytes(48, "big")


def ser_g2(pt):
    if pt is None:
        return b"\x40" + b"\x00" * 191
    (x0, x1), (y0, y1) = pt
    return b"".join(v.to_bytes(48, "big") for v in (x1, x0, y1, y0))


def ser_vec(items, ser):
    return struct.pack("<Q", len(items)) + b"".join(ser(i) for i in items)


def blake2b(data):
    return hashlib.blake2b(data, digest_size=64).digest()


def blake2s(data):
    return hashlib.blake2s(data, digest_size=32).digest()


def fr_from_random_bytes(digest):
    """ark-ff 0.4 Fp::from_random_bytes: first 32 bytes LE, top bit cleared, None if >= r."""
    v = int.from_bytes(digest[:32], "little") & ((1 << 255) - 1)
    return v if v < E.R else None


def challenge_u128(digest):
    """gipa.rs:248-251: u128::from_be_bytes(digest[0..16]) embedded in Fr."""
    return int.from_bytes(digest[:16], "big")


# ----------------------------------------------------------------------------- ChaCha20 (rand_chacha 0.3 ChaChaRng)
def _rotl(v, n):
    return ((v << n) & 0xFFFFFFFF) | (v >> (32 - n))


def _qr(s, a, b, c, d):
    s[a] = (s[a] + s[b]) & 0xFFFFFFFF
    s[d] = _rotl(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & 0xFFFFFFFF
    s[b] = _rotl(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b]) & 0xFFFFFFFF
    s[d] = _rotl(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & 0xFFFFFFFF
    s[b] = _rotl(s[b] ^ s[c], 7)


def chacha20_block(key32, counter=0, nonce_words=(0, 0)):
    init = list(struct.unpack("<4I", b"expand 32-byte k")) + list(struct.unpack("<8I", key32))
    init += [counter & 0xFFFFFFFF, counter >> 32, nonce_words[0], nonce_words[1]]
    s = list(init)
    for _ in range(10):
        _qr(s, 0, 4, 8, 12)
        _qr(s, 1, 5, 9, 13)
        _qr(s, 2, 6, 10, 14)
        _qr(s, 3, 7, 11, 15)
        _qr(s, 0, 5, 10, 15)
        _qr(s, 1, 6, 11, 12)
        _qr(s, 2, 7, 8, 13)
        _qr(s, 3, 4, 9, 14)
    return struct.pack("<16I", *[(a + b) & 0xFFFFFFFF for a, b in zip(s, init)])


class FiatShamirRng:
    """sipp/src/rng.rs: seed = H(material); absorb: seed = H(new || seed); ChaCha20 re-keyed each time."""

    def __init__(self, seed_material, digest=blake2s):
        self.digest = digest
        self.seed = digest(seed_material)
        self.block = 0
        self.buf = b""

    def absorb(self, data):
        self.seed = self.digest(data + self.seed)
        self.block = 0
        self.buf = b""

    def next_bytes(self, n):
        while len(self.buf) < n:
            self.buf += chacha20_block(self.seed, self.block)
            self.block += 1
        out, self.buf = self.buf[:n], self.buf[n:]
        return out

    def next_u128(self):
        """rand 0.8 Standard for u128: low = next_u64(), high = next_u64()."""
        return int.from_bytes(self.next_bytes(16), "little")
