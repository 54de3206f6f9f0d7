"""ORACLE (test infrastructure, NOT product code) -- RIPP protocol layer restated in Python.

PARITY UNPINNED (see oracle/bls12_381.py): the reference's tests are seeded round trips
only; this restatement is pinned by prove->verify acceptance,
tamper->reject and the commitment homomorphism, not by golden bytes.

Each function cites the reference lines it follows (paths relative to /root/reference).
Heavy group work goes through a `Backend` so the same protocol code can run on the pure
big-int arithmetic (default) or on the compiled CPU oracle (oracle/cpu) at full sizes.
Points are affine tuples, `None` = identity; Fr elements are ints mod r.
"""
import struct

from . import bls12_381 as E
from .encoding import (
    FiatShamirRng,
    blake2b,
    blake2s,
    challenge_u128,
    fr_from_random_bytes,
    ser_fr,
    ser_g1,
    ser_g2,
    ser_gt,
    ser_vec,
)

R = E.R


class InnerProductError(Exception):
    """inner_products/src/lib.rs:18-38 MessageLengthInvalid / ip_proofs/src/lib.rs:22-25."""


# ----------------------------------------------------------------------------- backend
class PyBackend:
    """Pure big-int heavy ops."""

    def pairing_product(self, g1s, g2s):
        return E.multi_pairing(g1s, g2s)

    def msm_g1(self, pts, scalars):
        return E.msm(pts, scalars, E.g1_add, E.g1_mul)

    def msm_g2(self, pts, scalars):
        return E.msm(pts, scalars, E.g2_add, E.g2_mul)

    def mul_vec_g1(self, pts, scalars):
        return [E.g1_mul(p, s) for p, s in zip(pts, scalars)]

    def mul_vec_g2(self, pts, scalars):
        return [E.g2_mul(p, s) for p, s in zip(pts, scalars)]

    def fold_g1(self, hi, lo, c):
        return [E.g1_add(E.g1_mul(h, c), l) for h, l in zip(hi, lo)]

    def fold_g2(self, hi, lo, c):
        return [E.g2_add(E.g2_mul(h, c), l) for h, l in zip(hi, lo)]


BACKEND = PyBackend()


# ----------------------------------------------------------------------------- algebraic types
def _vec(be, name, x):
    """Backend-native container for a vector (packed limb arrays on the compiled backend)."""
    f = getattr(be, name, None)
    return f(x) if f else list(x)


class G1T:
    ser = staticmethod(ser_g1)
    add = staticmethod(E.g1_add)
    mul = staticmethod(E.g1_mul)

    @staticmethod
    def vec(x, be):
        return _vec(be, "vec_g1", x)

    @staticmethod
    def fold(hi, lo, c, be):
        return be.fold_g1(hi, lo, c)

    @staticmethod
    def msm(pts, sc, be):
        return be.msm_g1(pts, sc)


class G2T:
    ser = staticmethod(ser_g2)
    add = staticmethod(E.g2_add)
    mul = staticmethod(E.g2_mul)

    @staticmethod
    def vec(x, be):
        return _vec(be, "vec_g2", x)

    @staticmethod
    def fold(hi, lo, c, be):
        return be.fold_g2(hi, lo, c)

    @staticmethod
    def msm(pts, sc, be):
        return be.msm_g2(pts, sc)


class GTT:
    """PairingOutput: additive notation over Fq12 multiplication (App. A-3)."""

    ser = staticmethod(ser_gt)
    add = staticmethod(E.gt_mul)
    mul = staticmethod(E.gt_pow)

    @staticmethod
    def fold(hi, lo, c, be):
        return [E.gt_mul(E.gt_pow(h, c), l) for h, l in zip(hi, lo)]

    @staticmethod
    def vec(x, be):
        return list(x)


class FrT:
    ser = staticmethod(ser_fr)

    @staticmethod
    def add(a, b):
        return (a + b) % R

    @staticmethod
    def mul(a, s):
        return a * s % R

    @staticmethod
    def fold(hi, lo, c, be):
        if hasattr(be, "fold_fr"):
            return be.fold_fr(hi, lo, c)
        return [(h * c + l) % R for h, l in zip(hi, lo)]

    @staticmethod
    def vec(x, be):
        return _vec(be, "vec_fr", x)


class PlaceholderT:
    """dh_commitments/src/identity/mod.rs:18-30: unit struct, no-op Add / MulAssign, 0 bytes."""

    @staticmethod
    def ser(_):
        return b""

    @staticmethod
    def add(a, b):
        return None

    @staticmethod
    def mul(a, s):
        return None

    @staticmethod
    def fold(hi, lo, c, be):
        return [None] * len(hi)

    @staticmethod
    def vec(x, be):
        return list(x)


def IdentityOutT(T):
    """dh_commitments/src/identity/mod.rs:33-62 IdentityOutput<T>(Vec<T>)."""

    class _Out:
        @staticmethod
        def ser(v):
            return ser_vec(v, T.ser)

        @staticmethod
        def add(a, b):
            return [T.add(x, y) for x, y in zip(a, b)]

        @staticmethod
        def mul(a, s):
            return [T.mul(x, s) for x in a]

    return _Out


# ----------------------------------------------------------------------------- inner products (inner_products/src/lib.rs)
def _check_len(l, r):
    if len(l) != len(r):
        raise InnerProductError("left length, right length: %d, %d" % (len(l), len(r)))


class PairingInnerProduct:
    """inner_products/src/lib.rs:56-74."""

    Left, Right, Out = G1T, G2T, GTT

    @staticmethod
    def inner_product(l, r, be=None):
        _check_len(l, r)
        return (be or BACKEND).pairing_product(l, r)


def MultiexponentiationInnerProduct(G):
    """inner_products/src/lib.rs:123-142."""

    class _IP:
        Left, Right, Out = G, FrT, G

        @staticmethod
        def inner_product(l, r, be=None):
            _check_len(l, r)
            return G.msm(l, r, be or BACKEND)

    return _IP


class ScalarInnerProduct:
    """inner_products/src/lib.rs:149-166."""

    Left, Right, Out = FrT, FrT, FrT

    @staticmethod
    def inner_product(l, r, be=None):
        _check_len(l, r)
        return sum(a * b for a, b in zip(l, r)) % R


# ----------------------------------------------------------------------------- commitments (dh_commitments/src)
class _Commitment:
    @classmethod
    def verify(cls, k, m, com, be=None):
        """dh_commitments/src/lib.rs:52-54."""
        return cls.commit(k, m, be) == com


class AFGHOCommitmentG1(_Commitment):
    """afgho16/mod.rs:20-33: message G1, key G2, IP(m, k)."""

    Message, Key, Output = G1T, G2T, GTT

    @staticmethod
    def commit(k, m, be=None):
        return PairingInnerProduct.inner_product(m, k, be)


class AFGHOCommitmentG2(_Commitment):
    """afgho16/mod.rs:35-48: message G2, key G1, IP(k, m)."""

    Message, Key, Output = G2T, G1T, GTT

    @staticmethod
    def commit(k, m, be=None):
        return PairingInnerProduct.inner_product(k, m, be)


def PedersenCommitment(G):
    """pedersen/mod.rs:14-27: MSM(keys, msgs)."""

    class _Ped(_Commitment):
        Message, Key, Output = FrT, G, G

        @staticmethod
        def commit(k, m, be=None):
            return MultiexponentiationInnerProduct(G).inner_product(k, m, be)

    return _Ped


def IdentityCommitment(T):
    """identity/mod.rs:64-89: the commitment is the message vector itself."""

    class _Id(_Commitment):
        Message, Key, Output = T, PlaceholderT, IdentityOutT(T)

        @staticmethod
        def commit(k, m, be=None):
            return list(m)

    return _Id


class SSMPlaceholderCommitment(_Commitment):
    """tipa/structured_scalar_message.rs:29-47: commit == Fr::zero()."""

    Message, Key, Output = FrT, PlaceholderT, FrT

    @staticmethod
    def commit(k, m, be=None):
        return 0


# ----------------------------------------------------------------------------- GIPA (ip_proofs/src/gipa.rs)
class GIPAProof:
    def __init__(self, steps, base):
        self.r_commitment_steps = steps  # reversed order, list of (com_1, com_2) triples
        self.r_base = base


class GIPA:
    def __init__(self, IP, LMC, RMC, IPC, digest=blake2b, be=None):
        self.IP, self.LMC, self.RMC, self.IPC = IP, LMC, RMC, IPC
        self.digest = digest
        self.be = be or BACKEND

    # gipa.rs:235-258 / :330-353
    def _challenge(self, prev, com_1, com_2):
        nonce = 0
        L, Rm, I = self.LMC.Output, self.RMC.Output, self.IPC.Output
        while True:
            h = struct.pack(">Q", nonce) + ser_fr(prev)
            for com in (com_1, com_2):
                h += L.ser(com[0]) + Rm.ser(com[1]) + I.ser(com[2])
            c = challenge_u128(self.digest(h))
            if c != 0:
                return E.fr_inv(c), c  # (c, c_inv) swapped: gipa.rs:253-255
            nonce += 1

    # gipa.rs:108-133
    def prove(self, values, ck, com):
        m_a, m_b, t = values
        if self.IP.inner_product(m_a, m_b, self.be) != t:
            raise InnerProductError("inner product not sound")
        if bin(len(m_a)).count("1") != 1:
            raise InnerProductError("left length, right length: %d, %d" % (len(m_a), len(m_b)))
        if not (
            self.LMC.verify(ck[0], m_a, com[0], self.be)
            and self.RMC.verify(ck[1], m_b, com[1], self.be)
            and self.IPC.verify([ck[2]], [t], com[2], self.be)
        ):
            raise InnerProductError("inner product not sound")
        proof, _ = self.prove_with_aux((m_a, m_b), (ck[0], ck[1], [ck[2]]))
        return proof

    # gipa.rs:162-312
    def prove_with_aux(self, values, ck):
        IP, LMC, RMC, IPC, be = self.IP, self.LMC, self.RMC, self.IPC, self.be
        m_a, m_b = LMC.Message.vec(values[0], be), RMC.Message.vec(values[1], be)
        ck_a, ck_b, ck_t = LMC.Key.vec(ck[0], be), RMC.Key.vec(ck[1], be), list(ck[2])
        steps, transcript = [], []
        assert len(m_a) & (len(m_a) - 1) == 0 and len(m_a) > 0
        while len(m_a) > 1:
            split = len(m_a) // 2
            m_a_1, m_a_2 = m_a[split:], m_a[:split]
            ck_a_1, ck_a_2 = ck_a[:split], ck_a[split:]
            m_b_1, m_b_2 = m_b[:split], m_b[split:]
            ck_b_1, ck_b_2 = ck_b[split:], ck_b[:split]
            com_1 = (
                LMC.commit(ck_a_1, m_a_1, be),
                RMC.commit(ck_b_1, m_b_1, be),
                IPC.commit(ck_t, [IP.inner_product(m_a_1, m_b_1, be)], be),
            )
            com_2 = (
                LMC.commit(ck_a_2, m_a_2, be),
                RMC.commit(ck_b_2, m_b_2, be),
                IPC.commit(ck_t, [IP.inner_product(m_a_2, m_b_2, be)], be),
            )
            c, c_inv = self._challenge(transcript[-1] if transcript else 0, com_1, com_2)
            m_a = LMC.Message.fold(m_a_1, m_a_2, c, be)
            m_b = RMC.Message.fold(m_b_2, m_b_1, c_inv, be)
            ck_a = LMC.Key.fold(ck_a_2, ck_a_1, c_inv, be)
            ck_b = RMC.Key.fold(ck_b_1, ck_b_2, c, be)
            steps.append((com_1, com_2))
            transcript.append(c)
        transcript.reverse()
        steps.reverse()
        return GIPAProof(steps, (m_a[0], m_b[0])), {"r_transcript": transcript, "ck_base": (ck_a[0], ck_b[0])}

    # gipa.rs:322-363
    def compute_recursive_challenges(self, com, proof):
        com_a, com_b, com_t = com
        L, Rm, I = self.LMC.Output, self.RMC.Output, self.IPC.Output
        transcript = []
        for com_1, com_2 in reversed(proof.r_commitment_steps):
            c, c_inv = self._challenge(transcript[-1] if transcript else 0, com_1, com_2)
            com_a = L.add(L.add(L.mul(com_1[0], c), com_a), L.mul(com_2[0], c_inv))
            com_b = Rm.add(Rm.add(Rm.mul(com_1[1], c), com_b), Rm.mul(com_2[1], c_inv))
            com_t = I.add(I.add(I.mul(com_1[2], c), com_t), I.mul(com_2[2], c_inv))
            transcript.append(c)
        transcript.reverse()
        return (com_a, com_b, com_t), transcript

    # gipa.rs:365-399
    def compute_final_commitment_keys(self, ck, transcript):
        ck_a, ck_b = ck[0], ck[1]
        ea, eb = [1], [1]
        for i, c in enumerate(transcript):
            c_inv = E.fr_inv(c)
            for j in range(2**i):
                ea.append(ea[j] * c_inv % R)
                eb.append(eb[j] * c % R)
        assert len(ea) == len(ck_a)
        KA, KB = self.LMC.Key, self.RMC.Key
        a = KA.mul(ck_a[0], ea[0])
        for g, x in zip(ck_a[1:], ea[1:]):
            a = KA.add(a, KA.mul(g, x))
        b = KB.mul(ck_b[0], eb[0])
        for g, x in zip(ck_b[1:], eb[1:]):
            b = KB.add(b, KB.mul(g, x))
        return a, b

    # gipa.rs:401-415
    def verify_base_commitment(self, base_ck, base_com, proof):
        com_a, com_b, com_t = base_com
        ck_a_base, ck_b_base, ck_t = base_ck
        a_base, b_base = [proof.r_base[0]], [proof.r_base[1]]
        t_base = [self.IP.inner_product(a_base, b_base, self.be)]
        return (
            self.LMC.verify([ck_a_base], a_base, com_a, self.be)
            and self.RMC.verify([ck_b_base], b_base, com_b, self.be)
            and self.IPC.verify(ck_t, t_base, com_t, self.be)
        )

    # gipa.rs:135-160
    def verify(self, ck, com, proof):
        if bin(len(ck[0])).count("1") != 1 or len(ck[0]) != len(ck[1]):
            raise InnerProductError("left length, right length: %d, %d" % (len(ck[0]), len(ck[1])))
        base_com, transcript = self.compute_recursive_challenges(com, proof)
        ck_a_base, ck_b_base = self.compute_final_commitment_keys(ck, transcript)
        return self.verify_base_commitment((ck_a_base, ck_b_base, [ck[2]]), base_com, proof)

    def ser_proof(self, proof):
        """derive(CanonicalSerialize) on GIPAProof (gipa.rs:24-51): Vec of steps, then r_base."""
        L, Rm, I = self.LMC.Output, self.RMC.Output, self.IPC.Output
        out = struct.pack("<Q", len(proof.r_commitment_steps))
        for com_1, com_2 in proof.r_commitment_steps:
            for com in (com_1, com_2):
                out += L.ser(com[0]) + Rm.ser(com[1]) + I.ser(com[2])
        out += self.LMC.Message.ser(proof.r_base[0]) + self.RMC.Message.ser(proof.r_base[1])
        return out


# ----------------------------------------------------------------------------- TIPA (ip_proofs/src/tipa/mod.rs)
def structured_scalar_power(num, s):
    """structured_scalar_message.rs:334-340."""
    out = [1]
    for _ in range(1, num):
        out.append(out[-1] * s % R)
    return out


class SRS:
    """tipa/mod.rs:96-128."""

    def __init__(self, g_alpha_powers, h_beta_powers, g_beta, h_alpha):
        self.g_alpha_powers, self.h_beta_powers = g_alpha_powers, h_beta_powers
        self.g_beta, self.h_alpha = g_beta, h_alpha

    def get_commitment_keys(self):
        return self.h_beta_powers[::2], self.g_alpha_powers[::2]

    def get_verifier_key(self):
        return {"g": self.g_alpha_powers[0], "h": self.h_beta_powers[0], "g_beta": self.g_beta, "h_alpha": self.h_alpha}


def tipa_setup(size, alpha, beta, be=None):
    """tipa/mod.rs:150-164 with alpha, beta supplied instead of drawn from an rng."""
    be = be or BACKEND
    n = 2 * size - 1
    ga = be.mul_vec_g1(G1T.vec([E.G1_GEN] * n, be), structured_scalar_power(n, alpha))
    hb = be.mul_vec_g2(G2T.vec([E.G2_GEN] * n, be), structured_scalar_power(n, beta))
    return SRS(ga, hb, E.g1_mul(E.G1_GEN, beta), E.g2_mul(E.G2_GEN, alpha))


def polynomial_evaluation_product_form_from_transcript(transcript, z, r_shift):
    """tipa/mod.rs:393-405."""
    power_2_zr = z * z % R * r_shift % R
    prod = 1
    for x in transcript:
        prod = prod * (1 + x * power_2_zr) % R
        power_2_zr = power_2_zr * power_2_zr % R
    return prod


def polynomial_coefficients_from_transcript(transcript, r_shift):
    """tipa/mod.rs:407-422."""
    coeffs = [1]
    power_2_r = r_shift % R
    for i, x in enumerate(transcript):
        for j in range(2**i):
            coeffs.append(coeffs[j] * (x * power_2_r % R) % R)
        power_2_r = power_2_r * power_2_r % R
    out = []
    for i, c in enumerate(coeffs):
        out.append(c)
        if i != len(coeffs) - 1:
            out.append(0)
    return out


def kzg_quotient_coeffs(transcript, r_shift, z, n_srs):
    """tipa/mod.rs:309-330: (f - f(z)) / (X - z), zero-padded to the SRS length."""
    f = polynomial_coefficients_from_transcript(transcript, r_shift)
    assert len(f) == n_srs
    q = [0] * n_srs
    carry = 0
    for i in range(len(f) - 1, 0, -1):
        carry = (f[i] + z * carry) % R
        q[i - 1] = carry
    return q


def prove_commitment_key_kzg_opening(G, srs_powers, transcript, r_shift, z, be=None):
    """tipa/mod.rs:304-337."""
    q = kzg_quotient_coeffs(transcript, r_shift, z, len(srs_powers))
    return MultiexponentiationInnerProduct(G).inner_product(srs_powers, FrT.vec(q, be or BACKEND), be)


def verify_commitment_key_g2_kzg_opening(v_srs, ck_final, ck_opening, transcript, r_shift, z):
    """tipa/mod.rs:340-354."""
    ev = polynomial_evaluation_product_form_from_transcript(transcript, z, r_shift)
    lhs = E.pairing(v_srs["g"], E.g2_add(ck_final, E.g2_neg(E.g2_mul(v_srs["h"], ev))))
    rhs = E.pairing(E.g1_add(v_srs["g_beta"], E.g1_neg(E.g1_mul(v_srs["g"], z))), ck_opening)
    return lhs == rhs


def verify_commitment_key_g1_kzg_opening(v_srs, ck_final, ck_opening, transcript, r_shift, z):
    """tipa/mod.rs:356-370."""
    ev = polynomial_evaluation_product_form_from_transcript(transcript, z, r_shift)
    lhs = E.pairing(E.g1_add(ck_final, E.g1_neg(E.g1_mul(v_srs["g"], ev))), v_srs["h"])
    rhs = E.pairing(ck_opening, E.g2_add(v_srs["h_alpha"], E.g2_neg(E.g2_mul(v_srs["h"], z))))
    return lhs == rhs


def _kzg_challenge(digest, parts):
    """tipa/mod.rs:195-209: from_random_bytes with nonce retry."""
    nonce = 0
    while True:
        c = fr_from_random_bytes(digest(struct.pack(">Q", nonce) + parts))
        if c is not None:
            return c
        nonce += 1


class TIPAProof:
    def __init__(self, gipa_proof, final_ck, final_ck_proof):
        self.gipa_proof, self.final_ck, self.final_ck_proof = gipa_proof, final_ck, final_ck_proof


class TIPA:
    """tipa/mod.rs:32-39,130-302; LMC.Key = G2, RMC.Key = G1."""

    def __init__(self, IP, LMC, RMC, IPC, digest=blake2b, be=None):
        self.gipa = GIPA(IP, LMC, RMC, IPC, digest, be)
        self.IP, self.LMC, self.RMC, self.IPC = IP, LMC, RMC, IPC
        self.digest, self.be = digest, be or BACKEND

    def prove(self, srs, values, ck):
        return self.prove_with_srs_shift(srs, values, ck, 1)

    # tipa/mod.rs:176-231
    def prove_with_srs_shift(self, srs, values, ck, r_shift):
        proof, aux = self.gipa.prove_with_aux(values, (ck[0], ck[1], [ck[2]]))
        ck_a_final, ck_b_final = aux["ck_base"]
        transcript = aux["r_transcript"]
        transcript_inverse = [E.fr_inv(x) for x in transcript]
        r_inverse = E.fr_inv(r_shift)
        c = _kzg_challenge(self.digest, ser_fr(transcript[0]) + ser_g2(ck_a_final) + ser_g1(ck_b_final))
        ck_a_open = prove_commitment_key_kzg_opening(G2T, srs.h_beta_powers, transcript_inverse, r_inverse, c, self.be)
        ck_b_open = prove_commitment_key_kzg_opening(G1T, srs.g_alpha_powers, transcript, 1, c, self.be)
        return TIPAProof(proof, (ck_a_final, ck_b_final), (ck_a_open, ck_b_open))

    def verify(self, v_srs, ck_t, com, proof):
        return self.verify_with_srs_shift(v_srs, ck_t, com, proof, 1)

    # tipa/mod.rs:242-301
    def verify_with_srs_shift(self, v_srs, ck_t, com, proof, r_shift):
        base_com, transcript = self.gipa.compute_recursive_challenges(com, proof.gipa_proof)
        transcript_inverse = [E.fr_inv(x) for x in transcript]
        ck_a_final, ck_b_final = proof.final_ck
        ck_a_proof, ck_b_proof = proof.final_ck_proof
        c = _kzg_challenge(self.digest, ser_fr(transcript[0]) + ser_g2(ck_a_final) + ser_g1(ck_b_final))
        ck_a_valid = verify_commitment_key_g2_kzg_opening(
            v_srs, ck_a_final, ck_a_proof, transcript_inverse, E.fr_inv(r_shift), c
        )
        ck_b_valid = verify_commitment_key_g1_kzg_opening(v_srs, ck_b_final, ck_b_proof, transcript, 1, c)
        com_a, com_b, com_t = base_com
        a_base, b_base = [proof.gipa_proof.r_base[0]], [proof.gipa_proof.r_base[1]]
        t_base = [self.IP.inner_product(a_base, b_base, self.be)]
        base_valid = (
            self.LMC.verify([ck_a_final], a_base, com_a, self.be)
            and self.RMC.verify([ck_b_final], b_base, com_b, self.be)
            and self.IPC.verify([ck_t], t_base, com_t, self.be)
        )
        return ck_a_valid and ck_b_valid and base_valid

    def ser_proof(self, proof):
        """derive(CanonicalSerialize) on TIPAProof (tipa/mod.rs:41-65)."""
        return (
            self.gipa.ser_proof(proof.gipa_proof)
            + ser_g2(proof.final_ck[0])
            + self.RMC.Key.ser(proof.final_ck[1])
            + ser_g2(proof.final_ck_proof[0])
            + ser_g1(proof.final_ck_proof[1])
        )


# ----------------------------------------------------------------------------- structured scalar message (tipa/structured_scalar_message.rs)
class GIPAWithSSM:
    """structured_scalar_message.rs:49-128."""

    def __init__(self, IP, LMC, IPC, digest=blake2b, be=None):
        self.gipa = GIPA(IP, LMC, SSMPlaceholderCommitment, IPC, digest, be)
        self.IP, self.LMC, self.IPC, self.be = IP, LMC, IPC, be or BACKEND

    def prove_with_structured_scalar_message(self, values, ck):
        proof, _ = self.gipa.prove_with_aux(values, (ck[0], [None] * len(values[1]), [ck[1]]))
        return proof

    def verify_with_structured_scalar_message(self, ck, com, scalar_b, proof):
        base_com, transcript = self.gipa.compute_recursive_challenges((com[0], 0, com[1]), proof)
        ck_a_base, ck_b_base = self.gipa.compute_final_commitment_keys(
            (ck[0], [None] * len(ck[0]), ck[1]), transcript
        )
        gipa_valid = self.gipa.verify_base_commitment((ck_a_base, ck_b_base, [ck[1]]), base_com, proof)
        b_base = _ssm_final_scalar(transcript, scalar_b)
        com_a, _, com_t = base_com
        a_base = [proof.r_base[0]]
        t_base = [self.IP.inner_product(a_base, [b_base], self.be)]
        base_valid = self.LMC.verify([ck_a_base], a_base, com_a, self.be) and self.IPC.verify(
            [ck[1]], t_base, com_t, self.be
        )
        return gipa_valid and base_valid


def _ssm_final_scalar(transcript, scalar_b):
    """structured_scalar_message.rs:112-118 / :315-321."""
    power_2_b = scalar_b % R
    prod = 1
    for x in transcript:
        prod = prod * (1 + E.fr_inv(x) * power_2_b) % R
        power_2_b = power_2_b * power_2_b % R
    return prod


class TIPAWithSSMProof:
    def __init__(self, gipa_proof, final_ck, final_ck_proof):
        self.gipa_proof, self.final_ck, self.final_ck_proof = gipa_proof, final_ck, final_ck_proof


class TIPAWithSSM:
    """structured_scalar_message.rs:130-332; LMC.Key = G2."""

    def __init__(self, IP, LMC, IPC, digest=blake2b, be=None):
        self.gipa = GIPA(IP, LMC, SSMPlaceholderCommitment, IPC, digest, be)
        self.IP, self.LMC, self.IPC = IP, LMC, IPC
        self.digest, self.be = digest, be or BACKEND

    # structured_scalar_message.rs:211-268
    def prove_with_structured_scalar_message(self, srs, values, ck):
        proof, aux = self.gipa.prove_with_aux(values, (ck[0], [None] * len(values[1]), [ck[1]]))
        ck_a_final, _ = aux["ck_base"]
        transcript = aux["r_transcript"]
        transcript_inverse = [E.fr_inv(x) for x in transcript]
        c = _kzg_challenge(self.digest, ser_fr(transcript[0]) + ser_g2(ck_a_final))
        opening = prove_commitment_key_kzg_opening(G2T, srs.h_beta_powers, transcript_inverse, 1, c, self.be)
        return TIPAWithSSMProof(proof, ck_a_final, opening)

    # structured_scalar_message.rs:270-331
    def verify_with_structured_scalar_message(self, v_srs, ck_t, com, scalar_b, proof):
        base_com, transcript = self.gipa.compute_recursive_challenges((com[0], scalar_b, com[1]), proof.gipa_proof)
        transcript_inverse = [E.fr_inv(x) for x in transcript]
        ck_a_final, ck_a_proof = proof.final_ck, proof.final_ck_proof
        c = _kzg_challenge(self.digest, ser_fr(transcript[0]) + ser_g2(ck_a_final))
        ck_a_valid = verify_commitment_key_g2_kzg_opening(v_srs, ck_a_final, ck_a_proof, transcript_inverse, 1, c)
        b_base = _ssm_final_scalar(transcript, scalar_b)
        com_a, _, com_t = base_com
        a_base = [proof.gipa_proof.r_base[0]]
        t_base = [self.IP.inner_product(a_base, [b_base], self.be)]
        base_valid = self.LMC.verify([ck_a_final], a_base, com_a, self.be) and self.IPC.verify(
            [ck_t], t_base, com_t, self.be
        )
        return ck_a_valid and base_valid

    def ser_proof(self, proof):
        """derive(CanonicalSerialize) on TIPAWithSSMProof (structured_scalar_message.rs:138-156)."""
        return self.gipa.ser_proof(proof.gipa_proof) + ser_g2(proof.final_ck) + ser_g2(proof.final_ck_proof)


# ----------------------------------------------------------------------------- Groth16 aggregation (applications/groth16_aggregation.rs)
def _tipp_ab(digest, be):
    return TIPA(PairingInnerProduct, AFGHOCommitmentG1, AFGHOCommitmentG2, IdentityCommitment(GTT), digest, be)


def _tipp_c(digest, be):
    return TIPAWithSSM(MultiexponentiationInnerProduct(G1T), AFGHOCommitmentG1, IdentityCommitment(G1T), digest, be)


def _agg_challenge(digest, com_a, com_b, com_c):
    """groth16_aggregation.rs:104-116 / :172-184."""
    return _kzg_challenge(digest, ser_gt(com_a) + ser_gt(com_b) + ser_gt(com_c))


class AggregateProof:
    """groth16_aggregation.rs:58-66."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def aggregate_proofs(srs, proofs, digest=blake2b, be=None):
    """groth16_aggregation.rs:77-160.  proofs = list of (A in G1, B in G2, C in G1)."""
    be = be or BACKEND
    if hasattr(proofs, "column"):  # columnar container of packed vectors (oracle/cpu_baseline.py)
        a, b, c = proofs.column(0), proofs.column(1), proofs.column(2)
    else:
        a = G1T.vec([p[0] for p in proofs], be)
        b = G2T.vec([p[1] for p in proofs], be)
        c = G1T.vec([p[2] for p in proofs], be)
    ck_1, ck_2 = srs.get_commitment_keys()
    ck_1, ck_2 = G2T.vec(ck_1, be), G1T.vec(ck_2, be)
    IP = PairingInnerProduct
    com_a = IP.inner_product(a, ck_1, be)
    com_b = IP.inner_product(ck_2, b, be)
    com_c = IP.inner_product(c, ck_1, be)
    r = _agg_challenge(digest, com_a, com_b, com_c)
    r_vec = structured_scalar_power(len(proofs), r)
    r_inv_vec = FrT.vec([E.fr_inv(x) for x in r_vec], be)
    r_vec = FrT.vec(r_vec, be)
    a_r = be.mul_vec_g1(a, r_vec)
    ip_ab = IP.inner_product(a_r, b, be)
    agg_c = MultiexponentiationInnerProduct(G1T).inner_product(c, r_vec, be)
    ck_1_r = be.mul_vec_g2(ck_1, r_inv_vec)
    assert com_a == IP.inner_product(a_r, ck_1_r, be)
    tipa_proof_ab = _tipp_ab(digest, be).prove_with_srs_shift(srs, (a_r, b), (ck_1_r, ck_2, None), r)
    tipa_proof_c = _tipp_c(digest, be).prove_with_structured_scalar_message(srs, (c, r_vec), (ck_1, None))
    return AggregateProof(
        com_a=com_a, com_b=com_b, com_c=com_c, ip_ab=ip_ab, agg_c=agg_c,
        tipa_proof_ab=tipa_proof_ab, tipa_proof_c=tipa_proof_c,
    )


def ser_aggregate_proof(proof, digest=blake2b):
    """Field-order uncompressed encoding of AggregateProof (the struct has no serialize derive in
    the reference, groth16_aggregation.rs:58-66; this is what the derive would emit)."""
    return (
        ser_gt(proof.com_a) + ser_gt(proof.com_b) + ser_gt(proof.com_c) + ser_gt(proof.ip_ab) + ser_g1(proof.agg_c)
        + _tipp_ab(digest, None).ser_proof(proof.tipa_proof_ab)
        + _tipp_c(digest, None).ser_proof(proof.tipa_proof_c)
    )


def verify_aggregate_proof(v_srs, vk, public_inputs, proof, digest=blake2b, be=None):
    """groth16_aggregation.rs:162-231.  vk = dict(alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1)."""
    be = be or BACKEND
    r = _agg_challenge(digest, proof.com_a, proof.com_b, proof.com_c)
    ab_valid = _tipp_ab(digest, be).verify_with_srs_shift(
        v_srs, None, (proof.com_a, proof.com_b, [proof.ip_ab]), proof.tipa_proof_ab, r
    )
    c_valid = _tipp_c(digest, be).verify_with_structured_scalar_message(
        v_srs, None, (proof.com_c, [proof.agg_c]), r, proof.tipa_proof_c
    )
    n = len(public_inputs)
    r_sum = (pow(r, n, R) - 1) * E.fr_inv((r - 1) % R) % R
    p1 = E.pairing(E.g1_mul(vk["alpha_g1"], r_sum), vk["beta_g2"])
    assert len(vk["gamma_abc_g1"]) == len(public_inputs[0]) + 1
    r_vec = structured_scalar_power(n, r)
    g_ic = E.g1_mul(vk["gamma_abc_g1"][0], r_sum)
    for i, bpt in enumerate(vk["gamma_abc_g1"][1:]):
        s = ScalarInnerProduct.inner_product([inp[i] for inp in public_inputs], r_vec)
        g_ic = E.g1_add(g_ic, E.g1_mul(bpt, s))
    p2 = E.pairing(g_ic, vk["gamma_g2"])
    p3 = E.pairing(proof.agg_c, vk["delta_g2"])
    ppe_valid = proof.ip_ab == E.gt_mul(E.gt_mul(p1, p2), p3)
    return ab_valid and c_valid and ppe_valid


# ----------------------------------------------------------------------------- SIPP (sipp/src/lib.rs)
def product_of_pairings_with_coeffs(a, b, r, be=None):
    """sipp/src/lib.rs:184-217."""
    be = be or BACKEND
    return be.pairing_product(be.mul_vec_g1(a, r), b)


def product_of_pairings(a, b, be=None):
    """sipp/src/lib.rs:221-224."""
    return (be or BACKEND).pairing_product(a, b)


def _sipp_rng(a, b, r, value, digest):
    """sipp/src/lib.rs:56-60: tuple (a, b, r, value) serialised uncompressed."""
    seed = ser_vec(a, ser_g1) + ser_vec(b, ser_g2) + ser_vec(r, ser_fr) + ser_gt(value)
    return FiatShamirRng(seed, digest)


def sipp_prove(a, b, r, value, digest=blake2s, be=None):
    """sipp/src/lib.rs:42-106.  Returns the list of (z_l, z_r)."""
    be = be or BACKEND
    assert len(a) == len(b) and bin(len(a)).count("1") == 1
    rng = _sipp_rng(a, b, r, value, digest)
    a = be.mul_vec_g1(a, r)
    b = list(b)
    length = len(a)
    proof = []
    while length != 1:
        length //= 2
        a_l, a_r = a[:length], a[length:]
        b_l, b_r = b[:length], b[length:]
        z_l = be.pairing_product(a_r, b_l)
        z_r = be.pairing_product(a_l, b_r)
        proof.append((z_l, z_r))
        rng.absorb(ser_gt(z_l) + ser_gt(z_r))
        x = rng.next_u128()
        a = be.fold_g1(a_r, a_l, x)
        b = be.fold_g2(b_r, b_l, E.fr_inv(x))
    return proof


def sipp_verify(a, b, r, claimed_value, proof, digest=blake2s, be=None):
    """sipp/src/lib.rs:109-180."""
    be = be or BACKEND
    length = len(a)
    assert bin(length).count("1") == 1 and length >= 2 and length == len(b)
    proof_len = len(proof)
    assert 1 << proof_len == length
    rng = _sipp_rng(a, b, r, claimed_value, digest)
    x_s = []
    for z_l, z_r in proof:
        rng.absorb(ser_gt(z_l) + ser_gt(z_r))
        x_s.append(rng.next_u128())
    x_invs = [E.fr_inv(x) for x in x_s]
    z_prime = claimed_value
    for (z_l, z_r), x, xi in zip(proof, x_s, x_invs):
        z_prime = E.gt_mul(z_prime, E.gt_mul(E.gt_pow(z_l, x), E.gt_pow(z_r, xi)))
    s = [1] * length
    s_invs = [1] * length
    for j, (x, xi) in enumerate(zip(x_s, x_invs)):
        for i in range(length):
            if i & (1 << (proof_len - j - 1)):
                s[i] = s[i] * x % R
                s_invs[i] = s_invs[i] * xi % R
    s = [x * ri % R for x, ri in zip(s, r)]
    a_prime = be.msm_g1(a, s)
    b_prime = be.msm_g2(b, s_invs)
    return E.pairing(a_prime, b_prime) == z_prime


def ser_sipp_proof(proof):
    return b"".join(ser_gt(zl) + ser_gt(zr) for zl, zr in proof)
