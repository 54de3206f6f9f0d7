"""ORACLE (test infrastructure, NOT product code) -- the polynomial-commitment application.

PARITY UNPINNED (see oracle/bls12_381.py).  Restates
/root/reference/ip_proofs/src/applications/poly_commit/mod.rs: KZG (:52-131), BivariatePolynomial (:133-151),
BivariatePolynomialCommitment (:153-284), UnivariatePolynomialCommitment (:286-377) on top of the oracle's
TIPAWithSSM / AFGHO / MSM restatements (oracle/protocols.py).  Polynomials are coefficient lists of ints mod r,
lowest degree first; a bivariate polynomial is the list of its Y polynomials (coefficient of X^i is y_polynomials[i]).
Setup takes alpha, beta explicitly where the reference draws them from an rng.
"""
import math

from . import bls12_381 as E
from . import protocols as O
from .encoding import ser_g1

R = E.R


def _powers_g1(n, s):
    out, cur = [], 1
    for _ in range(n):
        out.append(E.g1_mul(E.G1_GEN, cur))
        cur = cur * s % R
    return out


def _powers_g2(n, s):
    out, cur = [], 1
    for _ in range(n):
        out.append(E.g2_mul(E.G2_GEN, cur))
        cur = cur * s % R
    return out


def poly_eval(coeffs, z):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * z + c) % R
    return acc


def quotient_by_linear(coeffs, z):
    """mod.rs:98-104: polynomial / (X - z), remainder dropped (synthetic division)."""
    if len(coeffs) < 2:
        return []
    q = [0] * (len(coeffs) - 1)
    carry = 0
    for i in range(len(coeffs) - 1, 0, -1):
        carry = (coeffs[i] + z * carry) % R
        q[i - 1] = carry
    return q


class KZG:
    """mod.rs:52-131."""

    @staticmethod
    def setup(degree, alpha, beta):
        g, h = E.G1_GEN, E.G2_GEN
        v_srs = {"g": g, "h": h, "g_beta": E.g1_mul(g, beta), "h_alpha": E.g2_mul(h, alpha)}
        return _powers_g1(degree + 1, alpha), v_srs

    @staticmethod
    def commit(powers, coeffs):
        assert len(powers) >= len(coeffs)  # mod.rs:85
        c = list(coeffs) + [0] * (len(powers) - len(coeffs))
        return E.msm(powers, c, E.g1_add, E.g1_mul)

    @staticmethod
    def open(powers, coeffs, point):
        assert len(powers) >= len(coeffs)  # mod.rs:98
        q = quotient_by_linear(coeffs, point)
        q = q + [0] * (len(powers) - len(q))
        return E.msm(powers, q, E.g1_add, E.g1_mul)

    @staticmethod
    def verify(v_srs, com, point, evaluation, proof):
        """mod.rs:120-130: e(com - g eval, h) == e(proof, h_alpha - h point)."""
        lhs = E.pairing(E.g1_add(com, E.g1_neg(E.g1_mul(v_srs["g"], evaluation))), v_srs["h"])
        rhs = E.pairing(proof, E.g2_add(v_srs["h_alpha"], E.g2_neg(E.g2_mul(v_srs["h"], point))))
        return lhs == rhs


def bivariate_evaluate(y_polynomials, point):
    """mod.rs:137-150."""
    x, y = point
    acc, xp = 0, 1
    for yp in y_polynomials:
        acc = (acc + xp * poly_eval(yp, y)) % R
        xp = xp * x % R
    return acc


def _ipa():
    return O.TIPAWithSSM(O.MultiexponentiationInnerProduct(O.G1T), O.AFGHOCommitmentG1, O.IdentityCommitment(O.G1T))


class BivariatePolynomialCommitment:
    """mod.rs:153-284.  srs = (ip_srs: protocols.SRS with g_alpha_powers = [g], kzg_srs: list of G1)."""

    @staticmethod
    def setup(x_degree, y_degree, alpha, beta):
        g, h = E.G1_GEN, E.G2_GEN
        kzg_srs = _powers_g1(y_degree + 1, alpha)
        srs = O.SRS([g], _powers_g2(2 * x_degree + 1, beta), E.g1_mul(g, beta), E.g2_mul(h, alpha))
        return srs, kzg_srs

    @staticmethod
    def commit(srs, y_polynomials):
        ip_srs, kzg_srs = srs
        ck, _ = ip_srs.get_commitment_keys()
        assert len(ck) >= len(y_polynomials)  # mod.rs:186
        padded = list(y_polynomials) + [[]] * (len(ck) - len(y_polynomials))
        coms = [KZG.commit(kzg_srs, yp) for yp in padded]
        return O.AFGHOCommitmentG1.commit(ck, coms), coms

    @staticmethod
    def open(srs, y_polynomials, y_polynomial_comms, point):
        """-> dict(ip_proof, y_eval_comm, kzg_proof) (OpeningProof, mod.rs:145-149)."""
        x, y = point
        ip_srs, kzg_srs = srs
        ck_1, _ = ip_srs.get_commitment_keys()
        assert len(ck_1) >= len(y_polynomials)  # mod.rs:211
        powers_of_x = O.structured_scalar_power(len(ck_1), x)
        m = len(kzg_srs)
        y_eval_coeffs = [0] * m
        for i, yp in enumerate(y_polynomials):
            for j, c in enumerate(yp):
                y_eval_coeffs[j] = (y_eval_coeffs[j] + powers_of_x[i] * c) % R
        y_eval_comm = E.msm(kzg_srs, y_eval_coeffs, E.g1_add, E.g1_mul)
        ip_proof = _ipa().prove_with_structured_scalar_message(ip_srs, (y_polynomial_comms, powers_of_x), (ck_1, None))
        kzg_proof = KZG.open(kzg_srs, y_eval_coeffs, y)
        return {"ip_proof": ip_proof, "y_eval_comm": y_eval_comm, "kzg_proof": kzg_proof}

    @staticmethod
    def verify(v_srs, com, point, evaluation, proof):
        x, y = point
        ip_ok = _ipa().verify_with_structured_scalar_message(v_srs, None, (com, [proof["y_eval_comm"]]), x, proof["ip_proof"])
        return ip_ok and KZG.verify(v_srs, proof["y_eval_comm"], y, evaluation, proof["kzg_proof"])


def ser_opening_proof(proof):
    """Field order of OpeningProof (mod.rs:145-149); the struct itself derives no serialisation in the reference."""
    return _ipa().ser_proof(proof["ip_proof"]) + ser_g1(proof["y_eval_comm"]) + ser_g1(proof["kzg_proof"])


class UnivariatePolynomialCommitment:
    """mod.rs:286-377."""

    @staticmethod
    def bivariate_degrees(univariate_degree):
        """mod.rs:292-298 (the only floating point of the path: sqrt of the degree)."""
        sqrt = 1 << (math.ceil(math.sqrt(univariate_degree + 1)) - 1).bit_length()  # next_power_of_two
        skew = 16 if sqrt >= 32 else sqrt // 2
        return sqrt // skew - 1, sqrt * skew - 1

    @staticmethod
    def degrees_from_srs(srs):
        return (len(srs[0].h_beta_powers) - 1) // 2, len(srs[1]) - 1

    @staticmethod
    def bivariate_form(degrees, coeffs):
        xd, yd = degrees
        total = (xd + 1) * (yd + 1)
        c = (list(coeffs) + [0] * total)[:total]
        return [c[i * (yd + 1) : (i + 1) * (yd + 1)] for i in range(xd + 1)]

    @classmethod
    def setup(cls, degree, alpha, beta):
        xd, yd = cls.bivariate_degrees(degree)
        return BivariatePolynomialCommitment.setup(xd, yd, alpha, beta)

    @classmethod
    def commit(cls, srs, coeffs):
        return BivariatePolynomialCommitment.commit(srs, cls.bivariate_form(cls.degrees_from_srs(srs), coeffs))

    @classmethod
    def open(cls, srs, coeffs, y_polynomial_comms, point):
        xd, yd = cls.degrees_from_srs(srs)
        x = pow(point, yd + 1, R)
        return BivariatePolynomialCommitment.open(srs, cls.bivariate_form((xd, yd), coeffs), y_polynomial_comms, (x, point))

    @classmethod
    def verify(cls, v_srs, max_degree, com, point, evaluation, proof):
        _, yd = cls.bivariate_degrees(max_degree)
        return BivariatePolynomialCommitment.verify(v_srs, com, (pow(point, yd + 1, R), point), evaluation, proof)


# ------------------------------------------------------------------------------------------------
# Transparent variant (/root/reference/ip_proofs/src/applications/poly_commit/transparent.rs): no trusted setup;
# first tier = Pedersen<G1> commitments to the Y polynomials, opened with a scalar GIPA with structured message
# (:44-56), second tier = AFGHO over those commitments, opened with the multiexponentiation GIPA with structured
# message (:28-42).  The verifier recomputes the final commitment keys itself (O(n), gipa.rs:365-399).
# ------------------------------------------------------------------------------------------------
def _second_tier():
    return O.GIPAWithSSM(O.MultiexponentiationInnerProduct(O.G1T), O.AFGHOCommitmentG1, O.IdentityCommitment(O.G1T))


def _first_tier():
    return O.GIPAWithSSM(O.ScalarInnerProduct, O.PedersenCommitment(O.G1T), O.IdentityCommitment(O.FrT))


class TransparentBivariatePolynomialCommitment:
    """transparent.rs:85-213.  ck = (first_tier_ck: y_degree + 1 G1 points, second_tier_ck: x_degree + 1 G2 points)."""

    @staticmethod
    def commit(ck, y_polynomials):
        first, second = ck
        assert len(second) >= len(y_polynomials)  # :106
        padded = list(y_polynomials) + [[]] * (len(second) - len(y_polynomials))
        coms = []
        for yp in padded:
            assert len(first) >= len(yp)  # :116
            coms.append(E.msm(first, list(yp) + [0] * (len(first) - len(yp)), E.g1_add, E.g1_mul))
        return O.AFGHOCommitmentG1.commit(second, coms), coms

    @staticmethod
    def open(ck, y_polynomials, y_polynomial_comms, point):
        x, y = point
        first, second = ck
        assert len(second) >= len(y_polynomials)  # :137
        powers_of_x = O.structured_scalar_power(len(second), x)
        y_eval_coeffs = [0] * len(first)
        for i, yp in enumerate(y_polynomials):
            for j, c in enumerate(yp):
                y_eval_coeffs[j] = (y_eval_coeffs[j] + powers_of_x[i] * c) % R
        y_eval_comm = E.msm(first, y_eval_coeffs, E.g1_add, E.g1_mul)
        second_proof = _second_tier().prove_with_structured_scalar_message((y_polynomial_comms, powers_of_x), (second, None))
        powers_of_y = O.structured_scalar_power(len(first), y)
        first_proof = _first_tier().prove_with_structured_scalar_message((y_eval_coeffs, powers_of_y), (first, None))
        return {"second_tier_ip_proof": second_proof, "y_eval_comm": y_eval_comm, "first_tier_ip_proof": first_proof}

    @staticmethod
    def verify(ck, com, point, evaluation, proof):
        first, second = ck
        x, y = point
        ok2 = _second_tier().verify_with_structured_scalar_message((second, None), (com, [proof["y_eval_comm"]]), x,
                                                                  proof["second_tier_ip_proof"])
        ok1 = _first_tier().verify_with_structured_scalar_message((first, None), (proof["y_eval_comm"], [evaluation]), y,
                                                                 proof["first_tier_ip_proof"])
        return ok2 and ok1


def ser_transparent_opening_proof(proof):
    """Field order of transparent.rs:79-83."""
    return (_second_tier().gipa.ser_proof(proof["second_tier_ip_proof"]) + ser_g1(proof["y_eval_comm"])
            + _first_tier().gipa.ser_proof(proof["first_tier_ip_proof"]))


class TransparentUnivariatePolynomialCommitment:
    """transparent.rs:215-318."""

    @staticmethod
    def bivariate_degrees(univariate_degree):
        sqrt = 1 << (math.ceil(math.sqrt(univariate_degree + 1)) - 1).bit_length()  # :222-227
        skew = 4 if sqrt >= 8 else sqrt // 2
        return sqrt // skew - 1, sqrt * skew - 1

    @staticmethod
    def degrees_from_ck(ck):
        return len(ck[1]) - 1, len(ck[0]) - 1

    @classmethod
    def commit(cls, ck, coeffs):
        form = UnivariatePolynomialCommitment.bivariate_form(cls.degrees_from_ck(ck), coeffs)
        return TransparentBivariatePolynomialCommitment.commit(ck, form)

    @classmethod
    def open(cls, ck, coeffs, y_polynomial_comms, point):
        xd, yd = cls.degrees_from_ck(ck)
        form = UnivariatePolynomialCommitment.bivariate_form((xd, yd), coeffs)
        return TransparentBivariatePolynomialCommitment.open(ck, form, y_polynomial_comms, (pow(point, yd + 1, R), point))

    @classmethod
    def verify(cls, ck, com, point, evaluation, proof):
        _, yd = cls.degrees_from_ck(ck)
        return TransparentBivariatePolynomialCommitment.verify(ck, com, (pow(point, yd + 1, R), point), evaluation, proof)
