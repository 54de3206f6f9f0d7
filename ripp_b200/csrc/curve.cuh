// Short-Weierstrass (a = 0) group arithmetic for BLS12-381 G1 (over Fq) and G2 (over Fq2).
// Replaces ark-ec `short_weierstrass::{Affine, Projective}` (third-party):
// `Jac` is Jacobian (x = X/Z^2, y = Y/Z^3) exactly like arkworks' Projective; the packed affine
// form used for resident vectors encodes the identity as (0, 0), which is not on either curve.
#pragma once
#include "tower.cuh"

namespace ripp {

template <class F>
struct Aff {
  F x, y;
  RIPP_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  RIPP_HD static Aff inf() { return {F::zero(), F::zero()}; }
  RIPP_HD Aff neg() const { return {x, -y}; }
};

template <class F>
struct Jac {
  F x, y, z;
  RIPP_HD bool is_inf() const { return z.is_zero(); }
  RIPP_HD static Jac inf() { return {F::one(), F::one(), F::zero()}; }
  RIPP_HD static Jac from_affine(const Aff<F>& a) {
    if (a.is_inf()) return inf();
    return {a.x, a.y, F::one()};
  }
  RIPP_HD Jac neg() const { return {x, -y, z}; }

  // dbl-2009-l (a = 0): 2M + 5S
  RIPP_FN Jac dbl() const { return dbl_body(); }
  RIPP_HD Jac dbl_body() const {
    F A = x.sqr(), B = y.sqr(), C = B.sqr();
    F D = ((x + B).sqr() - A - C).dbl();
    F E = A.dbl() + A;
    F Fv = E.sqr();
    Jac r;
    r.z = (y * z).dbl();
    r.x = Fv - D.dbl();
    r.y = E * (D - r.x) - C.dbl().dbl().dbl();
    return r;
  }
  // madd-2007-bl: 7M + 4S
  RIPP_FN Jac add_mixed(const Aff<F>& q) const { return add_mixed_body(q); }
  RIPP_HD Jac add_mixed_body(const Aff<F>& q) const {
    if (q.is_inf()) return *this;
    if (is_inf()) return {q.x, q.y, F::one()};
    F Z1Z1 = z.sqr();
    F U2 = q.x * Z1Z1;
    F S2 = q.y * z * Z1Z1;
    F H = U2 - x;
    F rr = S2 - y;
    if (H.is_zero()) {
      if (rr.is_zero()) return dbl_body();
      return inf();
    }
    rr = rr.dbl();
    F HH = H.sqr();
    F I = HH.dbl().dbl();
    F J = H * I;
    F V = x * I;
    Jac r;
    r.x = rr.sqr() - J - V.dbl();
    r.y = rr * (V - r.x) - (y * J).dbl();
    r.z = (z + H).sqr() - Z1Z1 - HH;
    return r;
  }
  // add-2007-bl: 11M + 5S
  RIPP_FN Jac add(const Jac& q) const {
    if (q.is_inf()) return *this;
    if (is_inf()) return q;
    F Z1Z1 = z.sqr(), Z2Z2 = q.z.sqr();
    F U1 = x * Z2Z2, U2 = q.x * Z1Z1;
    F S1 = y * q.z * Z2Z2, S2 = q.y * z * Z1Z1;
    F H = U2 - U1;
    F rr = S2 - S1;
    if (H.is_zero()) {
      if (rr.is_zero()) return dbl();
      return inf();
    }
    rr = rr.dbl();
    F I = H.dbl().sqr();
    F J = H * I;
    F V = U1 * I;
    Jac r;
    r.x = rr.sqr() - J - V.dbl();
    r.y = rr * (V - r.x) - (S1 * J).dbl();
    r.z = ((z + q.z).sqr() - Z1Z1 - Z2Z2) * H;
    return r;
  }
  // register-resident variants: operands and result by value, so a running accumulator never
  // round-trips through local memory between iterations
  static RIPP_FN Jac dbl_fn(Jac a) { return a.dbl_body(); }
  static RIPP_FN Jac add_mixed_fn(Jac a, Aff<F> q) { return a.add_mixed_body(q); }
  // affine coordinates given zinv = 1/z (caller handles infinity)
  RIPP_HD Aff<F> to_affine_with(const F& zinv) const {
    F zi2 = zinv.sqr();
    return {x * zi2, y * zi2 * zinv};
  }
  RIPP_HD Aff<F> to_affine() const {
    if (is_inf()) return Aff<F>::inf();
    return to_affine_with(z.inv());
  }
};

// MSB-first double-and-add; `bits` little-endian 32-bit words of the canonical scalar, nbits significant.
// Control flow depends only on the scalar => uniform across a warp when the scalar is shared.
template <class F>
RIPP_FN Jac<F> scalar_mul(const Aff<F>& p, const uint32_t* bits, int nbits) {
  Jac<F> acc = Jac<F>::inf();
  for (int i = nbits - 1; i >= 0; i--) {
    acc = acc.dbl();
    if ((bits[i >> 5] >> (i & 31)) & 1) acc = acc.add_mixed(p);
  }
  return acc;
}

// Efficient endomorphisms (GLV / GLS).  G1: phi(x, y) = (beta x, y) = [lambda] with lambda = x^2 - 1 (128 bits).
// G2: psi(x, y) = (conj(x) cx, conj(y) cy) = [x] (x = -|x|, 64 bits), so -psi = [|x|].
RIPP_HD Aff<Fq> endo_map(const Aff<Fq>& p) {
  Fq beta;
#pragma unroll
  for (int i = 0; i < 12; i++) beta.v[i] = k::ENDO_BETA(i);
  return {p.x * beta, p.y};
}
RIPP_HD Aff<Fq2> endo_map(const Aff<Fq2>& p) {  // -psi(p) = [|x|] p
  Fq2 cx, cy;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    cx.c0.v[i] = k::PSI_CX(i);
    cx.c1.v[i] = k::PSI_CX(12 + i);
    cy.c0.v[i] = k::PSI_CY(i);
    cy.c1.v[i] = k::PSI_CY(12 + i);
  }
  if (p.is_inf()) return p;
  return {p.x.conj() * cx, -(p.y.conj() * cy)};
}

using G1Aff = Aff<Fq>;
using G1Jac = Jac<Fq>;
using G2Aff = Aff<Fq2>;
using G2Jac = Jac<Fq2>;

RIPP_HD G1Aff g1_generator() {
  G1Aff g;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    g.x.v[i] = k::G1_GEN(i);
    g.y.v[i] = k::G1_GEN(12 + i);
  }
  return g;
}
RIPP_HD G2Aff g2_generator() {
  G2Aff g;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    g.x.c0.v[i] = k::G2_GEN(i);
    g.x.c1.v[i] = k::G2_GEN(12 + i);
    g.y.c0.v[i] = k::G2_GEN(24 + i);
    g.y.c1.v[i] = k::G2_GEN(36 + i);
  }
  return g;
}

}  // namespace ripp
