// Optimal-ate Miller loop and final exponentiation for BLS12-381.
//
// Replaces ark-ec 0.4 `Bls12::multi_miller_loop` / `final_exponentiation` (third-party; call sites
// inner_products/src/lib.rs:83-115, sipp/src/lib.rs:196-216).  Unlike the
// CPU reference, the 68 line-coefficient triples of `G2Prepared` are never materialised: each
// line is produced on the fly from the running point T (homogeneous projective on the M-twist,
// Costello-Lange-Naehrig formulas) and consumed at once by the sparse `mul_by_014`.
// The output convention is arkworks': exponent 3 (p^12 - 1)/r (Hayashida-Hayasaka-Teruya chain).
#pragma once
#include "curve.cuh"

namespace ripp {

struct G2Proj {
  Fq2 x, y, z;
};
struct Line {
  Fq2 c0, c1, c2;  // evaluated at P as  c0 + (c1 xP) v + (c2 yP) v w
};

// T <- 2T, returns the tangent line at T.  3 M2 + 6 S2.
RIPP_FN Line dbl_step(G2Proj& r) {
  Fq2 a = (r.x * r.y).half();
  Fq2 b = r.y.sqr();
  Fq2 c = r.z.sqr();
  // e = b' * 3c with b' = 4 xi
  Fq2 e = (c.dbl() + c).mul_xi().dbl().dbl();
  Fq2 f = e.dbl() + e;
  Fq2 g = (b + f).half();
  Fq2 h = (r.y + r.z).sqr() - (b + c);
  Fq2 i = e - b;
  Fq2 j = r.x.sqr();
  Fq2 e2 = e.sqr();
  r.x = a * (b - f);
  r.y = g.sqr() - (e2.dbl() + e2);
  r.z = b * h;
  return {i, j.dbl() + j, -h};
}

// T <- T + Q, returns the chord through T and Q.  11 M2 + 2 S2.
RIPP_FN Line add_step(G2Proj& r, const G2Aff& q) {
  Fq2 theta = r.y - q.y * r.z;
  Fq2 lambda = r.x - q.x * r.z;
  Fq2 c = theta.sqr();
  Fq2 d = lambda.sqr();
  Fq2 e = lambda * d;
  Fq2 f = r.z * c;
  Fq2 g = r.x * d;
  Fq2 h = e + f - g.dbl();
  r.x = lambda * h;
  r.y = theta * (g - h) - e * r.y;
  r.z = r.z * e;
  Fq2 j = theta * q.x - lambda * q.y;
  return {j, -theta, lambda};
}

RIPP_HD void ell(Fq12& f, const Line& l, const G1Aff& p) {
  f = f.mul_by_014(l.c0, l.c1.mul_fq(p.x), l.c2.mul_fq(p.y));
}

// f_{|x|,Q}(P), conjugated because x < 0; 1 if either point is the identity.
RIPP_FN Fq12 miller_loop(const G1Aff& p, const G2Aff& q) {
  Fq12 f = Fq12::one();
  if (p.is_inf() || q.is_inf()) return f;
  G2Proj t = {q.x, q.y, Fq2::one()};
  for (int i = 62; i >= 0; i--) {
    f = f.sqr();
    Line l = dbl_step(t);
    ell(f, l, p);
    if ((k::X_ABS >> i) & 1) {
      l = add_step(t, q);
      ell(f, l, p);
    }
  }
  return f.conj();
}

// a^x for a in the cyclotomic subgroup (x = -|x|)
RIPP_FN Fq12 exp_by_x(const Fq12& a) {
  Fq12 r = a;
  for (int i = 62; i >= 0; i--) {
    r = r.cyclotomic_sqr();
    if ((k::X_ABS >> i) & 1) r = r * a;
  }
  return r.conj();
}

RIPP_FN Fq12 final_exponentiation(const Fq12& f) {
  // easy part: f^((p^6 - 1)(p^2 + 1))
  Fq12 r = f.conj() * f.inv();
  r = r.frob<2>() * r;
  // hard part: exponent (x-1)^2 (x+p)(x^2+p^2-1) + 3
  Fq12 y0 = r.cyclotomic_sqr();
  Fq12 y1 = exp_by_x(r);
  Fq12 y2 = r.conj();
  y1 = y1 * y2;
  y2 = exp_by_x(y1);
  y1 = y1.conj();
  y1 = y1 * y2;
  y2 = exp_by_x(y1);
  y1 = y1.frob<1>();
  y1 = y1 * y2;
  r = r * y0;
  y0 = exp_by_x(y1);
  y2 = exp_by_x(y0);
  y0 = y1.frob<2>();
  y1 = y1.conj();
  y1 = y1 * y2;
  y1 = y1 * y0;
  return r * y1;
}

}  // namespace ripp
