// BLS12-377: the second parameter set -- the reference's own SIPP curve (sipp/src/lib.rs:228-254
// `SIPP<Bls12_377, Blake2s>`, sipp/examples/scaling-ipp.rs:10).
//
// Replaces ark-bls12-377 0.4 (third-party): Fq (377 bits, 12 limbs), Fr (253 bits, 8 limbs),
//   Fq2 = Fq[u]/(u^2 + 5),  Fq12 = Fq2[w]/(w^6 - u) (flat basis: arkworks' c_i.c_j is the coefficient of w^(2j+i)),
//   G1: y^2 = x^3 + 1,  G2 on the D-type twist y^2 = x^3 + 1/u,  x = 0x8508c00000000001 > 0.
// The field layer (fp.cuh: CIOS Montgomery on 32-bit IMAD.WIDE chains, safegcd inversion) and the group law (curve.cuh:
// Jac<F> / Aff<F>, a = 0) are the SAME templates as BLS12-381, instantiated on this curve's parameters.  What is specific
// -- the Fq2 non-residue, xi = u, the line placement of the D-type twist (w^0, w^1, w^3: ark-ec's `mul_by_034`), no
// conjugation after the Miller loop -- lives here, one thread per pairing: this is the CORRECTNESS port of the curve
// (bit-exact against oracle/bls12_377.py); the lane-role throughput engines (l6.cuh, xt.cuh) are specialised on
// BLS12-381's tower and are not instantiated for it.
#pragma once
#include "constants377.cuh"
#include "curve.cuh"

namespace ripp {
namespace b377 {

struct FqP {
  static constexpr int N = 12;
  static constexpr uint32_t M0 = k377::FQ_M0;
  RIPP_HD static uint32_t p(int i) { return k377::FQ_P(i); }
  RIPP_HD static uint32_t one(int i) { return k377::FQ_ONE(i); }
  RIPP_HD static uint32_t r2(int i) { return k377::FQ_R2(i); }
  RIPP_HD static uint32_t pm2(int i) { return k377::FQ_PM2(i); }
  static constexpr int BITS = 377;
};
struct FrP {
  static constexpr int N = 8;
  static constexpr uint32_t M0 = k377::FR_M0;
  RIPP_HD static uint32_t p(int i) { return k377::FR_P(i); }
  RIPP_HD static uint32_t one(int i) { return k377::FR_ONE(i); }
  RIPP_HD static uint32_t r2(int i) { return k377::FR_R2(i); }
  RIPP_HD static uint32_t pm2(int i) { return k377::FR_PM2(i); }
  static constexpr int BITS = 253;
};
typedef Fp<FqP> Fq;
typedef Fp<FrP> Fr;

RIPP_HD Fq times5(const Fq& a) {
  Fq d = a.dbl();
  return d.dbl() + a;
}

// Fq[u] / (u^2 + 5)
struct Fq2 {
  Fq c0, c1;
  RIPP_HD static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  RIPP_HD static Fq2 one() { return {Fq::one(), Fq::zero()}; }
  RIPP_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  RIPP_HD bool operator==(const Fq2& b) const { return c0 == b.c0 && c1 == b.c1; }
  RIPP_HD bool operator!=(const Fq2& b) const { return !(*this == b); }
  RIPP_HD Fq2 operator+(const Fq2& b) const { return {c0 + b.c0, c1 + b.c1}; }
  RIPP_HD Fq2 operator-(const Fq2& b) const { return {c0 - b.c0, c1 - b.c1}; }
  RIPP_HD Fq2 operator-() const { return {-c0, -c1}; }
  RIPP_HD Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  RIPP_HD Fq2 conj() const { return {c0, -c1}; }
  // Karatsuba: (a0 + a1 u)(b0 + b1 u) = a0 b0 - 5 a1 b1 + ((a0 + a1)(b0 + b1) - a0 b0 - a1 b1) u
  static RIPP_FN Fq2 mul_fn(Fq2 a, Fq2 b) {
    Fq t0 = a.c0 * b.c0;
    Fq t1 = a.c1 * b.c1;
    Fq t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
    return {t0 - times5(t1), t2 - t0 - t1};
  }
  RIPP_HD Fq2 operator*(const Fq2& b) const { return mul_fn(*this, b); }
  RIPP_HD Fq2 sqr() const { return mul_fn(*this, *this); }
  RIPP_HD Fq2 mul_fq(const Fq& s) const { return {c0 * s, c1 * s}; }
  // times xi = u: (a0 + a1 u) u = -5 a1 + a0 u
  RIPP_HD Fq2 mul_xi() const { return {-times5(c1), c0}; }
  RIPP_HD Fq2 inv() const {
    Fq d = (c0 * c0 + times5(c1 * c1)).inv();
    return {c0 * d, -(c1 * d)};
  }
};

typedef Aff<Fq> G1Aff;
typedef Jac<Fq> G1Jac;
typedef Aff<Fq2> G2Aff;
typedef Jac<Fq2> G2Jac;

// Fq12 on the flat basis 1, w, .., w^5 (w^6 = u)
struct F12 {
  Fq2 c[6];
  RIPP_HD static F12 one() {
    F12 r;
    r.c[0] = Fq2::one();
    for (int i = 1; i < 6; i++) r.c[i] = Fq2::zero();
    return r;
  }
  RIPP_HD bool operator==(const F12& b) const {
    bool e = true;
    for (int i = 0; i < 6; i++) e = e && (c[i] == b.c[i]);
    return e;
  }
  RIPP_FN F12 operator*(const F12& b) const {
    F12 r;
    for (int k = 0; k < 6; k++) {
      Fq2 lo = Fq2::zero(), hi = Fq2::zero();
      for (int i = 0; i < 6; i++) {
        int j = k - i;
        if (j >= 0)
          lo = lo + c[i] * b.c[j];
        else
          hi = hi + c[i] * b.c[j + 6];
      }
      r.c[k] = lo + hi.mul_xi();
    }
    return r;
  }
  RIPP_HD F12 sqr() const { return *this * *this; }
  RIPP_HD F12 conj() const {
    F12 r;
    for (int i = 0; i < 6; i++) r.c[i] = (i & 1) ? -c[i] : c[i];
    return r;
  }
};
RIPP_HD Fq2 frob_const(int npow, int kk) {
  Fq2 r;
  for (int i = 0; i < 12; i++) {
    r.c0.v[i] = npow == 1 ? k377::FROB1(24 * kk + i) : k377::FROB2(24 * kk + i);
    r.c1.v[i] = npow == 1 ? k377::FROB1(24 * kk + 12 + i) : k377::FROB2(24 * kk + 12 + i);
  }
  return r;
}
// a^(p^n), n = 1 or 2
RIPP_FN F12 frob(const F12& a, int n) {
  F12 r;
  for (int kk = 0; kk < 6; kk++) r.c[kk] = (n == 1 ? a.c[kk].conj() : a.c[kk]) * frob_const(n, kk);
  return r;
}
// a^-1 through the norm to Fq2 (as l6.cuh / the oracle): N = a conj(a) in Fq6, d = N N^(p^2) N^(p^4) in Fq2
RIPP_FN F12 inv(const F12& a) {
  F12 ac = a.conj();
  F12 n = a * ac;
  F12 n2 = frob(n, 2);
  F12 n4 = frob(n2, 2);
  F12 t = n2 * n4;
  F12 d = n * t;
  Fq2 dinv = d.c[0].inv();
  F12 ninv;
  for (int i = 0; i < 6; i++) ninv.c[i] = t.c[i] * dinv;
  return ac * ninv;
}
// a^2 for a in the cyclotomic subgroup (Granger-Scott on the flat basis, pairs (a0, a3), (a1, a4), (a2, a5); the same
// formulas as l6.cuh's cyc_sqr with xi = u): six Fq2 products instead of thirty-six
RIPP_FN F12 cyc_sqr(const F12& a) {
  F12 r;
  Fq2 te[3], to[3];
  for (int pr = 0; pr < 3; pr++) {
    const Fq2 &ra = a.c[pr], &rb = a.c[pr + 3];
    Fq2 prod = ra * rb;
    Fq2 cross = (ra + rb) * (ra + rb.mul_xi());
    te[pr] = cross - prod - prod.mul_xi();  // ra^2 + xi rb^2
    to[pr] = prod.dbl();                    // 2 ra rb
  }
  auto three = [](const Fq2& t) { return t.dbl() + t; };
  r.c[0] = three(te[0]) - a.c[0].dbl();
  r.c[3] = three(to[0]) + a.c[3].dbl();
  r.c[1] = three(to[2].mul_xi()) + a.c[1].dbl();
  r.c[4] = three(te[2]) - a.c[4].dbl();
  r.c[2] = three(te[1]) - a.c[2].dbl();
  r.c[5] = three(to[1]) + a.c[5].dbl();
  return r;
}
// a^e for a 64-bit exponent, a in the cyclotomic subgroup (the final exponentiation's exp_by_x)
RIPP_FN F12 pow_u64(const F12& a, uint64_t e) {
  F12 r = F12::one();
  bool started = false;
  for (int i = 63; i >= 0; i--) {
    if (started) r = cyc_sqr(r);
    if ((e >> i) & 1) {
      r = started ? r * a : a;
      started = true;
    }
  }
  return r;
}
// a^e, e = 8 canonical words (GT exponentiation of the verifier)
RIPP_FN F12 pow_words(const F12& a, const uint32_t* e) {
  F12 r = F12::one();
  for (int i = 255; i >= 0; i--) {
    r = r.sqr();
    if ((e[i >> 5] >> (i & 31)) & 1) r = r * a;
  }
  return r;
}

// f * (l0 + l1 w + l3 w^3): the D-type line (ark-ec `mul_by_034`)
RIPP_FN F12 mul_line(const F12& f, const Fq2& l0, const Fq2& l1, const Fq2& l3) {
  F12 r;
  for (int k = 0; k < 6; k++) {
    Fq2 acc = f.c[k] * l0;
    int j1 = k - 1, j3 = k - 3;
    Fq2 t1 = f.c[j1 < 0 ? j1 + 6 : j1] * l1;
    Fq2 t3 = f.c[j3 < 0 ? j3 + 6 : j3] * l3;
    acc = acc + (j1 < 0 ? t1.mul_xi() : t1) + (j3 < 0 ? t3.mul_xi() : t3);
    r.c[k] = acc;
  }
  return r;
}

// f_{x,Q}(P): affine steps on the twist, line yP - lam xP w + (lam x' - y') w^3 (oracle/bls12_377.py: miller_loop)
RIPP_FN F12 miller_loop(const G1Aff& P, const G2Aff& Q) {
  if (P.is_inf() || Q.is_inf()) return F12::one();
  Fq2 tx = Q.x, ty = Q.y;
  F12 f = F12::one();
  const Fq2 yp = {P.y, Fq::zero()};
  for (int i = 62; i >= 0; i--) {
    Fq2 x2 = tx.sqr();
    Fq2 lam = (x2.dbl() + x2) * ty.dbl().inv();
    f = mul_line(f.sqr(), yp, -(lam.mul_fq(P.x)), lam * tx - ty);
    Fq2 nx = lam.sqr() - tx.dbl();
    ty = lam * (tx - nx) - ty;
    tx = nx;
    if ((k377::X >> i) & 1) {
      lam = (ty - Q.y) * (tx - Q.x).inv();
      f = mul_line(f, yp, -(lam.mul_fq(P.x)), lam * tx - ty);
      nx = lam.sqr() - tx - Q.x;
      ty = lam * (tx - nx) - ty;
      tx = nx;
    }
  }
  return f;
}

// ark-ec Bls12::final_exponentiation (same chain as BLS12-381; x > 0: exp_by_x has no conjugation)
RIPP_FN F12 final_exponentiation(const F12& f) {
  F12 r = f.conj() * inv(f);
  r = frob(r, 2) * r;
  F12 y0 = cyc_sqr(r);
  F12 y1 = pow_u64(r, k377::X);
  F12 y2 = r.conj();
  y1 = y1 * y2;
  y2 = pow_u64(y1, k377::X);
  y1 = y1.conj();
  y1 = y1 * y2;
  y2 = pow_u64(y1, k377::X);
  y1 = frob(y1, 1);
  y1 = y1 * y2;
  r = r * y0;
  y0 = pow_u64(y1, k377::X);
  y2 = pow_u64(y0, k377::X);
  y0 = frob(y1, 2);
  y1 = y1.conj();
  y1 = y1 * y2;
  y1 = y1 * y0;
  return r * y1;
}

// [s] P, s = 8 canonical words: plain double-and-add (no endomorphism on this curve's port)
template <class F>
RIPP_FN Jac<F> scalar_mul_words(const Aff<F>& p, const uint32_t* s) {
  Jac<F> acc = Jac<F>::inf();
  for (int i = 255; i >= 0; i--) {
    acc = acc.dbl();
    if ((s[i >> 5] >> (i & 31)) & 1) acc = acc.add_mixed(p);
  }
  return acc;
}

// memory order of an Fq12 value = arkworks' struct order c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2: tower slot 3 i + j holds
// the flat coefficient 2 j + i
RIPP_HD int tower_slot377(int flat) { return 3 * (flat & 1) + (flat >> 1); }

}  // namespace b377
}  // namespace ripp
