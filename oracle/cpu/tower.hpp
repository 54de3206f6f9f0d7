// ORACLE (not product code): Fq2 / Fq6 / Fq12 as ark-ff instantiates them for BLS12-381
// (Fp2 u^2 = -1; Fp6_3over2 v^3 = 1 + u; Fp12_2over3over2 w^2 = v).
#pragma once
#include "field.hpp"

struct Fq2 {
  Fq c0, c1;
  static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  static Fq2 one() { return {Fq::one(), Fq::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fq2& b) const { return c0 == b.c0 && c1 == b.c1; }
  Fq2 operator+(const Fq2& b) const { return {c0 + b.c0, c1 + b.c1}; }
  Fq2 operator-(const Fq2& b) const { return {c0 - b.c0, c1 - b.c1}; }
  Fq2 operator-() const { return {-c0, -c1}; }
  Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  Fq2 operator*(const Fq2& b) const {
    Fq v0 = c0 * b.c0, v1 = c1 * b.c1;
    return {v0 - v1, (c0 + c1) * (b.c0 + b.c1) - v0 - v1};
  }
  Fq2 sqr() const {
    Fq ab = c0 * c1;
    return {(c0 + c1) * (c0 - c1), ab.dbl()};
  }
  Fq2 scale(const Fq& s) const { return {c0 * s, c1 * s}; }
  Fq2 conj() const { return {c0, -c1}; }
  Fq2 mul_nr() const { return {c0 - c1, c0 + c1}; }  // * (1 + u)
  Fq2 inv() const {
    Fq d = (c0.sqr() + c1.sqr()).inv();
    return {c0 * d, -(c1 * d)};
  }
  Fq2 pow(const u64* e, int words) const {
    Fq2 r = one();
    for (int i = words * 64 - 1; i >= 0; i--) {
      r = r.sqr();
      if ((e[i / 64] >> (i % 64)) & 1) r = r * *this;
    }
    return r;
  }
};

struct Fq6 {
  Fq2 c0, c1, c2;
  static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
  static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
  bool operator==(const Fq6& b) const { return c0 == b.c0 && c1 == b.c1 && c2 == b.c2; }
  Fq6 operator+(const Fq6& b) const { return {c0 + b.c0, c1 + b.c1, c2 + b.c2}; }
  Fq6 operator-(const Fq6& b) const { return {c0 - b.c0, c1 - b.c1, c2 - b.c2}; }
  Fq6 operator-() const { return {-c0, -c1, -c2}; }
  Fq6 mul_by_v() const { return {c2.mul_nr(), c0, c1}; }
  // Devegili et al. "Multiplication and Squaring in Pairing-Friendly Fields", section 4 (Karatsuba)
  Fq6 operator*(const Fq6& o) const {
    Fq2 ad = c0 * o.c0, be = c1 * o.c1, cf = c2 * o.c2;
    Fq2 x = (c1 + c2) * (o.c1 + o.c2) - be - cf;
    Fq2 y = (c0 + c1) * (o.c0 + o.c1) - ad - be;
    Fq2 z = (c0 + c2) * (o.c0 + o.c2) - ad + be - cf;
    return {ad + x.mul_nr(), y + cf.mul_nr(), z};
  }
  Fq6 sqr() const { return *this * *this; }
  Fq6 mul_by_01(const Fq2& a0, const Fq2& a1) const {
    Fq2 aa = c0 * a0, bb = c1 * a1;
    Fq2 t1 = ((c1 + c2) * a1 - bb).mul_nr() + aa;
    Fq2 t3 = (c0 + c2) * a0 - aa + bb;
    Fq2 t2 = (a0 + a1) * (c0 + c1) - aa - bb;
    return {t1, t2, t3};
  }
  Fq6 mul_by_1(const Fq2& a1) const { return {(c2 * a1).mul_nr(), c0 * a1, c1 * a1}; }
  Fq6 inv() const {
    Fq2 t0 = c0.sqr() - (c1 * c2).mul_nr();
    Fq2 t1 = c2.sqr().mul_nr() - c0 * c1;
    Fq2 t2 = c1.sqr() - c0 * c2;
    Fq2 d = (c0 * t0 + (c2 * t1 + c1 * t2).mul_nr()).inv();
    return {t0 * d, t1 * d, t2 * d};
  }
};

struct Fq12 {
  Fq6 c0, c1;
  static Fq2 FROB1[6];  // (1+u)^(k (p-1)/6), k = 0..5
  static Fq2 FROB2[6];  // (1+u)^(k (p^2-1)/6)
  static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
  bool operator==(const Fq12& b) const { return c0 == b.c0 && c1 == b.c1; }
  Fq12 operator*(const Fq12& o) const {
    Fq6 v0 = c0 * o.c0, v1 = c1 * o.c1;
    return {v0 + v1.mul_by_v(), (c0 + c1) * (o.c0 + o.c1) - v0 - v1};
  }
  Fq12 sqr() const {
    Fq6 v0 = c0 - c1, v3 = c0 - c1.mul_by_v(), v2 = c0 * c1;
    v0 = v0 * v3 + v2;
    return {v0 + v2.mul_by_v(), v2 + v2};
  }
  Fq12 conj() const { return {c0, -c1}; }
  Fq12 inv() const {
    Fq6 d = (c0.sqr() - c1.sqr().mul_by_v()).inv();
    return {c0 * d, -(c1 * d)};
  }
  // ark-ff Fp12::mul_by_014
  Fq12 mul_by_014(const Fq2& d0, const Fq2& d1, const Fq2& d4) const {
    Fq6 aa = c0.mul_by_01(d0, d1), bb = c1.mul_by_1(d4);
    Fq6 m = (c1 + c0).mul_by_01(d0, d1 + d4) - aa - bb;
    return {bb.mul_by_v() + aa, m};
  }
  Fq2* slot(int k) {  // coefficient of w^k
    Fq6& h = (k & 1) ? c1 : c0;
    return k / 2 == 0 ? &h.c0 : (k / 2 == 1 ? &h.c1 : &h.c2);
  }
  Fq12 frobenius(int n) const {  // n in {1, 2}
    Fq12 r = *this;
    for (int k = 0; k < 6; k++) {
      Fq2* s = r.slot(k);
      *s = (n == 1) ? s->conj() * FROB1[k] : *s * FROB2[k];
    }
    return r;
  }
  // Granger-Scott, as in ark-ff Fp12 `cyclotomic_square`
  Fq12 cyclotomic_sqr() const {
    Fq2 r0 = c0.c0, r4 = c0.c1, r3 = c0.c2, r2 = c1.c0, r1 = c1.c1, r5 = c1.c2;
    Fq2 tmp = r0 * r1;
    Fq2 t0 = (r0 + r1) * (r1.mul_nr() + r0) - tmp - tmp.mul_nr(), t1 = tmp.dbl();
    tmp = r2 * r3;
    Fq2 t2 = (r2 + r3) * (r3.mul_nr() + r2) - tmp - tmp.mul_nr(), t3 = tmp.dbl();
    tmp = r4 * r5;
    Fq2 t4 = (r4 + r5) * (r5.mul_nr() + r4) - tmp - tmp.mul_nr(), t5 = tmp.dbl();
    Fq12 o;
    o.c0.c0 = (t0 - r0).dbl() + t0;
    o.c1.c1 = (t1 + r1).dbl() + t1;
    tmp = t5.mul_nr();
    o.c1.c0 = (r2 + tmp).dbl() + tmp;
    o.c0.c2 = (t4 - r3).dbl() + t4;
    o.c0.c1 = (t2 - r4).dbl() + t2;
    o.c1.c2 = (r5 + t3).dbl() + t3;
    return o;
  }
};
