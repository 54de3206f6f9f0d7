// Fq2 / Fq6 / Fq12 tower for BLS12-381 (replaces ark-ff's Fp2 / Fp6_3over2 / Fp12_2over3over2 as
// instantiated by ark-bls12-381 0.4):
//   Fq2  = Fq[u]  / (u^2 + 1)
//   Fq6  = Fq2[v] / (v^3 - xi),  xi = 1 + u
//   Fq12 = Fq6[w] / (w^2 - v)
// Memory order of an Fq12 equals arkworks' struct order c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2
// (each an Fq2 = c0, c1), which is also the order `serialize_uncompressed` emits.
#pragma once
#include "fp.cuh"

namespace ripp {

struct Fq2;
RIPP_HD Fq2 fq2_mul_lazy(const Fq2& a, const Fq2& b);

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
  // Karatsuba: 3 Fq products
  // (operands and results of the out-of-line products travel in registers)
  static RIPP_FN Fq2 mul_fn(Fq2 a, Fq2 b) { return fq2_mul_lazy(a, b); }
  RIPP_HD Fq2 operator*(const Fq2& b) const { return mul_fn(*this, b); }
  // complex squaring: 2 Fq products
  static RIPP_FN Fq2 sqr_fn(Fq2 a) {
    Fq t = a.c0 * a.c1;
    return {(a.c0 + a.c1) * (a.c0 - a.c1), t.dbl()};
  }
  RIPP_HD Fq2 sqr() const { return sqr_fn(*this); }
  RIPP_HD Fq2 mul_fq(const Fq& s) const { return {c0 * s, c1 * s}; }
  RIPP_HD Fq2 half() const { return {c0.half(), c1.half()}; }
  // times xi = 1 + u
  RIPP_HD Fq2 mul_xi() const { return {c0 - c1, c0 + c1}; }
  RIPP_HD Fq2 inv() const {
    Fq d = (c0.sqr() + c1.sqr()).inv();
    return {c0 * d, -(c1 * d)};
  }
  RIPP_HD Fq2& operator+=(const Fq2& b) { return *this = *this + b; }
  RIPP_HD Fq2& operator-=(const Fq2& b) { return *this = *this - b; }
  RIPP_HD Fq2& operator*=(const Fq2& b) { return *this = *this * b; }
};

// Karatsuba in the unreduced (768-bit) domain: three integer products and TWO Montgomery reductions instead of three
// Montgomery products -- 3 x 144 + 2 x 156 = 744 MAC32 for 900, same fully reduced result.
//   c0 = REDC(a0 b0 + p R - a1 b1)      (< (p^2 + p R) / R + p < 2.2 p: two conditional subtractions)
//   c1 = REDC((a0 + a1)(b0 + b1) - a0 b0 - a1 b1) = REDC(a0 b1 + a1 b0)      (< 2 p^2 / R + p < 1.3 p: one)
RIPP_HD Fq2 fq2_mul_lazy(const Fq2& a, const Fq2& b) {
  using namespace limb;
#if defined(RIPP_HOSTSIM) && !defined(RIPP_FP_UNSATURATED)
  detail::mul_count_[1] += 3;  // the op-count model counts the reference algorithm's three Fq products
#endif
  uint32_t s0[24], s1[24], kk[24], sa[12], sb[12];
  detail::wide_mul<FqParams>(s0, a.c0.v, b.c0.v);
  detail::wide_mul<FqParams>(s1, a.c1.v, b.c1.v);
  add_cc(sa[0], a.c0.v[0], a.c1.v[0]);
#pragma unroll
  for (int i = 1; i < 11; i++) addc_cc(sa[i], a.c0.v[i], a.c1.v[i]);
  addc(sa[11], a.c0.v[11], a.c1.v[11]);
  add_cc(sb[0], b.c0.v[0], b.c1.v[0]);
#pragma unroll
  for (int i = 1; i < 11; i++) addc_cc(sb[i], b.c0.v[i], b.c1.v[i]);
  addc(sb[11], b.c0.v[11], b.c1.v[11]);
  detail::wide_mul<FqParams>(kk, sa, sb);
  detail::wide_sub<24>(kk, s0);
  detail::wide_sub<24>(kk, s1);
  add_cc(s0[12], s0[12], FqParams::p(0));
#pragma unroll
  for (int i = 1; i < 11; i++) addc_cc(s0[12 + i], s0[12 + i], FqParams::p(i));
  addc(s0[23], s0[23], FqParams::p(11));
  detail::wide_sub<24>(s0, s1);
  Fq2 r;
  detail::redc_wide<FqParams>(r.c0.v, s0, 2);
  detail::redc_wide<FqParams>(r.c1.v, kk, 1);
  return r;
}

// gamma_n[k] = xi^(k (p^n - 1)/6)
template <int NPOW>
RIPP_HD Fq2 frob_const(int kk) {
  Fq2 r;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    r.c0.v[i] = NPOW == 1 ? k::FROB1(24 * kk + i) : (NPOW == 2 ? k::FROB2(24 * kk + i) : k::FROB3(24 * kk + i));
    r.c1.v[i] = NPOW == 1 ? k::FROB1(24 * kk + 12 + i) : (NPOW == 2 ? k::FROB2(24 * kk + 12 + i) : k::FROB3(24 * kk + 12 + i));
  }
  return r;
}

struct Fq6 {
  Fq2 c0, c1, c2;
  RIPP_HD static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
  RIPP_HD static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
  RIPP_HD bool is_zero() const { return c0.is_zero() && c1.is_zero() && c2.is_zero(); }
  RIPP_HD bool operator==(const Fq6& b) const { return c0 == b.c0 && c1 == b.c1 && c2 == b.c2; }
  RIPP_HD Fq6 operator+(const Fq6& b) const { return {c0 + b.c0, c1 + b.c1, c2 + b.c2}; }
  RIPP_HD Fq6 operator-(const Fq6& b) const { return {c0 - b.c0, c1 - b.c1, c2 - b.c2}; }
  RIPP_HD Fq6 operator-() const { return {-c0, -c1, -c2}; }
  // times v
  RIPP_HD Fq6 mul_v() const { return {c2.mul_xi(), c0, c1}; }
  // Karatsuba / Toom-style: 6 Fq2 products
  RIPP_FN Fq6 operator*(const Fq6& b) const {
    Fq2 a0 = c0 * b.c0, a1 = c1 * b.c1, a2 = c2 * b.c2;
    Fq2 t0 = ((c1 + c2) * (b.c1 + b.c2) - a1 - a2).mul_xi() + a0;
    Fq2 t1 = (c0 + c1) * (b.c0 + b.c1) - a0 - a1 + a2.mul_xi();
    Fq2 t2 = (c0 + c2) * (b.c0 + b.c2) - a0 - a2 + a1;
    return {t0, t1, t2};
  }
  RIPP_FN Fq6 sqr() const {
    // CH-SQR2: 2 Fq2 products + 3 squarings
    Fq2 s0 = c0.sqr();
    Fq2 ab = c0 * c1;
    Fq2 s1 = ab.dbl();
    Fq2 s2 = (c0 - c1 + c2).sqr();
    Fq2 bc = c1 * c2;
    Fq2 s3 = bc.dbl();
    Fq2 s4 = c2.sqr();
    return {s0 + s3.mul_xi(), s1 + s4.mul_xi(), s1 + s2 + s3 - s0 - s4};
  }
  // sparse: times (b0 + b1 v): 5 Fq2 products
  RIPP_FN Fq6 mul_by_01(const Fq2& b0, const Fq2& b1) const {
    Fq2 a0 = c0 * b0, a1 = c1 * b1;
    Fq2 t0 = ((c1 + c2) * b1 - a1).mul_xi() + a0;
    Fq2 t1 = (c0 + c1) * (b0 + b1) - a0 - a1;
    Fq2 t2 = (c0 + c2) * b0 - a0 + a1;
    return {t0, t1, t2};
  }
  // sparse: times (b1 v): 3 Fq2 products
  RIPP_FN Fq6 mul_by_1(const Fq2& b1) const { return {(c2 * b1).mul_xi(), c0 * b1, c1 * b1}; }
  RIPP_HD Fq6 mul_fq2(const Fq2& s) const { return {c0 * s, c1 * s, c2 * s}; }
  RIPP_FN Fq6 inv() const {
    Fq2 t0 = c0.sqr() - (c1 * c2).mul_xi();
    Fq2 t1 = c2.sqr().mul_xi() - c0 * c1;
    Fq2 t2 = c1.sqr() - c0 * c2;
    Fq2 d = (c0 * t0 + (c2 * t1 + c1 * t2).mul_xi()).inv();
    return {t0 * d, t1 * d, t2 * d};
  }
};

struct Fq12 {
  Fq6 c0, c1;
  RIPP_HD static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
  RIPP_HD bool operator==(const Fq12& b) const { return c0 == b.c0 && c1 == b.c1; }
  RIPP_FN Fq12 operator*(const Fq12& b) const {
    Fq6 aa = c0 * b.c0, bb = c1 * b.c1;
    Fq6 m = (c0 + c1) * (b.c0 + b.c1) - aa - bb;
    return {aa + bb.mul_v(), m};
  }
  RIPP_FN Fq12 sqr() const {
    // complex squaring over Fq6: 2 Fq6 products
    Fq6 ab = c0 * c1;
    Fq6 t = (c0 + c1) * (c0 + c1.mul_v()) - ab - ab.mul_v();
    return {t, ab + ab};
  }
  // w -> -w (= p^6 Frobenius; the inverse on the cyclotomic subgroup)
  RIPP_HD Fq12 conj() const { return {c0, -c1}; }
  RIPP_FN Fq12 inv() const {
    Fq6 d = (c0.sqr() - c1.sqr().mul_v()).inv();
    return {c0 * d, -(c1 * d)};
  }
  // times the sparse element (d0 + d1 v) + (d4 v) w   [M-twist line, ark-ec `mul_by_014`]: 13 Fq2 products
  RIPP_FN Fq12 mul_by_014(const Fq2& d0, const Fq2& d1, const Fq2& d4) const {
    Fq6 aa = c0.mul_by_01(d0, d1);
    Fq6 bb = c1.mul_by_1(d4);
    Fq6 m = (c0 + c1).mul_by_01(d0, d1 + d4) - aa - bb;
    return {aa + bb.mul_v(), m};
  }
  // flat view: coefficient of w^k (k = 2j + i for slot c_i.c_j)
  RIPP_HD Fq2& w(int kk) {
    Fq6& h = (kk & 1) ? c1 : c0;
    int j = kk >> 1;
    return j == 0 ? h.c0 : (j == 1 ? h.c1 : h.c2);
  }
  // p^n Frobenius for n = 1, 2, 3
  template <int NPOW>
  RIPP_FN Fq12 frob() const {
    Fq12 r = *this;
#pragma unroll
    for (int kk = 0; kk < 6; kk++) {
      Fq2 c = r.w(kk);
      if (NPOW & 1) c = c.conj();
      if (kk > 0) {
        c = c * frob_const<NPOW>(kk);
      }
      r.w(kk) = c;
    }
    return r;
  }
  // Granger-Scott squaring, valid only in the cyclotomic subgroup: 9 Fq2 squarings-equivalents (6 products)
  RIPP_FN Fq12 cyclotomic_sqr() const {
    const Fq2 &r0 = c0.c0, &r4 = c0.c1, &r3 = c0.c2, &r2 = c1.c0, &r1 = c1.c1, &r5 = c1.c2;
    Fq2 tmp = r0 * r1;
    Fq2 t0 = (r0 + r1) * (r1.mul_xi() + r0) - tmp - tmp.mul_xi();
    Fq2 t1 = tmp.dbl();
    tmp = r2 * r3;
    Fq2 t2 = (r2 + r3) * (r3.mul_xi() + r2) - tmp - tmp.mul_xi();
    Fq2 t3 = tmp.dbl();
    tmp = r4 * r5;
    Fq2 t4 = (r4 + r5) * (r5.mul_xi() + r4) - tmp - tmp.mul_xi();
    Fq2 t5 = tmp.dbl();
    Fq12 o;
    o.c0.c0 = (t0 - r0).dbl() + t0;
    o.c1.c1 = (t1 + r1).dbl() + t1;
    tmp = t5.mul_xi();
    o.c1.c0 = (tmp + r2).dbl() + tmp;
    o.c0.c2 = (t4 - r3).dbl() + t4;
    o.c0.c1 = (t2 - r4).dbl() + t2;
    o.c1.c2 = (t3 + r5).dbl() + t3;
    return o;
  }
};

}  // namespace ripp
