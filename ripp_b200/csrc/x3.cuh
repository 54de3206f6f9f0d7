// "x3": one curve point handled by three lanes.  State (coordinates) is replicated in the three lanes'
// registers; every field multiplication level is split three ways and the results are exchanged with
// warp shuffles, so the three lanes always hold identical values and take identical branches.
//   * G2 (Fq2 coordinates): the three Karatsuba sub-products of every Fq2 product / squaring go to the
//     three lanes (Fq2x3 below has the interface of Fq2, so the Jacobian formulas of curve.cuh are
//     reused unchanged);
//   * G1 (Fq coordinates): the independent Fq products of each level of dbl-2009-l / madd-2007-bl.
// Purpose: the folds / element-wise scalings / Horner tails of a GIPA round are latency chains of ~3000
// dependent field products per element; three lanes cut that chain ~2x at the vector lengths where the
// GPU is otherwise empty.  Ten groups (30 lanes) per warp; lanes 30 and 31 idle.
#pragma once
#include "curve.cuh"

namespace ripp {
namespace x3 {

#if defined(__CUDA_ARCH__)
// Lanes 30 and 31 of a warp must have exited before any of this is called (kernels return early for them).
__device__ __forceinline__ int lane_r() { return (threadIdx.x & 31) % 3; }
__device__ __forceinline__ int lane_base() {
  int l = threadIdx.x & 31;
  return l - l % 3;
}
// every lane contributes `mine`; returns the three lanes' values.  The shuffle mask names only the group's
// own lanes, so different groups of a warp may diverge (different scalars, exceptional cases).
__device__ __forceinline__ void gather3(const Fq& mine, Fq& t0, Fq& t1, Fq& t2) {
  const int b = lane_base();
  const unsigned m = 7u << b;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    t0.v[i] = __shfl_sync(m, mine.v[i], b);
    t1.v[i] = __shfl_sync(m, mine.v[i], b + 1);
    t2.v[i] = __shfl_sync(m, mine.v[i], b + 2);
  }
}
#else
int lane_r();                                             // provided by the host-simulation harness
void gather3(const Fq& mine, Fq& t0, Fq& t1, Fq& t2);
#endif

RIPP_HD Fq fqmul(const Fq& a, const Fq& b) { return Fq::mul_fn(a, b); }
RIPP_HD Fq sel3(int r, const Fq& a, const Fq& b, const Fq& c) {
  Fq o;
#pragma unroll
  for (int i = 0; i < 12; i++) o.v[i] = r == 0 ? a.v[i] : (r == 1 ? b.v[i] : c.v[i]);
  return o;
}

// Fq2 whose products are computed cooperatively by the three lanes of a group
struct Fq2x3 {
  Fq c0, c1;
  RIPP_HD static Fq2x3 zero() { return {Fq::zero(), Fq::zero()}; }
  RIPP_HD static Fq2x3 one() { return {Fq::one(), Fq::zero()}; }
  RIPP_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  RIPP_HD bool operator==(const Fq2x3& b) const { return c0 == b.c0 && c1 == b.c1; }
  RIPP_HD Fq2x3 operator+(const Fq2x3& b) const { return {c0 + b.c0, c1 + b.c1}; }
  RIPP_HD Fq2x3 operator-(const Fq2x3& b) const { return {c0 - b.c0, c1 - b.c1}; }
  RIPP_HD Fq2x3 operator-() const { return {-c0, -c1}; }
  RIPP_HD Fq2x3 dbl() const { return {c0.dbl(), c1.dbl()}; }
  // Karatsuba: lane 0: a0 b0, lane 1: a1 b1, lane 2: (a0 + a1)(b0 + b1)
  static RIPP_FN Fq2x3 mul_fn(Fq2x3 a, Fq2x3 b) {
    const int r = lane_r();
    Fq u = sel3(r, a.c0, a.c1, a.c0 + a.c1), v = sel3(r, b.c0, b.c1, b.c0 + b.c1);
    Fq t0, t1, t2;
    gather3(fqmul(u, v), t0, t1, t2);
    return {t0 - t1, t2 - t0 - t1};
  }
  RIPP_HD Fq2x3 operator*(const Fq2x3& b) const { return mul_fn(*this, b); }
  // complex squaring: lane 0: (a0 + a1)(a0 - a1), lane 1 (and 2): a0 a1
  static RIPP_FN Fq2x3 sqr_fn(Fq2x3 a) {
    const int r = lane_r();
    Fq u = sel3(r, a.c0 + a.c1, a.c0, a.c0), v = sel3(r, a.c0 - a.c1, a.c1, a.c1);
    Fq t0, t1, t2;
    gather3(fqmul(u, v), t0, t1, t2);
    return {t0, t1.dbl()};
  }
  RIPP_HD Fq2x3 sqr() const { return sqr_fn(*this); }
  RIPP_HD Fq2x3 inv() const {
    Fq d = (c0 * c0 + c1 * c1).inv();
    return {c0 * d, -(c1 * d)};
  }
};

// ---- G1: the independent Fq products of each formula level on three lanes ------------------------
// level helper: every lane multiplies its (u, v); all lanes get the three products
RIPP_HD void mul3(const Fq& u0, const Fq& v0, const Fq& u1, const Fq& v1, const Fq& u2, const Fq& v2, Fq& p0, Fq& p1,
                  Fq& p2) {
  const int r = lane_r();
  gather3(fqmul(sel3(r, u0, u1, u2), sel3(r, v0, v1, v2)), p0, p1, p2);
}

// dbl-2009-l in three levels; the identity (Z = 0) maps to itself
static RIPP_FN Jac<Fq> g1_dbl(Jac<Fq> p) {
  Fq A, B, YZ, C, T, Fv, M, d0, d1;
  mul3(p.x, p.x, p.y, p.y, p.y, p.z, A, B, YZ);
  Fq E = A.dbl() + A, XB = p.x + B;
  mul3(B, B, XB, XB, E, E, C, T, Fv);
  Fq D = (T - A - C).dbl();
  Jac<Fq> r;
  r.x = Fv - D.dbl();
  r.z = YZ.dbl();
  mul3(E, D - r.x, E, E, E, E, M, d0, d1);
  r.y = M - C.dbl().dbl().dbl();
  return r;
}
// madd-2007-bl in five levels; exceptional cases (either operand the identity, P = +-Q) fall back to the
// complete single-lane formulas, executed redundantly (identically) by the three lanes
static RIPP_FN Jac<Fq> g1_madd(Jac<Fq> p, Aff<Fq> q) {
  if (q.is_inf() || p.is_inf()) return p.add_mixed_body(q);
  Fq Z1Z1, d0, d1, U2, ZZZ, S2, HH, ZH2, RR, J, V, YJ, M;
  mul3(p.z, p.z, p.z, p.z, p.z, p.z, Z1Z1, d0, d1);
  mul3(q.x, Z1Z1, p.z, Z1Z1, p.z, Z1Z1, U2, ZZZ, d0);
  Fq H = U2 - p.x;
  mul3(q.y, ZZZ, H, H, p.z + H, p.z + H, S2, HH, ZH2);
  Fq rr = S2 - p.y;
  if (H.is_zero()) return p.add_mixed_body(q);  // doubling or the identity
  rr = rr.dbl();
  Fq I = HH.dbl().dbl();
  mul3(H, I, p.x, I, rr, rr, J, V, RR);
  Jac<Fq> r;
  r.x = RR - J - V.dbl();
  r.z = ZH2 - Z1Z1 - HH;
  mul3(rr, V - r.x, p.y, J, p.y, J, M, YJ, d0);
  r.y = M - YJ.dbl();
  return r;
}

// acc = sum digit_i 2^i * p with digits in {-1, 0, 1} (NAF bitmaps), MSB first
template <class F>
struct Ops;
template <>
struct Ops<Fq> {
  typedef Fq Field;
  RIPP_HD static Jac<Fq> dbl(const Jac<Fq>& a) { return g1_dbl(a); }
  RIPP_HD static Jac<Fq> madd(const Jac<Fq>& a, const Aff<Fq>& q) { return g1_madd(a, q); }
};
template <>
struct Ops<Fq2x3> {
  typedef Fq2x3 Field;
  RIPP_HD static Jac<Fq2x3> dbl(const Jac<Fq2x3>& a) { return a.dbl_body(); }
  RIPP_HD static Jac<Fq2x3> madd(const Jac<Fq2x3>& a, const Aff<Fq2x3>& q) { return a.add_mixed_body(q); }
};

template <class F>
RIPP_FN Jac<F> mul_naf(const Aff<F>& p, const uint32_t* pos, const uint32_t* neg, int ndigits) {
  Aff<F> np = p.neg();
  Jac<F> acc = Jac<F>::inf();
  for (int j = ndigits - 1; j >= 0; j--) {
    acc = Ops<F>::dbl(acc);
    if ((pos[j >> 5] >> (j & 31)) & 1) acc = Ops<F>::madd(acc, p);
    if ((neg[j >> 5] >> (j & 31)) & 1) acc = Ops<F>::madd(acc, np);
  }
  return acc;
}

}  // namespace x3
}  // namespace ripp
