// "x3": one vector of curve points handled by a TEAM OF THREE WARPS.  Lane l of each of the three warps works on
// element l; the point state (coordinates) is replicated in the three warps' registers; every field
// multiplication level is split three ways, each warp computes one product, and the results are exchanged
// through shared memory, so the three warps always hold identical values and take identical branches.
//   * G2 (Fq2 coordinates): the three Karatsuba sub-products of every Fq2 product / squaring go to the
//     three warps (Fq2x3 below has the interface of Fq2, so the Jacobian formulas of curve.cuh are reused);
//   * G1 (Fq coordinates): the independent Fq products of each level of dbl-2009-l / madd-2007-bl.
// Why warps and not lanes: a carry-chain Montgomery product is ~300 IMAD.WIDE.U32.X, each holding the heavy pipe
// of ITS sub-partition for 4 cycles per warp instruction (profiles/README.md), so a lone warp runs at the speed of
// one sub-partition's multiplier however many of its lanes are active.  The first version of this file put the
// three roles in three LANES of one warp (shuffles): same pipe, no gain -- measured.  Three warps of a CTA sit on
// three different sub-partitions: three multipliers per element, for the folds / Horner tails of the late GIPA
// rounds, which are latency chains of ~2000-4000 dependent field products on an otherwise empty GPU.
// A CTA is ONE team (blockDim.x == 96); every thread must reach every exchange, so the point formulas below run
// their cooperative levels unconditionally and patch exceptional lanes afterwards with non-cooperative code.
#pragma once
#include "endo.cuh"

namespace ripp {
namespace x3 {

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ int lane_r() { return (threadIdx.x >> 5) % 3; }
// every warp contributes `mine` (per lane); returns the three warps' values.  [role][limb][lane]: conflict-free.
__device__ __forceinline__ void gather3(const Fq& mine, Fq& t0, Fq& t1, Fq& t2) {
  __shared__ uint32_t bus[3 * 12 * 32];
  const int r = lane_r(), l = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 12; i++) bus[(r * 12 + i) * 32 + l] = mine.v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 12; i++) {
    t0.v[i] = bus[i * 32 + l];
    t1.v[i] = bus[(12 + i) * 32 + l];
    t2.v[i] = bus[(24 + i) * 32 + l];
  }
  __syncthreads();
}
#else
int lane_r();                                             // provided by the host-simulation harness
void gather3(const Fq& mine, Fq& t0, Fq& t1, Fq& t2);
#endif

// The exchanges are CTA barriers (bar.sync, warp-aligned): a warp whose lanes left a per-lane branch -- the patch of an
// exceptional lane -- must be whole again before it reaches the next one, or its parts may arrive separately and the
// barrier release early.  With a shared scalar the patch is rare (identity inputs); an experiment with per-lane digits
// (element-wise scalings on these teams) takes it in most iterations.
RIPP_HD void reconverge() {
#if defined(__CUDA_ARCH__)
  __syncwarp();
#endif
}
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
// madd-2007-bl in five levels, executed unconditionally (every thread reaches every exchange); lanes in an
// exceptional case (either operand the identity, P = +-Q) then recompute with the complete single-thread formulas
static RIPP_FN Jac<Fq> g1_madd(Jac<Fq> p, Aff<Fq> q) {
  Fq Z1Z1, d0, d1, U2, ZZZ, S2, HH, ZH2, RR, J, V, YJ, M;
  mul3(p.z, p.z, p.z, p.z, p.z, p.z, Z1Z1, d0, d1);
  mul3(q.x, Z1Z1, p.z, Z1Z1, p.z, Z1Z1, U2, ZZZ, d0);
  Fq H = U2 - p.x;
  mul3(q.y, ZZZ, H, H, p.z + H, p.z + H, S2, HH, ZH2);
  Fq rr = (S2 - p.y).dbl();
  Fq I = HH.dbl().dbl();
  mul3(H, I, p.x, I, rr, rr, J, V, RR);
  Jac<Fq> r;
  r.x = RR - J - V.dbl();
  r.z = ZH2 - Z1Z1 - HH;
  mul3(rr, V - r.x, p.y, J, p.y, J, M, YJ, d0);
  r.y = M - YJ.dbl();
  if (q.is_inf() || p.is_inf() || H.is_zero()) r = p.add_mixed_body(q);
  reconverge();
  return r;
}

// Fq2x3 <-> Fq2 (same layout): the non-cooperative fallbacks and the endomorphism use the plain tower
RIPP_HD Fq2 plain(const Fq2x3& a) { return {a.c0, a.c1}; }
RIPP_HD Fq2x3 coop(const Fq2& a) { return {a.c0, a.c1}; }
RIPP_HD Aff<Fq2> plain(const Aff<Fq2x3>& a) { return {plain(a.x), plain(a.y)}; }
RIPP_HD Aff<Fq2x3> coop(const Aff<Fq2>& a) { return {coop(a.x), coop(a.y)}; }
RIPP_HD Jac<Fq2> plain(const Jac<Fq2x3>& a) { return {plain(a.x), plain(a.y), plain(a.z)}; }
RIPP_HD Jac<Fq2x3> coop(const Jac<Fq2>& a) { return {coop(a.x), coop(a.y), coop(a.z)}; }

// the same for G2: madd-2007-bl over the cooperative Fq2 (11 exchanges instead of 29 Fq products in a row)
// (inlined at its few call sites: 120 words of operands do not fit the register-passing convention of an
// out-of-line call and would travel through local memory -- 20 % of the kernel's instructions when they did)
RIPP_HD Jac<Fq2x3> g2_madd(const Jac<Fq2x3>& p, const Aff<Fq2x3>& q) {
  typedef Fq2x3 F;
  F Z1Z1 = p.z.sqr();
  F U2 = q.x * Z1Z1;
  F S2 = q.y * p.z * Z1Z1;
  F H = U2 - p.x;
  F rr = (S2 - p.y).dbl();
  F HH = H.sqr();
  F I = HH.dbl().dbl();
  F J = H * I;
  F V = p.x * I;
  Jac<F> r;
  r.x = rr.sqr() - J - V.dbl();
  r.y = rr * (V - r.x) - (p.y * J).dbl();
  r.z = (p.z + H).sqr() - Z1Z1 - HH;
  if (q.is_inf() || p.is_inf() || H.is_zero()) r = coop(plain(p).add_mixed_body(plain(q)));
  reconverge();
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
  RIPP_HD static Jac<Fq2x3> madd(const Jac<Fq2x3>& a, const Aff<Fq2x3>& q) { return g2_madd(a, q); }
};

RIPP_HD Aff<Fq> endo_map_x(const Aff<Fq>& p) { return endo_map(p); }
RIPP_HD Aff<Fq2x3> endo_map_x(const Aff<Fq2x3>& p) { return coop(endo_map(plain(p))); }

// affine form without the identity's early exit (Z = 0 inverts to 0, which gives (0, 0) = the packed identity)
template <class F>
RIPP_HD Aff<F> to_affine(const Jac<F>& a) {
  return a.to_affine_with(a.z.inv());
}

// sum_i d_i E^i(p) (endo.cuh) with the team's formulas; control flow depends only on the SHARED scalar
template <class F>
RIPP_FN Jac<F> endo_mul(const Aff<F>& p, const EndoBits& c) {
  Aff<F> base[4];
  base[0] = p;
  for (int t = 1; t < c.m; t++) base[t] = endo_map_x(base[t - 1]);
  Jac<F> acc = Jac<F>::inf();
#pragma unroll 1
  for (int j = c.nbits - 1; j >= 0; j--) {
    acc = Ops<F>::dbl(acc);
#pragma unroll 1
    for (int t = 0; t < c.m; t++) {
      const bool ps = (c.pos[t][j >> 5] >> (j & 31)) & 1, ng = (c.neg[t][j >> 5] >> (j & 31)) & 1;
      if (ps || ng) {  // one call site: the addition body is inlined once
        Aff<F> q = base[t];
        if (ng) q = q.neg();
        acc = Ops<F>::madd(acc, q);
      }
    }
  }
  return acc;
}

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
