// "L6": an Fq12 value spread over six lanes of a warp, one Fq2 coefficient of the flat basis
// 1, w, ..., w^5 (w^6 = xi) per lane, with the group's shared-memory scratch as the exchange medium.
//
// Why: a thread-per-pairing Miller loop needs ~250 live 32-bit registers of state and ends up
// streaming its tower through local memory; worse, at the small vector lengths of the late GIPA
// rounds (and for the final exponentiation, of which there is one per product) a single thread is a
// ~10^7-instruction serial chain.  Spreading one Fq12 over six lanes keeps every coefficient in
// registers, turns Fq12 products into six independent Fq2 accumulations (schoolbook in w, which is as
// cheap as the Karatsuba tower once the outputs are computed in parallel) and cuts the latency of a
// Miller loop / final exponentiation by ~4-5x while staying work-efficient at large n.
//
// Five groups (30 lanes) per warp; the last two lanes run the same instruction stream on a junk slot.
// All lanes execute identical code: per-lane differences are data (smem addresses, selects), never
// branches.  Host build (tests/hostsim): each lane is a std::thread, sync() is a barrier.
#pragma once
#include "pairing.cuh"

namespace ripp {
namespace l6 {

// ---- group scratch layout (32-bit words) ---------------------------------------------------------
constexpr int FQ2W = 24;
constexpr int F12W = 6 * FQ2W;
constexpr int OFF_R = 0;                  // round results R0..R5
constexpr int OFF_LINE = OFF_R + F12W;    // line coefficients d0, d1, d4
constexpr int OFF_F = OFF_LINE + 3 * FQ2W;  // Fq12 registers F0..F(nreg-1), then the per-pair blocks
// per-pair block (a group can walk several pairs that share one accumulator, ark-ec's multi_miller_loop shape)
constexpr int PB_T = 0;                   // running point T: x, y, z
constexpr int PB_P = PB_T + 3 * FQ2W;     // xP, -yP (Fq each; stage_pair_p)
constexpr int PB_QPTR = PB_P + FQ2W;      // address of Q (affine, global memory): read at the start and by the five
                                          // addition steps only -- 192 B of shared memory per pair bought a third CTA per SM
constexpr int PB_VALID = PB_QPTR + 2;     // 1 = finite pair, 0 = contributes the constant line 1
constexpr int PAIR_WORDS = PB_QPTR + 8;   // 104 words
RIPP_HD constexpr int group_words(int nreg, int npairs) { return OFF_F + nreg * F12W + npairs * PAIR_WORDS; }

// W = 1: the coefficient's lane does the whole Fq2 arithmetic (throughput shape: five groups per warp).
// W = 3: THREE lanes per coefficient (18 lanes per Fq12, one group per warp; lanes 18..31 mirror lanes 0..13 and
// recompute the same values), one Karatsuba role each -- role 0: a0 b0, role 1: a1 b1, role 2: (a0 + a1)(b0 + b1) --
// so every Fq2 product level costs ONE Fq product of latency instead of three.  A warp instruction holds the
// multiplier pipe for the same time however many lanes are active, so for the lone warps of the late GIPA rounds,
// the final exponentiation and the verifier's GT powers this is a ~2.5x shorter chain at no extra pipe time.
// The three lanes of a coefficient hold identical Fq2 state; partial products are exchanged through `bus`.
constexpr int BUS_WORDS = 6 * 3 * 12;  // one Fq per (coefficient, role); two buffers are used alternately
constexpr int BUS_ZERO = 2 * BUS_WORDS;   // twelve words that stay zero: the padding operand of the address-driven sums below
constexpr int BUS_TOTAL = BUS_ZERO + 16;  // what a W = 3 group reserves behind its registers
template <int W_>
struct CtxT {
  static constexpr int W = W_;
  int k;          // coefficient (lane within the group for W = 1), 0..5
  uint32_t* sm;   // group scratch
  void* bar;      // host build (tests/hostsim): barrier object; unused on the device
  int role;       // W = 3: Karatsuba role of this lane, 0..2
  uint32_t* bus;  // W = 3: BUS_TOTAL words of the group's scratch (zero slot cleared by whoever builds the context)
  mutable int par;  // W = 3: which bus buffer the next exchange uses
  const uint32_t* zero;  // Miller loop: twelve words that stay zero (padding operand of lz_pass); W = 3: bus + BUS_ZERO
};
using Ctx = CtxT<1>;
using Ctx3 = CtxT<3>;

#if defined(__CUDA_ARCH__)
template <class C>
__device__ __forceinline__ void sync(const C&) { __syncwarp(); }
#else
void host_barrier(void* bar);
template <class C>
inline void sync(const C& c) { host_barrier(c.bar); }
#endif

// ---- force-inlined field helpers (everything stays in registers inside an L6 kernel) -------------
RIPP_HD Fq fqmul(const Fq& a, const Fq& b) {
  Fq r;
  detail::mont_mul<FqParams>(r.v, a.v, b.v);
  return r;
}
RIPP_HD Fq2 f2add(const Fq2& a, const Fq2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
RIPP_HD Fq2 f2sub(const Fq2& a, const Fq2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
RIPP_HD Fq2 f2neg(const Fq2& a) { return {-a.c0, -a.c1}; }
RIPP_HD Fq2 f2dbl(const Fq2& a) { return {a.c0 + a.c0, a.c1 + a.c1}; }
RIPP_HD Fq2 f2half(const Fq2& a) { return {a.c0.half(), a.c1.half()}; }
RIPP_HD Fq2 f2xi(const Fq2& a) { return {a.c0 - a.c1, a.c0 + a.c1}; }
RIPP_HD Fq2 f2conj(const Fq2& a) { return {a.c0, -a.c1}; }
RIPP_HD Fq ld1(const uint32_t* p) {
  Fq r;
#if defined(__CUDA_ARCH__)
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    uint4 v = q[i];
    r.v[4 * i] = v.x;
    r.v[4 * i + 1] = v.y;
    r.v[4 * i + 2] = v.z;
    r.v[4 * i + 3] = v.w;
  }
#else
  for (int i = 0; i < 12; i++) r.v[i] = p[i];
#endif
  return r;
}
RIPP_HD void st1(uint32_t* p, const Fq& a) {
#if defined(__CUDA_ARCH__)
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 3; i++) q[i] = make_uint4(a.v[4 * i], a.v[4 * i + 1], a.v[4 * i + 2], a.v[4 * i + 3]);
#else
  for (int i = 0; i < 12; i++) p[i] = a.v[i];
#endif
}
RIPP_HD Fq sel3(int r, const Fq& a, const Fq& b, const Fq& c) {
  Fq o;
#pragma unroll
  for (int i = 0; i < 12; i++) o.v[i] = r == 0 ? a.v[i] : (r == 1 ? b.v[i] : c.v[i]);
  return o;
}
// W = 3: every lane of a coefficient's triple contributes `mine`; all three get (t0, t1, t2) = the values of roles 0, 1, 2.
// Two buffers used alternately: a lane may start writing exchange n + 2 only after the barrier of exchange n + 1,
// which every lane reaches after it has read exchange n -- so one barrier per exchange is enough.
template <class C>
RIPP_HD void gather3(const C& c, const Fq& mine, Fq& t0, Fq& t1, Fq& t2) {
  uint32_t* b = c.bus + c.par * BUS_WORDS + c.k * 36;
  c.par ^= 1;
  st1(b + c.role * 12, mine);
  sync(c);
  t0 = ld1(b);
  t1 = ld1(b + 12);
  t2 = ld1(b + 24);
}
// c0 + c1 over the integers (< 2p < 2^382): operand of the Karatsuba cross product
RIPP_HD void add_unreduced(uint32_t* s, const Fq& a, const Fq& b) {
  using namespace limb;
  add_cc(s[0], a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 11; i++) addc_cc(s[i], a.v[i], b.v[i]);
  addc(s[11], a.v[11], b.v[11]);
}
// acc (12 limbs) += x;  the caller's bound keeps the sum below 2^384
RIPP_HD void lz_add12(uint32_t* acc, const Fq& x) {
  using namespace limb;
  add_cc(acc[0], acc[0], x.v[0]);
#pragma unroll
  for (int i = 1; i < 11; i++) addc_cc(acc[i], acc[i], x.v[i]);
  addc(acc[11], acc[11], x.v[11]);
}
RIPP_HD void lz_sub12(uint32_t* acc, const Fq& x) {
  using namespace limb;
  sub_cc(acc[0], acc[0], x.v[0]);
#pragma unroll
  for (int i = 1; i < 11; i++) subc_cc(acc[i], acc[i], x.v[i]);
  subc(acc[11], acc[11], x.v[11]);
}
// One out-of-line copy of the Fq2 product on the device (operands by value, in registers) instead of an inlined
// copy per use: the Miller loop body drops from 24 k to 18 k instructions (2^16 pairs: 21.4 -> 21.1 ms, small
// batches 3.0 -> 2.7 ms: fewer instruction-cache misses for the lone warps of the late GIPA rounds).
RIPP_HD Fq2 f2mul_body(const Fq2& a, const Fq2& b) { return fq2_mul_lazy(a, b); }  // two reductions, not three (tower.cuh)
// the role's Karatsuba product of a b (W = 3); the caller exchanges the three and combines
// (the cross operands stay unreduced, < 2p each: the product is < 4 p^2 / R + p < 2p, which the final subtraction handles)
RIPP_HD Fq f2mul_part_body(const Fq2& a, const Fq2& b, int role) {
  Fq sa, sb;
  add_unreduced(sa.v, a.c0, a.c1);
  add_unreduced(sb.v, b.c0, b.c1);
  return fqmul(sel3(role, a.c0, a.c1, sa), sel3(role, b.c0, b.c1, sb));
}
#if defined(__CUDA_ARCH__) && !defined(RIPP_L6_INLINE_F2MUL)
static __device__ __noinline__ Fq2 f2mul_fn(Fq2 a, Fq2 b) { return f2mul_body(a, b); }
static __device__ __noinline__ Fq f2mul_part_fn(Fq2 a, Fq2 b, int role) { return f2mul_part_body(a, b, role); }
#else
RIPP_HD Fq2 f2mul_fn(const Fq2& a, const Fq2& b) { return f2mul_body(a, b); }
RIPP_HD Fq f2mul_part_fn(const Fq2& a, const Fq2& b, int role) { return f2mul_part_body(a, b, role); }
#endif
template <class C>
RIPP_HD Fq2 f2mul(const C& c, const Fq2& a, const Fq2& b) {
  if constexpr (C::W == 1) {
    return f2mul_fn(a, b);
  } else {
    Fq t0, t1, t2;
    gather3(c, f2mul_part_fn(a, b, c.role), t0, t1, t2);
    return {t0 - t1, t2 - t0 - t1};
  }
}
// ---- lazy Fq2 multiply-accumulate: sum_i a_i b_i with ONE Montgomery reduction per output limb vector ----
// Karatsuba in the wide (768-bit) domain: S0 += a0 b0, S1 += a1 b1, K += (a0 + a1)(b0 + b1); then
// c0 = REDC(S0 + p R - S1), c1 = REDC(K - S0 - S1).  Operand components must be < p; up to six products.
// W = 1: the lane keeps all three sums (Acc3).  W = 3: each role keeps its own (Acc1), reduces it, and the three
// reduced values are exchanged (REDC is linear modulo p, so REDC(K) - REDC(S0) - REDC(S1) is the same c1).
struct Acc3 {
  uint32_t s0[24], s1[24], k[24];
};
struct Acc1 {
  uint32_t s[24];
};
template <class C> struct AccOf { typedef Acc3 type; };
template <> struct AccOf<CtxT<3>> { typedef Acc1 type; };
RIPP_HD void acc_zero(Acc3& A) {
#pragma unroll
  for (int i = 0; i < 24; i++) A.s0[i] = A.s1[i] = A.k[i] = 0;
}
RIPP_HD void acc_zero(Acc1& A) {
#pragma unroll
  for (int i = 0; i < 24; i++) A.s[i] = 0;
}
template <class C>
RIPP_HD void f2_mac(const C&, Acc3& A, const Fq2& a, const Fq2& b) {
  uint32_t t[24];
  detail::wide_mul<FqParams>(t, a.c0.v, b.c0.v);
  detail::wide_add<24>(A.s0, t);
  detail::wide_mul<FqParams>(t, a.c1.v, b.c1.v);
  detail::wide_add<24>(A.s1, t);
  uint32_t sa[12], sb[12];
  add_unreduced(sa, a.c0, a.c1);
  add_unreduced(sb, b.c0, b.c1);
  detail::wide_mul<FqParams>(t, sa, sb);
  detail::wide_add<24>(A.k, t);
}
// W = 3, operands by ADDRESS (canonical Fq2 values a, b in shared memory): the role's operand pair is formed as an
// unreduced sum of slots instead of loading both components, reducing xi b and selecting --
//   u = a0 | a1 | a0 + a1;   v = b0 | b1 | b0 + b1;   with `wrap` (the term carries xi):  v = b0 - b1 + p | b0 + b1 | 2 b0
// (the three are the Karatsuba operands of xi b = (b0 - b1, b0 + b1) modulo p).  u, v < 2p: six products stay below
// 24 p^2, the bound f2_finish already reduces from.
template <class C>
RIPP_HD void f2_mac_addr(const C& c, Acc1& A, const uint32_t* a, const uint32_t* b, bool wrap) {
  using namespace limb;
  const int r = c.role;
  const uint32_t* Z = c.zero;
  Fq u = ld1(r == 1 ? a + 12 : a);
  lz_add12(u.v, ld1(r == 2 ? a + 12 : Z));
  Fq v = ld1((!wrap && r == 1) ? b + 12 : b);
  lz_add12(v.v, ld1(wrap ? (r == 0 ? Z : (r == 1 ? b + 12 : b)) : (r == 2 ? b + 12 : Z)));
  const bool neg = wrap && r == 0;
  const uint32_t m = neg ? 0xffffffffu : 0u;
  add_cc(v.v[0], v.v[0], FqParams::p(0) & m);
#pragma unroll
  for (int i = 1; i < 11; i++) addc_cc(v.v[i], v.v[i], FqParams::p(i) & m);
  addc(v.v[11], v.v[11], FqParams::p(11) & m);
  lz_sub12(v.v, ld1(neg ? b + 12 : Z));
  uint32_t t[24];
  detail::wide_mul<FqParams>(t, u.v, v.v);
  detail::wide_add<24>(A.s, t);
}
// The squaring's term  [2] [xi] a_i a_j  the same way: `on` = false adds nothing (coefficients with three terms), `twice`
// doubles v (cross terms).  u < 2p, v < 4p: four terms stay below 32 p^2, reduced by f2_finish(c, acc, 4).
template <class C>
RIPP_HD void f2_mac_addr_sq(const C& c, Acc1& A, const uint32_t* a, const uint32_t* b, bool wrap, bool twice, bool on) {
  using namespace limb;
  const int r = c.role;
  const uint32_t* Z = c.zero;
  Fq u = ld1(!on ? Z : (r == 1 ? a + 12 : a));
  lz_add12(u.v, ld1((on && r == 2) ? a + 12 : Z));
  Fq v = ld1((!wrap && r == 1) ? b + 12 : b);
  lz_add12(v.v, ld1(wrap ? (r == 0 ? Z : (r == 1 ? b + 12 : b)) : (r == 2 ? b + 12 : Z)));
  const bool neg = wrap && r == 0;
  const uint32_t m = neg ? 0xffffffffu : 0u;
  add_cc(v.v[0], v.v[0], FqParams::p(0) & m);
#pragma unroll
  for (int i = 1; i < 11; i++) addc_cc(v.v[i], v.v[i], FqParams::p(i) & m);
  addc(v.v[11], v.v[11], FqParams::p(11) & m);
  lz_sub12(v.v, ld1(neg ? b + 12 : Z));
  {
    const uint32_t m2 = twice ? 0xffffffffu : 0u;
    Fq w;
#pragma unroll
    for (int i = 0; i < 12; i++) w.v[i] = v.v[i] & m2;
    lz_add12(v.v, w);
  }
  uint32_t t[24];
  detail::wide_mul<FqParams>(t, u.v, v.v);
  detail::wide_add<24>(A.s, t);
}
template <class C>
RIPP_HD Fq2 f2_finish(const C&, Acc3& A) {
  using namespace limb;
  detail::wide_sub<24>(A.k, A.s0);
  detail::wide_sub<24>(A.k, A.s1);
  // S0 + p 2^384 - S1 >= 0 (S1 < 6 p^2 < p 2^384): one reduction gives c0 directly
  add_cc(A.s0[12], A.s0[12], FqParams::p(0));
#pragma unroll
  for (int i = 1; i < 11; i++) addc_cc(A.s0[12 + i], A.s0[12 + i], FqParams::p(i));
  addc(A.s0[23], A.s0[23], FqParams::p(11));
  detail::wide_sub<24>(A.s0, A.s1);
  Fq c0, c1;
  detail::redc_wide<FqParams>(c0.v, A.s0, 2);  // < (6 p^2 + p R) / R + p < 2.7 p
  detail::redc_wide<FqParams>(c1.v, A.k, 2);   // < 12 p^2 / R + p < 2.3 p
  return {c0, c1};
}
template <class C>
RIPP_HD Fq2 f2_finish(const C& c, Acc1& A, int subs = 3) {
  Fq r, t0, t1, t2;
  detail::redc_wide<FqParams>(r.v, A.s, subs);  // role 2: < 24 p^2 / R + p < 3.5 p  (squaring by address: 32 p^2, 4.3 p, subs = 4)
  gather3(c, r, t0, t1, t2);
  return {t0 - t1, t2 - t0 - t1};
}

// Q of a pair block (pointer stored as two words)
RIPP_HD const uint32_t* pair_q(const uint32_t* pb) {
  uint64_t a = (uint64_t)pb[PB_QPTR] | ((uint64_t)pb[PB_QPTR + 1] << 32);
  return reinterpret_cast<const uint32_t*>(a);
}
RIPP_HD void set_pair_q(uint32_t* pb, const void* q) {
  uint64_t a = (uint64_t)reinterpret_cast<uintptr_t>(q);
  pb[PB_QPTR] = (uint32_t)a;
  pb[PB_QPTR + 1] = (uint32_t)(a >> 32);
}
RIPP_HD Fq2 f2sel(bool c, const Fq2& a, const Fq2& b) {  // c ? a : b, branch-free
  Fq2 r;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    r.c0.v[i] = c ? a.c0.v[i] : b.c0.v[i];
    r.c1.v[i] = c ? a.c1.v[i] : b.c1.v[i];
  }
  return r;
}
RIPP_HD Fq2 ld2(const uint32_t* p) {
  Fq2 r;
#if defined(__CUDA_ARCH__)
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 6; i++) {
    uint4 v = q[i];
    uint32_t* d = (i < 3 ? r.c0.v : r.c1.v) + 4 * (i % 3);
    d[0] = v.x;
    d[1] = v.y;
    d[2] = v.z;
    d[3] = v.w;
  }
#else
  for (int i = 0; i < 12; i++) {
    r.c0.v[i] = p[i];
    r.c1.v[i] = p[12 + i];
  }
#endif
  return r;
}
RIPP_HD void st2(uint32_t* p, const Fq2& a) {
#if defined(__CUDA_ARCH__)
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const uint32_t* s = (i < 3 ? a.c0.v : a.c1.v) + 4 * (i % 3);
    q[i] = make_uint4(s[0], s[1], s[2], s[3]);
  }
#else
  for (int i = 0; i < 12; i++) {
    p[i] = a.c0.v[i];
    p[12 + i] = a.c1.v[i];
  }
#endif
}

template <class C>
RIPP_HD uint32_t* freg(const C& c, int r) { return c.sm + OFF_F + r * F12W; }

// ---- Fq12 ops on smem registers (collective: all lanes of the group call with the same arguments) ---------
// D = A * B on raw coefficient arrays (6 x Fq2, flat w-basis); D may alias A or B
template <class C>
RIPP_HD void mul_p_body(const C& c, uint32_t* D, const uint32_t* A, const uint32_t* B) {
  typename AccOf<C>::type acc;
  acc_zero(acc);
#pragma unroll 1
  for (int i = 0; i < 6; i++) {
    int j = c.k - i;
    bool wrap = j < 0;
    j += wrap ? 6 : 0;
    if constexpr (C::W == 3) {
      f2_mac_addr(c, acc, A + i * FQ2W, B + j * FQ2W, wrap);
    } else {
      Fq2 b = ld2(B + j * FQ2W);
      f2_mac(c, acc, ld2(A + i * FQ2W), f2sel(wrap, f2xi(b), b));  // xi applied to the operand: stays linear
    }
  }
  Fq2 out = f2_finish(c, acc);
  sync(c);
  st2(D + c.k * FQ2W, out);
  sync(c);
}
#if defined(__CUDA_ARCH__) && defined(RIPP_L6_CALL_W3)
static __device__ __noinline__ void mul_p_w3(const CtxT<3>& c, uint32_t* D, const uint32_t* A, const uint32_t* B) { mul_p_body(c, D, A, B); }
template <class C>
RIPP_HD void mul_p(const C& c, uint32_t* D, const uint32_t* A, const uint32_t* B) {
  if constexpr (C::W == 3) mul_p_w3(c, D, A, B); else mul_p_body(c, D, A, B);
}
#else
template <class C>
RIPP_HD void mul_p(const C& c, uint32_t* D, const uint32_t* A, const uint32_t* B) { mul_p_body(c, D, A, B); }
#endif
// dst = a * b;  dst may alias a or b
template <class C>
RIPP_HD void mul(const C& c, int dst, int a, int b) { mul_p(c, freg(c, dst), freg(c, a), freg(c, b)); }
// dst = a^2: 21 distinct products over six coefficients (cross terms doubled)
template <class C>
RIPP_HD void sqr_body(const C& c, int dst, int a) {
  const uint32_t* A = freg(c, a);
  typename AccOf<C>::type acc;
  acc_zero(acc);
  // pairs (i, j), i <= j, i + j = k or k + 6:  i runs over 0..3 slots; slots beyond the coefficient's count are masked
#pragma unroll 1
  for (int s = 0; s < 4; s++) {
    // s-th solution for this coefficient: enumerate i = 0..5 with j = (k - i) mod 6 >= i
    int cnt = -1, ii = 0, jj = 0;
    bool wrap = false, found = false;
#pragma unroll
    for (int i = 0; i < 6; i++) {
      int j = c.k - i;
      bool w = j < 0;
      j += w ? 6 : 0;
      bool ok = j >= i;
      cnt += ok ? 1 : 0;
      bool take = ok && cnt == s && !found;
      ii = take ? i : ii;
      jj = take ? j : jj;
      wrap = take ? w : wrap;
      found = found || take;
    }
    if constexpr (C::W == 3) {
      f2_mac_addr_sq(c, acc, A + ii * FQ2W, A + jj * FQ2W, wrap, ii != jj, found);
    } else {
      Fq2 x = ld2(A + ii * FQ2W), y = ld2(A + jj * FQ2W);
      y = f2sel(ii != jj, f2dbl(y), y);
      y = f2sel(wrap, f2xi(y), y);
      x = f2sel(found, x, Fq2::zero());  // coefficients with only three terms add 0 in the fourth slot
      f2_mac(c, acc, x, y);
    }
  }
  Fq2 out;
  if constexpr (C::W == 3) out = f2_finish(c, acc, 4); else out = f2_finish(c, acc);
  sync(c);
  st2(freg(c, dst) + c.k * FQ2W, out);
  sync(c);
}
// dst = a * (d0 + d1 w^2 + d4 w^3), line coefficients at OFF_LINE (the ark-ec `mul_by_014` shape)
template <class C>
RIPP_HD void mul_line_body(const C& c, int dst, int a) {
  const uint32_t* A = freg(c, a);
  const uint32_t* L = c.sm + OFF_LINE;
  int k = c.k;
  int k2 = k - 2, k3 = k - 3;
  bool w2 = k2 < 0, w3 = k3 < 0;
  k2 += w2 ? 6 : 0;
  k3 += w3 ? 6 : 0;
  typename AccOf<C>::type acc;
  acc_zero(acc);
  if constexpr (C::W == 3) {
    f2_mac_addr(c, acc, A + k * FQ2W, L, false);
    f2_mac_addr(c, acc, A + k2 * FQ2W, L + FQ2W, w2);
    f2_mac_addr(c, acc, A + k3 * FQ2W, L + 2 * FQ2W, w3);
  } else {
    f2_mac(c, acc, ld2(A + k * FQ2W), ld2(L));
    Fq2 d = ld2(L + FQ2W);
    f2_mac(c, acc, ld2(A + k2 * FQ2W), f2sel(w2, f2xi(d), d));
    d = ld2(L + 2 * FQ2W);
    f2_mac(c, acc, ld2(A + k3 * FQ2W), f2sel(w3, f2xi(d), d));
  }
  Fq2 out = f2_finish(c, acc);
  sync(c);
  st2(freg(c, dst) + k * FQ2W, out);
  sync(c);
}
// dst = conj(a)  (w -> -w)
template <class C>
RIPP_HD void conj(const C& c, int dst, int a) {
  Fq2 v = ld2(freg(c, a) + c.k * FQ2W);
  v = f2sel(c.k & 1, f2neg(v), v);
  sync(c);
  st2(freg(c, dst) + c.k * FQ2W, v);
  sync(c);
}
template <class C>
RIPP_HD void copy(const C& c, int dst, int a) {
  Fq2 v = ld2(freg(c, a) + c.k * FQ2W);
  sync(c);
  st2(freg(c, dst) + c.k * FQ2W, v);
  sync(c);
}
template <class C>
RIPP_HD void set_one(const C& c, int dst) {
  st2(freg(c, dst) + c.k * FQ2W, f2sel(c.k == 0, Fq2::one(), Fq2::zero()));
  sync(c);
}
RIPP_HD Fq2 frob_gamma(int npow, int k) {
  Fq2 g;
#pragma unroll 1
  for (int i = 0; i < 12; i++) {
    g.c0.v[i] = npow == 1 ? k::FROB1(24 * k + i) : k::FROB2(24 * k + i);
    g.c1.v[i] = npow == 1 ? k::FROB1(24 * k + 12 + i) : k::FROB2(24 * k + 12 + i);
  }
  return g;
}
// dst = a^(p^npow), npow in {1, 2}: coefficient-local
template <class C>
RIPP_HD void frob(const C& c, int dst, int a, int npow) {
  Fq2 v = ld2(freg(c, a) + c.k * FQ2W);
  v = f2sel(npow & 1, f2conj(v), v);
  v = f2mul(c, v, frob_gamma(npow, c.k));
  sync(c);
  st2(freg(c, dst) + c.k * FQ2W, v);
  sync(c);
}
// dst = a^-1 via the norm to Fq2: N = a conj(a) in Fq6, d = N N^(p^2) N^(p^4) in Fq2, a^-1 = conj(a) N^(p^2) N^(p^4) / d.
// Clobbers registers t0, t1, t2 (all distinct from a and dst).
template <class C>
RIPP_HD void inv(const C& c, int dst, int a, int t0, int t1, int t2) {
  conj(c, t0, a);
  mul(c, t1, a, t0);      // N
  frob(c, t2, t1, 2);     // N^(p^2)
  frob(c, dst, t2, 2);    // N^(p^4)
  mul(c, t2, t2, dst);    // T = N^(p^2) N^(p^4)
  mul(c, dst, t1, t2);    // d = N T, only the w^0 coefficient is non-zero
  Fq2 d = ld2(freg(c, dst));
  Fq dn = (fqmul(d.c0, d.c0) + fqmul(d.c1, d.c1)).inv();
  Fq2 dinv = {fqmul(d.c0, dn), -fqmul(d.c1, dn)};
  Fq2 tk = f2mul(c, ld2(freg(c, t2) + c.k * FQ2W), dinv);
  sync(c);
  st2(freg(c, t2) + c.k * FQ2W, tk);  // N^-1
  sync(c);
  mul(c, dst, t0, t2);
}

// ---- address-driven lazy sums (W = 3) --------------------------------------------------------------
// On the eighteen-lane shape the modular additions around a product, not the product, are the chain: the
// Granger-Scott squaring below spent 31 of them (~37 instructions each: carry chain, trial subtraction, select) and
// twelve 24-word selects on one 330-instruction product.  Here a lane forms a SUM of values read from lane-dependent
// shared-memory ADDRESSES (padding slots read twelve zero words), carries it unreduced in 12 or 13 limbs
// (2^384 = 9.84 p), and reduces once: operands of a product only need (bound of u) x (bound of v) < 9.84 p^2 for the
// Montgomery product to come out below 2p; results are brought to [0, p) by one quotient estimate.
RIPP_DEFCONST(FQ_2P, 12, 0xffff5556u, 0x73fdffffu, 0x62a7ffffu, 0x3d57fffdu, 0xed61ec48u, 0xce61a541u, 0xe70a257eu, 0xc8ee9709u, 0x869759aeu, 0x96374f6cu, 0x72ffcd34u, 0x340223d4u)
RIPP_DEFCONST(FQ_3P, 12, 0xffff0001u, 0x2dfcffffu, 0x13fbffffu, 0x5c03fffcu, 0xe412e26cu, 0x359277e2u, 0xda8f383eu, 0x2d65e28eu, 0xc9e30686u, 0xe152f722u, 0xac7fb3ceu, 0x4e0335beu)
RIPP_DEFCONST(FQ_4P, 12, 0xfffeaaacu, 0xe7fbffffu, 0xc54ffffeu, 0x7aaffffau, 0xdac3d890u, 0x9cc34a83u, 0xce144afdu, 0x91dd2e13u, 0x0d2eb35du, 0x2c6e9ed9u, 0xe5ff9a69u, 0x680447a8u)
RIPP_DEFCONST(FQ_5P, 12, 0xfffe5557u, 0xa1faffffu, 0x76a3fffeu, 0x995bfff9u, 0xd174ceb4u, 0x03f41d24u, 0xc1995dbdu, 0xf6547998u, 0x507a6034u, 0x778a468fu, 0x1f7f8103u, 0x82055993u)
template <int K>
RIPP_HD uint32_t fq_kp(int i) {
  static_assert(K >= 1 && K <= 5, "tabulated multiples of p");
  return K == 1 ? FqParams::p(i) : (K == 2 ? FQ_2P(i) : (K == 3 ? FQ_3P(i) : (K == 4 ? FQ_4P(i) : FQ_5P(i))));
}
template <int K>
RIPP_HD void lz_addk12(uint32_t* acc) {  // acc += K p
  using namespace limb;
  add_cc(acc[0], acc[0], fq_kp<K>(0));
#pragma unroll
  for (int i = 1; i < 11; i++) addc_cc(acc[i], acc[i], fq_kp<K>(i));
  addc(acc[11], acc[11], fq_kp<K>(11));
}
// v >= K p ? v - K p : v
template <int K>
RIPP_HD void lz_csub12(uint32_t* v) {
  using namespace limb;
  uint32_t t[12], borrow;
  sub_cc(t[0], v[0], fq_kp<K>(0));
#pragma unroll
  for (int i = 1; i < 12; i++) subc_cc(t[i], v[i], fq_kp<K>(i));
  subc(borrow, 0, 0);
#pragma unroll
  for (int i = 0; i < 12; i++) v[i] = borrow ? v[i] : t[i];
}
// 13-limb accumulator (values below 32 p < 2^386)
RIPP_HD void lz_add13(uint32_t* acc, const uint32_t* x, uint32_t x12) {
  using namespace limb;
  add_cc(acc[0], acc[0], x[0]);
#pragma unroll
  for (int i = 1; i < 12; i++) addc_cc(acc[i], acc[i], x[i]);
  addc(acc[12], acc[12], x12);
}
RIPP_HD void lz_sub13(uint32_t* acc, const uint32_t* x) {
  using namespace limb;
  sub_cc(acc[0], acc[0], x[0]);
#pragma unroll
  for (int i = 1; i < 12; i++) subc_cc(acc[i], acc[i], x[i]);
  subc(acc[12], acc[12], 0);
}
// T (13 limbs, T < 2^386 = 39.4 p) -> T mod p.  h = floor(T / 2^376) < 1024 and p / 2^376 = 26.0042..., so
// q = floor(h * 2520 / 2^16) (2520 / 2^16 = 1 / 26.0063) is floor(T / p) or one less for every such T (checked
// exhaustively over h in tests/test_hostsim.py); T - q p < 2p fits 12 limbs and one trial subtraction finishes.
RIPP_HD Fq lz_reduce13(const uint32_t* T) {
  using namespace limb;
  const uint32_t h = (T[12] << 8) | (T[11] >> 24);
  const uint32_t q = (h * 2520u) >> 16;
  uint32_t lo[12], hi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint64_t m = mul_wide(q, FqParams::p(i));
    lo[i] = (uint32_t)m;
    hi[i] = (uint32_t)(m >> 32);
  }
  Fq r;
  sub_cc(r.v[0], T[0], lo[0]);
#pragma unroll
  for (int i = 1; i < 11; i++) subc_cc(r.v[i], T[i], lo[i]);
  subc(r.v[11], T[11], lo[11]);
  sub_cc(r.v[1], r.v[1], hi[0]);
#pragma unroll
  for (int i = 2; i < 11; i++) subc_cc(r.v[i], r.v[i], hi[i - 1]);
  subc(r.v[11], r.v[11], hi[10]);
  detail::final_sub<FqParams>(r.v);
  return r;
}

// One lane's share of a "pass": out = MUL * (sum of NP slots - sum of NM slots) [/ 2] mod p, fully reduced.  The slots
// are twelve-word canonical values at lane-dependent shared-memory addresses (unused ones point at the zero slot), so
// lanes producing DIFFERENT derived values run one instruction stream; NP + NM + 1 <= 9 keeps the sum in 12 limbs.
RIPP_HD uint32_t shr_pair(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, sh);
#else
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> sh);
#endif
}
template <int NP, int NM, int MUL, bool HALF>
RIPP_HD Fq lz_pass(const uint32_t* base, const int* P, const int* M, bool half) {
  static_assert(NP + NM + 1 <= 9 && (MUL == 1 || (MUL == 12 && NP + NM <= 3 && !HALF)), "bounds: 2^384 = 9.84 p, 2^386 = 39.4 p");
  using namespace limb;
  uint32_t S[13];
  {
    Fq t = ld1(base + P[0]);
    lz_addk12<NM>(t.v);  // + NM p: the subtractions below cannot go negative
#pragma unroll
    for (int i = 0; i < 12; i++) S[i] = t.v[i];
  }
#pragma unroll
  for (int j = 1; j < NP; j++) lz_add12(S, ld1(base + P[j]));
#pragma unroll
  for (int j = 0; j < NM; j++) lz_sub12(S, ld1(base + M[j]));
  if (HALF) {  // (S + (S odd ? p : 0)) / 2 for the lanes that halve
    const uint32_t m = (half && (S[0] & 1u)) ? 0xffffffffu : 0u, sh = half ? 1u : 0u;
    add_cc(S[0], S[0], FqParams::p(0) & m);
#pragma unroll
    for (int i = 1; i < 11; i++) addc_cc(S[i], S[i], FqParams::p(i) & m);
    addc(S[11], S[11], FqParams::p(11) & m);
#pragma unroll
    for (int i = 0; i < 11; i++) S[i] = shr_pair(S[i], S[i + 1], sh);
    S[11] >>= sh;
  }
  S[12] = 0;
  if (MUL == 12) {
    uint32_t T[13];
#pragma unroll
    for (int i = 0; i < 13; i++) T[i] = S[i];
    lz_add13(T, S, 0);
    lz_add13(T, S, 0);
#pragma unroll
    for (int i = 12; i >= 1; i--) S[i] = (T[i] << 2) | (T[i - 1] >> 30);
    S[0] = T[0] << 2;
  }
  return lz_reduce13(S);
}

// Granger-Scott squaring on eighteen lanes (same formulas as cyc_sqr_body below).  Lane (k, role) multiplies ONE pair of
// operand sums u v -- role 0: u0 v0, role 1: u1 v1, role 2: (u0 + u1)(v0 + v1) of this coefficient's Fq2 product, with
// (u, v) = (ra, rb) for k < 3 and (ra + rb, ra + xi rb) for k >= 3 -- and after the exchange forms ONE component
// (role 1: c1, roles 0 and 2: c0) of its output coefficient straight from the six raw Karatsuba parts
// p0 p1 p2 (product ra rb) and x0 x1 x2 (cross product) of its source pair:
//   t_even = (x0 - x1 - 3 p0 + p1 + p2,  x2 - x0 - x1 - 2 p2 + p0 + 3 p1)
//   t_odd  = (2 p0 - 2 p1,  2 p2 - 2 p0 - 2 p1),      xi t_odd = (4 p0 - 2 p2,  2 p2 - 4 p1)
//   out = 3 t -+ 2 own        (t_even, -: k = 0, 2, 4;  t_odd, +: k = 3, 5;  xi t_odd, +: k = 1)
// as S = 5p + (five added slots) - (four subtracted slots) in [p, 10p), T = 3 S + 2 own_added - 2 own_subtracted < 32 p.
template <class C>
RIPP_HD void cyc_sqr_w3_body(const C& c, int dst, int a) {
  const uint32_t* A = freg(c, a);
  const uint32_t* Z = c.bus + BUS_ZERO;
  const int k = c.k, pr = k % 3, role = c.role;
  const bool hi = k >= 3;
  const uint32_t *ra0 = A + pr * FQ2W, *ra1 = ra0 + 12, *rb0 = A + (pr + 3) * FQ2W, *rb1 = rb0 + 12;
  // u:  k < 3: ra0 | ra1 | ra0 + ra1;   k >= 3: ra0 + rb0 | ra1 + rb1 | ra0 + ra1 + rb0 + rb1      (< 4p, then < 2p)
  // v:  k < 3: rb0 | rb1 | rb0 + rb1;   k >= 3: ra0 + rb0 - rb1 | ra1 + rb0 + rb1 | ra0 + ra1 + 2 rb0   (+ p; < 5p, then < 4p)
  const uint32_t* U1 = role == 1 ? ra1 : ra0;
  const uint32_t* U2 = role == 2 ? ra1 : Z;
  const uint32_t* U3 = hi ? (role == 1 ? rb1 : rb0) : Z;
  const uint32_t* U4 = (hi && role == 2) ? rb1 : Z;
  const uint32_t* V1 = hi ? (role == 1 ? ra1 : ra0) : (role == 1 ? rb1 : rb0);
  const uint32_t* V2 = hi ? rb0 : (role == 2 ? rb1 : Z);
  const uint32_t* V3 = hi ? (role == 0 ? Z : (role == 1 ? rb1 : ra1)) : Z;
  const uint32_t* V4 = (hi && role == 2) ? rb0 : Z;
  const uint32_t* N1 = (hi && role == 0) ? rb1 : Z;
  Fq u = ld1(U1), v = ld1(V1);
  lz_add12(u.v, ld1(U2));
  lz_add12(u.v, ld1(U3));
  lz_add12(u.v, ld1(U4));
  lz_csub12<2>(u.v);
  lz_add12(v.v, ld1(V2));
  lz_add12(v.v, ld1(V3));
  lz_add12(v.v, ld1(V4));
  lz_addk12<1>(v.v);
  lz_sub12(v.v, ld1(N1));
  lz_csub12<4>(v.v);
  uint32_t* b = c.bus + c.par * BUS_WORDS;
  c.par ^= 1;
  st1(b + k * 36 + role * 12, fqmul(u, v));  // < (2p)(4p) / R + p < 2p before the product's own final subtraction
  sync(c);
  // source pair of this coefficient's output: a0, a3 <- pair 0; a2, a5 <- pair 1; a1, a4 <- pair 2
  const int src = pr == 0 ? 0 : (pr == 2 ? 1 : 2);
  const uint32_t *p0 = b + src * 36, *p1 = p0 + 12, *p2 = p0 + 24;
  const uint32_t *x0 = b + (src + 3) * 36, *x1 = x0 + 12, *x2 = x0 + 24;
  const bool c1 = role == 1;
  const bool tE = (k & 1) == 0, tX = k == 1;  // else t_odd (k = 3, 5)
  const uint32_t *P1, *P2, *P3, *P4, *P5, *M1, *M2, *M3, *M4;
  //            even, c0        even, c1        xi odd, c0   xi odd, c1   odd, c0      odd, c1
  // added      x0 p1 p2 .  .   x2 p0 p1 p1 p1  p0 p0 p0 p0  p2 p2 . . .  p0 p0 . . .  p2 p2 . . .
  // subtracted x1 p0 p0 p0     x0 x1 p2 p2     p2 p2 .  .   p1 p1 p1 p1  p1 p1 .  .   p0 p0 p1 p1
  P1 = tE ? (c1 ? x2 : x0) : (c1 ? p2 : p0);
  P2 = tE ? (c1 ? p0 : p1) : (c1 ? p2 : p0);
  P3 = tE ? (c1 ? p1 : p2) : ((tX && !c1) ? p0 : Z);
  P4 = (tE && c1) ? p1 : ((tX && !c1) ? p0 : Z);
  P5 = (tE && c1) ? p1 : Z;
  M1 = tE ? (c1 ? x0 : x1) : (tX ? (c1 ? p1 : p2) : (c1 ? p0 : p1));
  M2 = tE ? (c1 ? x1 : p0) : (tX ? (c1 ? p1 : p2) : (c1 ? p0 : p1));
  M3 = tE ? (c1 ? p2 : p0) : (c1 ? p1 : Z);
  M4 = tE ? (c1 ? p2 : p0) : (c1 ? p1 : Z);
  const uint32_t* own = A + k * FQ2W + (c1 ? 12 : 0);
  const uint32_t* ownP = tE ? Z : own;
  const uint32_t* ownM = tE ? own : Z;
  uint32_t S[13];
#pragma unroll
  for (int i = 0; i < 12; i++) S[i] = fq_kp<5>(i);
  S[12] = 0;
  {
    Fq t;
    t = ld1(P1); lz_add13(S, t.v, 0);
    t = ld1(P2); lz_add13(S, t.v, 0);
    t = ld1(P3); lz_add13(S, t.v, 0);
    t = ld1(P4); lz_add13(S, t.v, 0);
    t = ld1(P5); lz_add13(S, t.v, 0);
    t = ld1(M1); lz_sub13(S, t.v);
    t = ld1(M2); lz_sub13(S, t.v);
    t = ld1(M3); lz_sub13(S, t.v);
    t = ld1(M4); lz_sub13(S, t.v);
  }
  uint32_t T[13];
#pragma unroll
  for (int i = 0; i < 13; i++) T[i] = S[i];
  lz_add13(T, S, S[12]);
  lz_add13(T, S, S[12]);
  {
    Fq t = ld1(ownP);
    lz_add13(T, t.v, 0);
    lz_add13(T, t.v, 0);
    t = ld1(ownM);
    lz_sub13(T, t.v);
    lz_sub13(T, t.v);
  }
  Fq out = lz_reduce13(T);
  sync(c);  // every lane has read `a` (dst may be the same register) before anyone overwrites it
  if (role != 2) st1(freg(c, dst) + k * FQ2W + (c1 ? 12 : 0), out);
  sync(c);
}

// dst = a^2 for a in the cyclotomic subgroup (Granger-Scott): ONE Fq2 product per coefficient.
// In the flat basis the three Fq4 pairs are (a0, a3), (a1, a4), (a2, a5); coefficient k < 3 forms a_k a_{k+3},
// coefficient k >= 3 forms (a_{k-3} + a_k)(a_{k-3} + xi a_k), and with
//   t_even(pair) = (ra + rb)(ra + xi rb) - (1 + xi) ra rb,  t_odd(pair) = 2 ra rb
// the outputs are  a0' = 3 t_even(A) - 2 a0,  a3' = 3 t_odd(A) + 2 a3,  a1' = 3 xi t_odd(C) + 2 a1,
// a4' = 3 t_even(C) - 2 a4,  a2' = 3 t_even(B) - 2 a2,  a5' = 3 t_odd(B) + 2 a5   (A, B, C = pairs 0, 1, 2).
template <class C>
RIPP_HD void cyc_sqr_body(const C& c, int dst, int a) {
  const uint32_t* A = freg(c, a);
  uint32_t* R = c.sm + OFF_R;
  const int k = c.k, pr = k % 3;
  Fq2 ra = ld2(A + pr * FQ2W), rb = ld2(A + (pr + 3) * FQ2W);
  Fq2 u = f2sel(k < 3, ra, f2add(ra, rb));
  Fq2 v = f2sel(k < 3, rb, f2add(ra, f2xi(rb)));
  Fq2 own = ld2(A + k * FQ2W);
  st2(R + k * FQ2W, f2mul(c, u, v));
  sync(c);
  // source pair of this coefficient's output: a0,a3 <- A(0); a2,a5 <- B(1); a1,a4 <- C(2)
  const int src = (k % 3 == 0) ? 0 : (k % 3 == 2 ? 1 : 2);
  Fq2 prod = ld2(R + src * FQ2W), cross = ld2(R + (src + 3) * FQ2W);
  Fq2 t_even = f2sub(cross, f2add(prod, f2xi(prod)));
  Fq2 t_odd = f2dbl(prod);
  // k = 0, 2, 4 take 3 t_even - 2 own; k = 3, 5 take 3 t_odd + 2 own; k = 1 takes 3 xi t_odd + 2 own
  Fq2 t = f2sel(k == 0 || k == 2 || k == 4, t_even, f2sel(k == 1, f2xi(t_odd), t_odd));
  Fq2 s2 = f2dbl(own);
  Fq2 three_t = f2add(f2dbl(t), t);
  Fq2 out = f2sel(k == 0 || k == 2 || k == 4, f2sub(three_t, s2), f2add(three_t, s2));
  sync(c);
  st2(freg(c, dst) + k * FQ2W, out);
  sync(c);
}

// -DRIPP_L6_CALL_W3: the eighteen-lane shape's Fq12 operations as real calls (one copy each) instead of inlined.  ncu
// shows 14 % of k_final_exp18's stall samples as "no instruction" (30 k instructions inlined), but the calls measured
// no better: final exponentiation 1.39 -> 1.44 ms, k_miller18<1> 1.12 -> 1.21 ms (gpurun_out r2g).  Kept for A/B runs.
#if defined(__CUDA_ARCH__) && defined(RIPP_L6_CALL_W3)
#define RIPP_L6_OP(name, params, args)                                                              \
  static __device__ __noinline__ void name##_w3(const Ctx3& c, params) { name##_body(c, args); }    \
  template <class C>                                                                                \
  RIPP_HD void name(const C& c, params) {                                                           \
    if constexpr (C::W == 3) name##_w3(c, args); else name##_body(c, args);                         \
  }
#else
#define RIPP_L6_OP(name, params, args) \
  template <class C>                   \
  RIPP_HD void name(const C& c, params) { name##_body(c, args); }
#endif
#define RIPP_L6_COMMA ,
RIPP_L6_OP(sqr, int dst RIPP_L6_COMMA int a, dst RIPP_L6_COMMA a)
RIPP_L6_OP(mul_line, int dst RIPP_L6_COMMA int a, dst RIPP_L6_COMMA a)
#if defined(__CUDA_ARCH__) && defined(RIPP_L6_CALL_W3)
RIPP_L6_OP(cyc_sqr, int dst RIPP_L6_COMMA int a, dst RIPP_L6_COMMA a)
#else
// -DRIPP_L6_CYC_EAGER: the eighteen-lane shape on the generic body (the round-2 baseline, kept for A/B runs)
template <class C>
RIPP_HD void cyc_sqr(const C& c, int dst, int a) {
#if !defined(RIPP_L6_CYC_EAGER)
  if constexpr (C::W == 3) {
    cyc_sqr_w3_body(c, dst, a);
    return;
  }
#endif
  cyc_sqr_body(c, dst, a);
}
#endif

// a^x (x = -|x|) for a in the cyclotomic subgroup; dst != a; clobbers nothing else
template <class C>
RIPP_HD void exp_by_x(const C& c, int dst, int a) {
  copy(c, dst, a);
#pragma unroll 1
  for (int i = 62; i >= 0; i--) {
    cyc_sqr(c, dst, dst);
    if ((k::X_ABS >> i) & 1) mul(c, dst, dst, a);
  }
  conj(c, dst, dst);
}

// acc = a^e for ANY a in Fq12 (generic squarings: verifier inputs are not trusted to be cyclotomic) and a
// 256-bit canonical exponent e (8 words in the group's scratch).  `PairingOutput *= Fr` of the verifiers
// (mul_helper on GT, gipa.rs:355-357; sipp/src/lib.rs:148-156).  Fixed 2-bit windows; the window digit
// selects an operand ADDRESS (1, a, a^2, a^3), so groups with different exponents share one instruction stream.
// a in register t1; clobbers t2, t3, one.
template <class C>
RIPP_HD void pow_fr(const C& c, int acc, int t1, int t2, int t3, int one, const uint32_t* e) {
  set_one(c, one);
  set_one(c, acc);
  sqr(c, t2, t1);
  mul(c, t3, t2, t1);
#pragma unroll 1
  for (int i = 127; i >= 0; i--) {
    sqr(c, acc, acc);
    sqr(c, acc, acc);
    uint32_t w = (e[i >> 4] >> ((i & 15) * 2)) & 3u;
    int src = w == 0 ? one : (w == 1 ? t1 : (w == 2 ? t2 : t3));
    mul(c, acc, acc, src);
  }
}

// Register 0 <- final_exponentiation(register 0) (ark-ec convention, see pairing.cuh); needs 8 registers.
template <class C>
RIPP_HD void final_exp(const C& c) {
  enum { F = 0, R = 1, Y0 = 2, Y1 = 3, Y2 = 4, T0 = 5, T1 = 6, T2 = 7 };
  inv(c, R, F, T0, T1, T2);   // R = f^-1
  conj(c, Y0, F);
  mul(c, R, Y0, R);           // f^(p^6 - 1)
  frob(c, Y0, R, 2);
  mul(c, R, Y0, R);           // ^(p^2 + 1)
  cyc_sqr(c, Y0, R);          // y0 = r^2
  exp_by_x(c, Y1, R);         // y1 = r^x
  conj(c, Y2, R);
  mul(c, Y1, Y1, Y2);
  exp_by_x(c, Y2, Y1);
  conj(c, Y1, Y1);
  mul(c, Y1, Y1, Y2);
  exp_by_x(c, Y2, Y1);
  frob(c, Y1, Y1, 1);
  mul(c, Y1, Y1, Y2);
  mul(c, R, R, Y0);
  exp_by_x(c, Y0, Y1);
  exp_by_x(c, Y2, Y0);
  frob(c, Y0, Y1, 2);
  conj(c, Y1, Y1);
  mul(c, Y1, Y1, Y2);
  mul(c, Y1, Y1, Y0);
  mul(c, F, R, Y1);
}

// ---- Miller loop --------------------------------------------------------------------------------
// smem: T = (x, y, z) at PB_T, (xP, -yP) at PB_P, accumulator in register 0.
RIPP_HD void stage_pair_p(uint32_t* pb, const G1Aff& P) { st2(pb + PB_P, Fq2{P.x, -P.y}); }
template <class C>
RIPP_HD int lane_index(const C& c) { return C::W == 1 ? c.k : c.k * 3 + c.role; }
template <class C>
RIPP_HD bool lane_stores(const C& c) { return C::W == 1 || c.role == 0; }  // W = 3: the triple holds one value

// One doubling step (ark-ec's homogeneous-projective formulas): two rounds of six parallel Fq2 products, every operand
// read from a lane-dependent ADDRESS, and between / after them three lz_pass rounds in which each lane forms one Fq
// component of one derived value.  (The first version had every lane compute every intermediate and select its own:
// 48 modular additions and ~22 24-word selects per step -- 2300 of the ~9100 instructions of a Miller-loop bit.)
//   round 1:  R0 = x y, R1 = y^2 (= b), R2 = z^2, R3 = y z, R4 = x^2
//   pass E:   E = 12 xi R2                        (e = 4 xi 3 z^2; f = 3 E)                        -> slot R5
//   pass B:   BFh = (R1 - 3E) / 2,  G = (R1 + 3E) / 2,  H = 2 R3 (= h = (y + z)^2 - y^2 - z^2)     -> the line slots
//   round 2:  x' = R0 BFh,  R2 = G^2,  z' = R1 H,  R3 = E^2,  R4 = R4 xP,  d4 = H (-yP)
//   pass F:   y' = R2 - 3 R3,  d0 = E - R1,  d1 = 3 R4         (masked pair: the line is the constant 1)
template <class C>
RIPP_HD void dbl_step(const C& c, uint32_t* pb) {
  const int k = c.k, li = lane_index(c);
  uint32_t* sm = c.sm;
  const int oT = (int)(pb - sm) + PB_T, oP = (int)(pb - sm) + PB_P, oZ = (int)(c.zero - sm);
  constexpr int oR = OFF_R, oL = OFF_LINE, oE = OFF_R + 5 * FQ2W;
  const bool valid = pb[PB_VALID] != 0;
  {
    const int ua = oT + ((k == 0 || k >= 4) ? 0 : (k == 2 ? 2 : 1)) * FQ2W;
    const int va = oT + (k <= 1 ? 1 : (k <= 3 ? 2 : 0)) * FQ2W;
    Fq2 r = f2mul(c, ld2(sm + ua), ld2(sm + va));
    if (lane_stores(c)) st2(sm + oR + k * FQ2W, r);
  }
  sync(c);
  {
    const int cc = li & 1;
    const int P[2] = {oR + 2 * FQ2W, cc ? oR + 2 * FQ2W + 12 : oZ}, M[1] = {cc ? oZ : oR + 2 * FQ2W + 12};
    Fq e = lz_pass<2, 1, 12, false>(sm, P, M, false);
    if (li < 2) st1(sm + oE + 12 * cc, e);
  }
  sync(c);
  {
    const int ci = li % 6, vi = ci >> 1, cc = 12 * (ci & 1);
    const int r1 = oR + FQ2W + cc, r3 = oR + 3 * FQ2W + cc, e = oE + cc;
    const int P[4] = {vi == 2 ? r3 : r1, vi == 0 ? oZ : (vi == 1 ? e : r3), vi == 1 ? e : oZ, vi == 1 ? e : oZ};
    const int m = vi == 0 ? e : oZ;
    const int M[3] = {m, m, m};
    Fq v = lz_pass<4, 3, 1, true>(sm, P, M, vi != 2);
    if (li < 6) st1(sm + oL + vi * FQ2W + cc, v);
  }
  sync(c);
  {
    const int ua = k == 0 ? oR : (k == 1 ? oL + FQ2W : (k == 2 ? oR + FQ2W : (k == 3 ? oE : (k == 4 ? oR + 4 * FQ2W : oL + 2 * FQ2W))));
    const int v0 = k == 0 ? oL : (k == 1 ? oL + FQ2W : (k == 2 ? oL + 2 * FQ2W : (k == 3 ? oE : (k == 4 ? oP : oP + 12))));
    const int v1 = k < 4 ? v0 + 12 : oZ;
    Fq2 u = ld2(sm + ua), v = {ld1(sm + v0), ld1(sm + v1)};
    sync(c);  // every operand is in registers before a product lands on its slot
    Fq2 r = f2mul(c, u, v);
    r = f2sel(valid || k < 4, r, Fq2::zero());  // masked pair: d4 = 0 (d1 is zeroed in pass F)
    const int dst = k == 0 ? oT : (k == 1 ? oR + 2 * FQ2W : (k == 2 ? oT + 2 * FQ2W : (k == 3 ? oR + 3 * FQ2W : (k == 4 ? oR + 4 * FQ2W : oL + 2 * FQ2W))));
    if (lane_stores(c)) st2(sm + dst, r);
  }
  sync(c);
  {
    const int ci = li % 6, vi = ci >> 1, cc = 12 * (ci & 1);
    const int r1 = oR + FQ2W + cc, r2 = oR + 2 * FQ2W + cc, r3 = oR + 3 * FQ2W + cc, r4 = oR + 4 * FQ2W + cc, e = oE + cc;
    const int P[3] = {vi == 0 ? r2 : (vi == 1 ? e : r4), vi == 2 ? r4 : oZ, vi == 2 ? r4 : oZ};
    const int M[3] = {vi == 0 ? r3 : (vi == 1 ? r1 : oZ), vi == 0 ? r3 : oZ, vi == 0 ? r3 : oZ};
    Fq v = lz_pass<3, 3, 1, false>(sm, P, M, false);
    if (vi != 0 && !valid) v = (vi == 1 && cc == 0) ? Fq::one() : Fq::zero();
    const int dst = vi == 0 ? oT + FQ2W + cc : (vi == 1 ? oL + cc : oL + FQ2W + cc);
    if (li < 6) st1(sm + dst, v);
  }
  sync(c);
}

// One addition step T <- T + Q with the chord line.
template <class C>
RIPP_HD void add_step(const C& c, uint32_t* pb) {
  uint32_t* T = pb + PB_T;
  uint32_t* R = c.sm + OFF_R;
  const int k = c.k;
  Fq2 x = ld2(T), y = ld2(T + FQ2W), z = ld2(T + 2 * FQ2W);
  const uint32_t* Q = pair_q(pb);
  Fq2 qx = ld2(Q), qy = ld2(Q + FQ2W);
  // round 1: R0 = qy z, R1 = qx z
  st2(R + k * FQ2W, f2mul(c, f2sel(k == 0, qy, qx), z));
  sync(c);
  Fq2 theta = f2sub(y, ld2(R)), lambda = f2sub(x, ld2(R + FQ2W));
  sync(c);
  // round 2: R0 = theta^2, R1 = lambda^2, R2 = theta qx, R3 = lambda qy
  Fq2 u = f2sel(k == 0 || k == 2, theta, lambda);
  Fq2 v = f2sel(k == 0, theta, f2sel(k == 1, lambda, f2sel(k == 2, qx, qy)));
  st2(R + k * FQ2W, f2mul(c, u, v));
  sync(c);
  Fq2 cc = ld2(R), d = ld2(R + FQ2W), jv = f2sub(ld2(R + 2 * FQ2W), ld2(R + 3 * FQ2W));
  sync(c);
  // round 3: R0 = lambda d (= e), R1 = z c (= f), R2 = x d (= g)
  u = f2sel(k == 0, lambda, f2sel(k == 1, z, x));
  v = f2sel(k == 1, cc, d);
  st2(R + k * FQ2W, f2mul(c, u, v));
  sync(c);
  Fq2 e = ld2(R), f = ld2(R + FQ2W), g = ld2(R + 2 * FQ2W);
  Fq2 h = f2sub(f2add(e, f), f2dbl(g));
  sync(c);
  // round 4: R0 = lambda h, R1 = theta (g - h), R2 = e y, R3 = z e, R4 = -theta xP, R5 = (-lambda) (-yP)
  Fq xp = ld2(pb + PB_P).c0, myp = ld2(pb + PB_P).c1;  // xP, -yP
  const bool valid = pb[PB_VALID] != 0;
  Fq2 sxp = {xp, Fq::zero()}, syp = {myp, Fq::zero()};
  u = f2sel(k == 0, lambda, f2sel(k == 1, theta, f2sel(k == 2, e, f2sel(k == 3, z, f2neg(f2sel(k == 4, theta, lambda))))));
  v = f2sel(k == 0, h, f2sel(k == 1, f2sub(g, h), f2sel(k == 2, y, f2sel(k == 3, e, f2sel(k == 4, sxp, syp)))));
  st2(R + k * FQ2W, f2mul(c, u, v));
  sync(c);
  Fq2 nx = ld2(R), ny = f2sub(ld2(R + FQ2W), ld2(R + 2 * FQ2W)), nz = ld2(R + 3 * FQ2W);
  Fq2 d1 = ld2(R + 4 * FQ2W), d4 = ld2(R + 5 * FQ2W);
  sync(c);
  if (k == 0) {
    st2(T, nx);
    st2(T + FQ2W, ny);
    st2(T + 2 * FQ2W, nz);
  }
  if (k == 1) {
    st2(c.sm + OFF_LINE, f2sel(valid, jv, Fq2::one()));
    st2(c.sm + OFF_LINE + FQ2W, f2sel(valid, d1, Fq2::zero()));
    st2(c.sm + OFF_LINE + 2 * FQ2W, f2sel(valid, d4, Fq2::zero()));
  }
  sync(c);
}

// Register 0 <- prod_j f_{|x|,Q_j}(P_j), conjugated, over the npairs pair blocks at `pairs` (P, Q and the
// valid flag staged by the caller; masked pairs must still hold finite points).  One accumulator squaring per
// bit is shared by all the group's pairs.
template <class C>
RIPP_HD void miller(const C& c, uint32_t* pairs, int npairs) {
  if (c.k == 0) {
    for (int j = 0; j < npairs; j++) {
      uint32_t* pb = pairs + j * PAIR_WORDS;
      const uint32_t* Q = pair_q(pb);
      st2(pb + PB_T, ld2(Q));
      st2(pb + PB_T + FQ2W, ld2(Q + FQ2W));
      st2(pb + PB_T + 2 * FQ2W, Fq2::one());
    }
  }
  set_one(c, 0);
#pragma unroll 1
  for (int i = 62; i >= 0; i--) {
    sqr(c, 0, 0);
#pragma unroll 1
    for (int j = 0; j < npairs; j++) {
      dbl_step(c, pairs + j * PAIR_WORDS);
      mul_line(c, 0, 0);
    }
    if ((k::X_ABS >> i) & 1) {
#pragma unroll 1
      for (int j = 0; j < npairs; j++) {
        add_step(c, pairs + j * PAIR_WORDS);
        mul_line(c, 0, 0);
      }
    }
  }
  conj(c, 0, 0);
}

}  // namespace l6
}  // namespace ripp
