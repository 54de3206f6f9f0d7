// Prime-field arithmetic on N x 32-bit limbs, Montgomery form, values always fully reduced in [0, p).
//
// Replaces ark-ff `Fp<MontBackend<_, N>, N>` (third-party) for
// BLS12-381 Fq (N = 12, R = 2^384) and Fr (N = 8, R = 2^256).  Same R as arkworks' 64-bit-limb
// representation, so host Montgomery limbs are ingested unchanged.
//
// Multiplication is the interleaved (CIOS) Montgomery product laid out as two carry chains per
// row ("even"/"odd" accumulators) so that lo/hi halves of 32x32 products never need a carry
// shuffle: 2N^2 + N MAC32 per product (Fq: 300, Fr: 136).
#pragma once
#include "limb.cuh"
#include "constants.cuh"
#include "modinv.cuh"

namespace ripp {

struct FqParams {
  static constexpr int N = 12;
  static constexpr uint32_t M0 = k::FQ_M0;
  RIPP_HD static uint32_t p(int i) { return k::FQ_P(i); }
  RIPP_HD static uint32_t one(int i) { return k::FQ_ONE(i); }
  RIPP_HD static uint32_t r2(int i) { return k::FQ_R2(i); }
  RIPP_HD static uint32_t pm2(int i) { return k::FQ_PM2(i); }
  static constexpr int BITS = 381;
};
struct FrParams {
  static constexpr int N = 8;
  static constexpr uint32_t M0 = k::FR_M0;
  RIPP_HD static uint32_t p(int i) { return k::FR_P(i); }
  RIPP_HD static uint32_t one(int i) { return k::FR_ONE(i); }
  RIPP_HD static uint32_t r2(int i) { return k::FR_R2(i); }
  RIPP_HD static uint32_t pm2(int i) { return k::FR_PM2(i); }
  static constexpr int BITS = 255;
};

namespace detail {
using namespace limb;

// r in [0, 2p)  ->  [0, p)
template <class P>
RIPP_HD void final_sub(uint32_t* r) {
  constexpr int N = P::N;
  uint32_t t[N], borrow;
  sub_cc(t[0], r[0], P::p(0));
#pragma unroll
  for (int i = 1; i < N; i++) subc_cc(t[i], r[i], P::p(i));
  subc(borrow, 0, 0);  // 0 or 0xffffffff
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = borrow ? r[i] : t[i];
}

// One row of the interleaved product after the first.  X is the aligned accumulator, Y the one
// offset by a limb; both are held as N/2 aligned (lo, hi) register pairs (DESIGN.md "Montgomery").
template <class P>
RIPP_HD void mont_row(uint64_t* X, uint64_t* Y, const uint32_t* a, uint32_t bi) {
  constexpr int H = P::N / 2;
  add_lo_hi_cc(X[0], Y[0]);
#pragma unroll
  for (int k = 0; k < H - 1; k++) mwc_cc(Y[k], a[2 * k + 1], bi, Y[k + 1]);
  mwc(Y[H - 1], a[2 * H - 1], bi, 0);
  mw_cc(X[0], a[0], bi, X[0]);
#pragma unroll
  for (int k = 1; k < H; k++) mwc_cc(X[k], a[2 * k], bi, X[k]);
  addc_hi(Y[H - 1]);
}

// Add m*p with m chosen so the lowest limb of X cancels.  CIN: the carry flag left by a preceding
// add_lo_hi_cc (fold of the retired pair's high limb into X's lowest limb) enters the offset chain, whose
// lowest position is exactly one limb above the fold.
template <class P, bool CIN = false>
RIPP_HD void mont_redc(uint64_t* X, uint64_t* Y) {
  constexpr int H = P::N / 2;
  uint32_t m = mul_lo((uint32_t)X[0], P::M0);
  if (CIN)
    mwc_cc(Y[0], P::p(1), m, Y[0]);
  else
    mw_cc(Y[0], P::p(1), m, Y[0]);
#pragma unroll
  for (int k = 1; k < H; k++) mwc_cc(Y[k], P::p(2 * k + 1), m, Y[k]);
  mw_cc(X[0], P::p(0), m, X[0]);
#pragma unroll
  for (int k = 1; k < H; k++) mwc_cc(X[k], P::p(2 * k), m, X[k]);
  addc_hi(Y[H - 1]);
}

// ---- lazy (unreduced) arithmetic for sums of products -----------------------------------------------
// wide_mul: t[0..2N) = a * b over the integers (operands any N-limb values), same even/odd carry chains as
// the Montgomery product minus the reduction rows.  redc_wide: T (2N limbs, T < 2^(32 N) * k p) -> T / R mod p,
// fully reduced with up to `subs` conditional subtractions.  Summing several products before one reduction is
// what makes schoolbook Fq12 arithmetic over six lanes as cheap as the Karatsuba tower (l6.cuh).
template <class P>
RIPP_HD void wide_mul(uint32_t* t, const uint32_t* a, const uint32_t* b) {
  constexpr int N = P::N;
  constexpr int H = N / 2;
  uint64_t ev[H], od[H];
#pragma unroll
  for (int k = 0; k < H; k++) {
    ev[k] = mul_wide(a[2 * k], b[0]);
    od[k] = mul_wide(a[2 * k + 1], b[0]);
  }
  t[0] = (uint32_t)ev[0];
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    mont_row<P>(od, ev, a, b[i]);
    t[i] = (uint32_t)od[0];
    if (i + 1 < N) {
      mont_row<P>(ev, od, a, b[i + 1]);
      t[i + 1] = (uint32_t)ev[0];
    }
  }
  // window now starts at limb N-1 with od aligned (its lowest limb already emitted): t[N + k] = ev[k] + od[k+1]
  uint32_t e[N], o[N];
#pragma unroll
  for (int k = 0; k < H; k++) {
    e[2 * k] = (uint32_t)ev[k];
    e[2 * k + 1] = (uint32_t)(ev[k] >> 32);
    o[2 * k] = (uint32_t)od[k];
    o[2 * k + 1] = (uint32_t)(od[k] >> 32);
  }
  add_cc(t[N], e[0], o[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) addc_cc(t[N + k], e[k], o[k + 1]);
  addc(t[2 * N - 1], e[N - 1], 0);
}
template <int NW>
RIPP_HD void wide_add(uint32_t* acc, const uint32_t* t) {
  add_cc(acc[0], acc[0], t[0]);
#pragma unroll
  for (int k = 1; k < NW - 1; k++) addc_cc(acc[k], acc[k], t[k]);
  addc(acc[NW - 1], acc[NW - 1], t[NW - 1]);
}
template <int NW>
RIPP_HD void wide_sub(uint32_t* acc, const uint32_t* t) {
  sub_cc(acc[0], acc[0], t[0]);
#pragma unroll
  for (int k = 1; k < NW - 1; k++) subc_cc(acc[k], acc[k], t[k]);
  subc(acc[NW - 1], acc[NW - 1], t[NW - 1]);
}
template <class P>
RIPP_HD void redc_wide(uint32_t* r, const uint32_t* t, int subs) {
  constexpr int N = P::N;
  constexpr int H = N / 2;
  uint64_t X[H], Y[H];
#pragma unroll
  for (int k = 0; k < H; k++) {
    X[k] = (uint64_t)t[2 * k] | ((uint64_t)t[2 * k + 1] << 32);
    Y[k] = 0;
  }
  // N rounds: add m p so the lowest limb cancels, slide the window by one limb, inject the next high limb.
  // (A aligned, B offset); after a round the roles swap.  The carry of the fold rides into the next round.
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint64_t* A = (i & 1) ? Y : X;
    uint64_t* B = (i & 1) ? X : Y;
    if (i == 0)
      mont_redc<P, false>(A, B);
    else
      mont_redc<P, true>(A, B);
    uint64_t a0 = A[0];
#pragma unroll
    for (int k = 0; k < H - 1; k++) A[k] = A[k + 1];
    A[H - 1] = (uint64_t)t[N + i];
    add_lo_hi_cc(B[0], a0);  // B.limb0 += a0.limb1, carry pending for the next chain
  }
  // N even: X aligned again; value = X + (Y << 32) + (pending carry << 32)
  uint32_t e[N], o[N];
#pragma unroll
  for (int k = 0; k < H; k++) {
    e[2 * k] = (uint32_t)X[k];
    e[2 * k + 1] = (uint32_t)(X[k] >> 32);
    o[2 * k] = (uint32_t)Y[k];
    o[2 * k + 1] = (uint32_t)(Y[k] >> 32);
  }
  r[0] = e[0];
#pragma unroll
  for (int k = 1; k < N - 1; k++) addc_cc(r[k], e[k], o[k - 1]);
  addc(r[N - 1], e[N - 1], o[N - 2]);
  for (int s2 = 0; s2 < subs; s2++) final_sub<P>(r);
}

#if defined(RIPP_HOSTSIM)
inline thread_local uint64_t mul_count_[2] = {0, 0};  // [Fr, Fq] products, tests only (op-count model)
#endif

template <class P>
RIPP_HD void mont_mul_chain(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = P::N;
  constexpr int H = N / 2;
  static_assert(N % 2 == 0, "even limb count");
#if defined(RIPP_HOSTSIM) && !defined(RIPP_FP_UNSATURATED)
  mul_count_[N == 12]++;
#endif
  uint64_t ev[H], od[H];
#pragma unroll
  for (int k = 0; k < H; k++) {
    ev[k] = mul_wide(a[2 * k], b[0]);
    od[k] = mul_wide(a[2 * k + 1], b[0]);
  }
  mont_redc<P>(ev, od);
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    mont_row<P>(od, ev, a, b[i]);
    mont_redc<P>(od, ev);
    if (i + 1 < N) {
      mont_row<P>(ev, od, a, b[i + 1]);
      mont_redc<P>(ev, od);
    }
  }
  // od's lowest limb is 0 now; value / 2^32 = ev[k] + od[k+1] at limb k
  uint32_t e[N], o[N];
#pragma unroll
  for (int k = 0; k < H; k++) {
    e[2 * k] = (uint32_t)ev[k];
    e[2 * k + 1] = (uint32_t)(ev[k] >> 32);
    o[2 * k] = (uint32_t)od[k];
    o[2 * k + 1] = (uint32_t)(od[k] >> 32);
  }
  add_cc(r[0], e[0], o[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) addc_cc(r[k], e[k], o[k + 1]);
  addc(r[N - 1], e[N - 1], 0);
  final_sub<P>(r);
}


// ---- carry-free product on unsaturated limbs -------------------------------------------------------
// IMAD.WIDE.U32 with carry-in/out (the .X form the 32-bit-limb chains above compile to) issues at half
// the rate of a plain IMAD.WIDE.U32 on sm_100 (ripp_bench_imad: 9.2 vs 17.2 TMAC/s) and serialises a
// warp on the carry predicate.  Here the operands are re-sliced into NL = ceil(32 N / W) limbs of W = 28
// bits so that 64-bit column accumulators absorb a whole column of the product AND of the reduction
// (<= 2 NL products < 2^56 each) without any carry: every multiply is an independent full-rate
// IMAD.WIDE.U32.  The Montgomery radix stays R = 2^(32 N): NL - 1 reduction rounds retire W bits each
// and the last one the remaining 32 N - W (NL - 1) bits, so values remain bit-compatible with arkworks.
#ifndef RIPP_LIMB_BITS
#define RIPP_LIMB_BITS 28
#endif
template <class P>
struct LimbW {
  static constexpr int W = RIPP_LIMB_BITS;
  static constexpr int N = P::N;
  static constexpr int NL = (32 * N + W - 1) / W;
  static constexpr int WL = 32 * N - W * (NL - 1);  // bits retired by the last round
  static constexpr uint32_t MW = (1u << W) - 1;
  // a column collects <= NL products of the multiplication and <= NL of the reduction, each < 2^(2W)
  static constexpr bool MID_NORMALISE = (2 * W + 5) > 63 || ((uint64_t)(2 * NL) << (2 * W)) >= ((uint64_t)1 << 63);
  RIPP_HD static uint32_t pw(int j) {  // limb j of the modulus; folds to an immediate after unrolling
    int bit = W * j, w = bit >> 5, sh = bit & 31;
    uint64_t v = P::p(w);
    if (w + 1 < N) v |= (uint64_t)P::p(w + 1) << 32;
    return (uint32_t)(v >> sh) & MW;
  }
  RIPP_HD static void unpack(uint32_t* L, const uint32_t* a) {
#pragma unroll
    for (int j = 0; j < NL; j++) {
      int bit = W * j, w = bit >> 5, sh = bit & 31;
      uint32_t lo = a[w], hi = (w + 1 < N) ? a[w + 1] : 0u;
      uint32_t v = sh == 0 ? lo : ((lo >> sh) | (sh > 32 - W ? (hi << (32 - sh)) : 0u));
      L[j] = v & MW;
    }
  }
};

template <class P>
RIPP_HD void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if !defined(RIPP_FP_UNSATURATED)
  // Default: the carry-chain product.  The carry-free variant below is kept for the record: it is
  // bit-exact (tests pass with -DRIPP_FP_UNSATURATED) but measured 1.5x SLOWER on B200 (Miller 2^16:
  // 65.6 ms vs 40.9 ms; TIPP 2^12: 607 ms vs 401 ms, profiles/README.md): 406 full-rate IMAD.WIDE plus
  // ~370 shift/mask/64-bit-add ALU instructions lose to 300 half-rate IMAD.WIDE.X plus ~30 -- at the
  // occupancies these kernels run at, instruction count, not the .X issue rate, is what matters.
  mont_mul_chain<P>(r, a, b);
  return;
#endif
  using L = LimbW<P>;
  constexpr int N = P::N, NL = L::NL, WL = L::WL, W = L::W;
  constexpr uint32_t MW = L::MW;
#if defined(RIPP_HOSTSIM)
  mul_count_[N == 12]++;
#endif
  uint32_t A[NL], B[NL];
  L::unpack(A, a);
  L::unpack(B, b);
  uint64_t acc[2 * NL + 1];
#pragma unroll
  for (int k = 0; k < 2 * NL + 1; k++) acc[k] = 0;
  // schoolbook columns, no carries
#pragma unroll
  for (int i = 0; i < NL; i++) {
#pragma unroll
    for (int j = 0; j < NL; j++) acc[i + j] += (uint64_t)A[j] * B[i];
  }
  if (L::MID_NORMALISE) {
#pragma unroll
    for (int k = 0; k < 2 * NL; k++) {
      acc[k + 1] += acc[k] >> W;
      acc[k] &= MW;
    }
  }
  // NL - 1 rounds retire W bits each
#pragma unroll
  for (int i = 0; i < NL - 1; i++) {
    uint32_t m = ((uint32_t)acc[i] * P::M0) & MW;
#pragma unroll
    for (int j = 0; j < NL; j++) acc[i + j] += (uint64_t)m * L::pw(j);
    acc[i + 1] += acc[i] >> W;
  }
  // last round retires the remaining WL bits
  {
    constexpr int i = NL - 1;
    uint32_t m = ((uint32_t)acc[i] * P::M0) & ((1u << WL) - 1);
#pragma unroll
    for (int j = 0; j < NL; j++) acc[i + j] += (uint64_t)m * L::pw(j);
  }
  // value = sum_{k >= NL-1} acc[k] 2^(W (k - NL + 1)), divisible by 2^WL; normalise and repack to 32-bit words
#pragma unroll
  for (int k = NL - 1; k < 2 * NL; k++) {
    acc[k + 1] += acc[k] >> W;
    acc[k] &= MW;
  }
  {
    uint64_t buf = acc[NL - 1] >> WL;  // bit buffer holding `have` valid bits
    int have = W - WL;
    int k = NL;
#pragma unroll
    for (int w = 0; w < N; w++) {
#pragma unroll
      for (int t = 0; t < 2; t++) {
        if (have < 32 && k <= 2 * NL) {
          buf |= acc[k] << have;
          have += W;
          k++;
        }
      }
      r[w] = (uint32_t)buf;
      buf >>= 32;
      have -= 32;
    }
  }
  final_sub<P>(r);
}

}  // namespace detail

template <class P>
struct alignas(16) Fp {
  static constexpr int N = P::N;
  using Params = P;
  uint32_t v[P::N];

  RIPP_HD static Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = 0;
    return r;
  }
  RIPP_HD static Fp one() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::one(i);
    return r;
  }
  RIPP_HD bool is_zero() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= v[i];
    return o == 0;
  }
  RIPP_HD bool operator==(const Fp& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= v[i] ^ b.v[i];
    return o == 0;
  }
  RIPP_HD bool operator!=(const Fp& b) const { return !(*this == b); }

  RIPP_HD Fp operator+(const Fp& b) const {
    using namespace limb;
    Fp r;
    add_cc(r.v[0], v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r.v[i], v[i], b.v[i]);
    addc(r.v[N - 1], v[N - 1], b.v[N - 1]);
    detail::final_sub<P>(r.v);
    return r;
  }
  // a - b, plus p if that borrowed.  The add-back chain does not wait for the borrow (it would make the two
  // 12-link carry chains strictly sequential: ~110 cycles on a lone warp): t + p is formed limb by limb right
  // behind t = a - b, and the borrow only selects between the two at the end.
  RIPP_HD Fp operator-(const Fp& b) const {
    using namespace limb;
    Fp r;
    uint32_t t[N], u[N], mask;
    sub_cc(t[0], v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(t[i], v[i], b.v[i]);
    subc(mask, 0, 0);
    add_cc(u[0], t[0], P::p(0));
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(u[i], t[i], P::p(i));
    addc(u[N - 1], t[N - 1], P::p(N - 1));
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = mask ? u[i] : t[i];
    return r;
  }
  RIPP_HD Fp operator-() const { return zero() - *this; }
  RIPP_HD Fp dbl() const { return *this + *this; }
  // a / 2
  RIPP_HD Fp half() const {
    using namespace limb;
    Fp r;
    uint32_t mask = 0u - (v[0] & 1u);
    add_cc(r.v[0], v[0], P::p(0) & mask);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r.v[i], v[i], P::p(i) & mask);
    addc(r.v[N - 1], v[N - 1], P::p(N - 1) & mask);
#pragma unroll
    for (int i = 0; i < N - 1; i++) r.v[i] = (r.v[i] >> 1) | (r.v[i + 1] << 31);
    r.v[N - 1] >>= 1;
    return r;
  }
  // one out-of-line copy of the product; operands and result travel in registers (by value)
  static RIPP_FN Fp mul_fn(Fp a, Fp b) {
    Fp r;
    detail::mont_mul<P>(r.v, a.v, b.v);
    return r;
  }
  RIPP_HD Fp operator*(const Fp& b) const { return mul_fn(*this, b); }
  RIPP_HD Fp sqr() const { return *this * *this; }
  RIPP_HD Fp& operator+=(const Fp& b) { return *this = *this + b; }
  RIPP_HD Fp& operator-=(const Fp& b) { return *this = *this - b; }
  RIPP_HD Fp& operator*=(const Fp& b) { return *this = *this * b; }

  // canonical integer <-> Montgomery
  RIPP_HD Fp to_mont() const {
    Fp r2;
#pragma unroll
    for (int i = 0; i < N; i++) r2.v[i] = P::r2(i);
    return *this * r2;
  }
  RIPP_HD Fp from_mont() const {
    Fp o = zero();
    o.v[0] = 1;
    return *this * o;
  }
  // a^-1 (0 for a = 0) by safegcd divsteps (modinv.cuh): ~25 k simple instructions instead of the ~570 products in
  // series of Fermat's a^(p-2).  The Montgomery value a R is inverted as an integer, (a R)^-1 = a^-1 R^-1, and
  // brought back to Montgomery form with one product by R^3.
  RIPP_FN Fp inv() const {
    Fp x, r2;
    modinv::inverse_words<P>(x.v, v);
#pragma unroll
    for (int i = 0; i < N; i++) r2.v[i] = P::r2(i);
    return x * (r2 * r2);
  }
  // a^(p-2), kept as the independent cross-check of inv() (tests) and for A/B timing
  RIPP_FN Fp inv_fermat() const {
    Fp r = one();
    for (int i = P::BITS - 1; i >= 0; i--) {
      r = r.sqr();
      if ((P::pm2(i >> 5) >> (i & 31)) & 1) r = r * *this;
    }
    return r;
  }
};

using Fq = Fp<FqParams>;
using Fr = Fp<FrParams>;

}  // namespace ripp
