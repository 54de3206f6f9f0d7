// Modular inversion by Bernstein-Yang "safegcd" divsteps (half-delta variant) in batches of 30, on signed 30-bit
// limbs: the published algorithm of "Fast constant-time gcd computation and modular inversion" (TCHES 2019) in
// the batched form popularised by libsecp256k1's modinv32 (restated here from the description, for 381 / 255-bit
// moduli).  Replaces ark-ff's `Field::inverse` (third-party) wherever the kernels normalise a
// point (Jacobian -> affine: folds, scalings, MSM tails) or invert in the tower (final exponentiation).
//
// Why: Fermat's a^(p-2) is 381 squarings + ~190 products IN SERIES (~0.38 ms on a lone warp: after the three-warp
// teams it was 40 % of a G1 fold).  A batch of 30 divsteps is ~500 simple ALU instructions on one 32-bit word pair plus
// two 13-limb matrix applications (~100 IMAD.WIDE): 31 batches = ~25 k instructions, about 10x less latency.
// The instruction stream is data independent (masks, no branches), so the lanes of a warp never diverge.
// 0 maps to 0 (as Fp::inv did).  Values stay in Montgomery form: inv(a R) = a^-1 R^-1, times R^3 (Montgomery) = a^-1 R.
#pragma once
#include "limb.cuh"

namespace ripp {
namespace modinv {

template <class P>
struct Cfg {
  static constexpr int N = P::N;                       // 32-bit words
  static constexpr int L = (32 * N + 29) / 30;         // signed 30-bit limbs (13 for Fq, 9 for Fr)
  // hddivsteps needed for a d-bit modulus: floor((45907 d + 26313) / 19929) (879 for 381 bits, 589 for 255 bits);
  // two spare batches on top (extra batches are harmless: once g = 0 a batch is the identity on f and d mod p)
  static constexpr int BATCHES = ((45907 * P::BITS + 26313) / 19929 + 29) / 30 + 2;
};

static constexpr int32_t M30 = (int32_t)(0xffffffffu >> 2);

// 30 divsteps on the low words of (f, g): transition matrix [[u, v], [q, r]] (entries in [-2^30, 2^30]) with
// 2^30 [f', g'] = M [f, g].  zeta = -(delta + 1/2).
RIPP_HD int32_t divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, int32_t& tu, int32_t& tv, int32_t& tq, int32_t& tr) {
  uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
#pragma unroll 2
  for (int i = 0; i < 30; i++) {
    uint32_t mask1 = (uint32_t)(zeta >> 31);          // zeta < 0
    uint32_t mask2 = 0u - (g & 1u);                   // g odd
    uint32_t x = (f ^ mask1) - mask1, y = (u ^ mask1) - mask1, z = (v ^ mask1) - mask1;  // conditionally negated f, u, v
    g += x & mask2;
    q += y & mask2;
    r += z & mask2;
    mask1 &= mask2;
    zeta = (int32_t)((uint32_t)zeta ^ mask1) - 1;     // -zeta - 2 or zeta - 1
    f += g & mask1;
    u += q & mask1;
    v += r & mask1;
    g >>= 1;
    u <<= 1;
    v <<= 1;
  }
  tu = (int32_t)u;
  tv = (int32_t)v;
  tq = (int32_t)q;
  tr = (int32_t)r;
  return zeta;
}

// [f, g] <- M [f, g] / 2^30 (exact)
template <int L>
RIPP_HD void update_fg(int32_t* f, int32_t* g, int32_t u, int32_t v, int32_t q, int32_t r) {
  int64_t cf = (int64_t)u * f[0] + (int64_t)v * g[0];
  int64_t cg = (int64_t)q * f[0] + (int64_t)r * g[0];
  cf >>= 30;
  cg >>= 30;
#pragma unroll
  for (int i = 1; i < L; i++) {
    cf += (int64_t)u * f[i] + (int64_t)v * g[i];
    cg += (int64_t)q * f[i] + (int64_t)r * g[i];
    f[i - 1] = (int32_t)cf & M30;
    cf >>= 30;
    g[i - 1] = (int32_t)cg & M30;
    cg >>= 30;
  }
  f[L - 1] = (int32_t)cf;
  g[L - 1] = (int32_t)cg;
}

// [d, e] <- M [d, e] / 2^30 mod p, with d, e kept in (-2p, p): multiples of p are added so the division is exact
template <int L>
RIPP_HD void update_de(int32_t* d, int32_t* e, int32_t u, int32_t v, int32_t q, int32_t r, const int32_t* m, uint32_t m_inv30) {
  int32_t sd = d[L - 1] >> 31, se = e[L - 1] >> 31;
  int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
  int64_t cd = (int64_t)u * d[0] + (int64_t)v * e[0];
  int64_t ce = (int64_t)q * d[0] + (int64_t)r * e[0];
  md -= (int32_t)((m_inv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
  me -= (int32_t)((m_inv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
  cd += (int64_t)m[0] * md;
  ce += (int64_t)m[0] * me;
  cd >>= 30;
  ce >>= 30;
#pragma unroll
  for (int i = 1; i < L; i++) {
    cd += (int64_t)u * d[i] + (int64_t)v * e[i] + (int64_t)m[i] * md;
    ce += (int64_t)q * d[i] + (int64_t)r * e[i] + (int64_t)m[i] * me;
    d[i - 1] = (int32_t)cd & M30;
    cd >>= 30;
    e[i - 1] = (int32_t)ce & M30;
    ce >>= 30;
  }
  d[L - 1] = (int32_t)cd;
  e[L - 1] = (int32_t)ce;
}

// r in (-2p, p), negated if sign < 0, brought to [0, p)
template <int L>
RIPP_HD void normalize(int32_t* r, int32_t sign, const int32_t* m) {
  int32_t cond_add = r[L - 1] >> 31;
#pragma unroll
  for (int i = 0; i < L; i++) r[i] += m[i] & cond_add;
  int32_t cond_neg = sign >> 31;
#pragma unroll
  for (int i = 0; i < L; i++) r[i] = (r[i] ^ cond_neg) - cond_neg;
#pragma unroll
  for (int i = 0; i < L - 1; i++) {
    r[i + 1] += r[i] >> 30;
    r[i] &= M30;
  }
  cond_add = r[L - 1] >> 31;
#pragma unroll
  for (int i = 0; i < L; i++) r[i] += m[i] & cond_add;
#pragma unroll
  for (int i = 0; i < L - 1; i++) {
    r[i + 1] += r[i] >> 30;
    r[i] &= M30;
  }
}

// 32-bit words (little endian, value < 2^(32 N)) -> L limbs of 30 bits
template <int N, int L>
RIPP_HD void to_limbs30(int32_t* o, const uint32_t* w) {
#pragma unroll
  for (int j = 0; j < L; j++) {
    int bit = 30 * j, k = bit >> 5, sh = bit & 31;
    uint64_t v = k < N ? w[k] : 0u;
    if (k + 1 < N) v |= (uint64_t)w[k + 1] << 32;
    o[j] = (int32_t)((uint32_t)(v >> sh) & (uint32_t)M30);
  }
}
template <int N, int L>
RIPP_HD void from_limbs30(uint32_t* w, const int32_t* l) {
#pragma unroll
  for (int k = 0; k < N; k++) w[k] = 0;
#pragma unroll
  for (int j = 0; j < L; j++) {
    int bit = 30 * j, k = bit >> 5, sh = bit & 31;
    uint64_t v = (uint64_t)(uint32_t)l[j] << sh;
    if (k < N) w[k] |= (uint32_t)v;
    if (k + 1 < N) w[k + 1] |= (uint32_t)(v >> 32);
  }
}

// x^-1 mod p as plain integers (x < p, canonical words in / out); 0 -> 0
template <class P>
RIPP_HD void inverse_words(uint32_t* out, const uint32_t* x) {
  constexpr int N = Cfg<P>::N, L = Cfg<P>::L;
  int32_t m[L], f[L], g[L], d[L], e[L];
  uint32_t pw[N];
#pragma unroll
  for (int i = 0; i < N; i++) pw[i] = P::p(i);
  to_limbs30<N, L>(m, pw);
  to_limbs30<N, L>(g, x);
#pragma unroll
  for (int i = 0; i < L; i++) {
    f[i] = m[i];
    d[i] = 0;
    e[i] = i == 0 ? 1 : 0;
  }
  const uint32_t m_inv30 = (0u - P::M0) & (uint32_t)M30;  // p^-1 mod 2^30 (M0 = -p^-1 mod 2^32)
  int32_t zeta = -1;
#pragma unroll 1
  for (int it = 0; it < Cfg<P>::BATCHES; it++) {
    int32_t u, v, q, r;
    zeta = divsteps_30(zeta, (uint32_t)f[0], (uint32_t)g[0], u, v, q, r);
    update_de<L>(d, e, u, v, q, r, m, m_inv30);
    update_fg<L>(f, g, u, v, q, r);
  }
  normalize<L>(d, f[L - 1], m);
  from_limbs30<N, L>(out, d);
}

}  // namespace modinv
}  // namespace ripp
