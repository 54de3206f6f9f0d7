// ORACLE (test / baseline infrastructure, NOT product code; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this).
//
// CPU restatement, with the reference's own parallel structure (rayon -> OpenMP), of the heavy leaf
// operations of arkworks-rs/ripp's inner-pairing-product path:
//   cfg_multi_pairing            inner_products/src/lib.rs:77-116  (normalise, G2 prepare, one
//                                Miller-loop chunk per thread, product, ONE final exponentiation)
//   G::msm (VariableBaseMSM)     inner_products/src/lib.rs:140 ; tipa/mod.rs:333-334
//   element-wise scalar muls     gipa.rs:261-291 (mul_helper, ip_proofs/src/lib.rs:15-19),
//                                groth16_aggregation.rs:118-131, sipp/src/lib.rs:61-65,87-100
// The algorithms are the published ones ark-ec 0.4 implements (Costello-Lange-Naehrig line
// functions with precomputed G2 coefficients, Hayashida-Hayasaka-Teruya final exponentiation,
// bucket-method MSM); PARITY UNPINNED against arkworks itself, validated against oracle/*.py.
//
// Array layouts are those of include/ripp_b200.h (arkworks' in-memory Montgomery limbs).
#include <math.h>
#include <omp.h>
#include <stdio.h>

#include <vector>

#include "tower.hpp"

Fq2 Fq12::FROB1[6];
Fq2 Fq12::FROB2[6];

static const u64 Q_MOD[6] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                             0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};
static const u64 R_MOD[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
static const u64 X_ABS = 0xd201000000010000ull;  // |x|, x < 0
static Fq TWO_INV;
static Fq2 B2;  // 4 (1 + u)
static bool g_init = false;

static void init_fields() {
  if (g_init) return;
  Fq::init(Q_MOD);
  Fr::init(R_MOD);
  TWO_INV = Fq::from_u64(2).inv();
  B2 = {Fq::from_u64(4), Fq::from_u64(4)};
  // (p - 1) / 6 by schoolbook division of the limbs
  u64 e[6], rem = 0;
  u64 pm1[6];
  memcpy(pm1, Q_MOD, sizeof(pm1));
  pm1[0] -= 1;
  for (int i = 5; i >= 0; i--) {
    u128 cur = ((u128)rem << 64) | pm1[i];
    e[i] = (u64)(cur / 6);
    rem = (u64)(cur % 6);
  }
  Fq2 xi = {Fq::one(), Fq::one()};
  Fq2 g = xi.pow(e, 6);
  Fq12::FROB1[0] = Fq2::one();
  for (int k = 1; k < 6; k++) Fq12::FROB1[k] = Fq12::FROB1[k - 1] * g;
  // xi^((p^2-1)/6) = g^(p+1) = g * conj(g)
  for (int k = 0; k < 6; k++) Fq12::FROB2[k] = Fq12::FROB1[k] * Fq12::FROB1[k].conj();
  g_init = true;
}

// ------------------------------------------------------------------------------------------------
// groups: Jacobian `Projective` of ark-ec short_weierstrass (a = 0)
// ------------------------------------------------------------------------------------------------
template <class F>
struct Aff {
  F x, y;
  bool is_inf() const { return x.is_zero() && y.is_zero(); }
};
template <class F>
struct Jac {
  F x, y, z;
  static Jac inf() { return {F::one(), F::one(), F::zero()}; }
  bool is_inf() const { return z.is_zero(); }
  Jac dbl() const {
    if (is_inf()) return *this;
    F a = x.sqr(), b = y.sqr(), c = b.sqr();
    F d = ((x + b).sqr() - a - c).dbl();
    F e = a + a.dbl(), f = e.sqr();
    Jac r;
    r.z = (y * z).dbl();
    r.x = f - d.dbl();
    r.y = e * (d - r.x) - c.dbl().dbl().dbl();
    return r;
  }
  Jac add_affine(const Aff<F>& q) const {
    if (q.is_inf()) return *this;
    if (is_inf()) return {q.x, q.y, F::one()};
    F z1z1 = z.sqr(), u2 = q.x * z1z1, s2 = q.y * z * z1z1;
    if (x == u2) {
      if (y == s2) return dbl();
      return inf();
    }
    F h = u2 - x, hh = h.sqr(), i = hh.dbl().dbl(), j = h * i, r = (s2 - y).dbl(), v = x * i;
    Jac o;
    o.x = r.sqr() - j - v.dbl();
    o.y = r * (v - o.x) - (y * j).dbl();
    o.z = (z + h).sqr() - z1z1 - hh;
    return o;
  }
  Jac add(const Jac& q) const {
    if (is_inf()) return q;
    if (q.is_inf()) return *this;
    F z1z1 = z.sqr(), z2z2 = q.z.sqr();
    F u1 = x * z2z2, u2 = q.x * z1z1, s1 = y * q.z * z2z2, s2 = q.y * z * z1z1;
    if (u1 == u2) {
      if (s1 == s2) return dbl();
      return inf();
    }
    F h = u2 - u1, i = h.dbl().sqr(), j = h * i, r = (s2 - s1).dbl(), v = u1 * i;
    Jac o;
    o.x = r.sqr() - j - v.dbl();
    o.y = r * (v - o.x) - (s1 * j).dbl();
    o.z = ((z + q.z).sqr() - z1z1 - z2z2) * h;
    return o;
  }
};

// batch normalisation (ark-ec `normalize_batch`: Montgomery's trick on the z coordinates)
template <class F>
static void normalize_batch(const Jac<F>* in, size_t n, Aff<F>* out) {
  std::vector<F> pref(n);
  F acc = F::one();
  for (size_t i = 0; i < n; i++) {
    pref[i] = acc;
    if (!in[i].is_inf()) acc = acc * in[i].z;
  }
  F inv = acc.inv();
  for (size_t i = n; i-- > 0;) {
    if (in[i].is_inf()) {
      out[i] = {F::zero(), F::zero()};
      continue;
    }
    F zi = inv * pref[i];
    inv = inv * in[i].z;
    F zi2 = zi.sqr();
    out[i] = {in[i].x * zi2, in[i].y * zi2 * zi};
  }
}

// mul_bigint: plain MSB-first double-and-add over the canonical scalar (ark-ec `mul_bigint`)
template <class F>
static Jac<F> mul_scalar(const Aff<F>& p, const u64* k) {
  Jac<F> acc = Jac<F>::inf();
  bool started = false;
  for (int i = 255; i >= 0; i--) {
    if (started) acc = acc.dbl();
    if ((k[i / 64] >> (i % 64)) & 1) {
      acc = acc.add_affine(p);
      started = true;
    }
  }
  return acc;
}

// ------------------------------------------------------------------------------------------------
// pairing: ark-ec models::bls12 (G2Prepared + multi_miller_loop + final_exponentiation)
// ------------------------------------------------------------------------------------------------
struct Ell {
  Fq2 c0, c1, c2;
};
struct G2Prepared {
  std::vector<Ell> coeffs;
  bool inf;
};

static Ell doubling_step(Fq2& rx, Fq2& ry, Fq2& rz) {
  Fq2 a = (rx * ry).scale(TWO_INV), b = ry.sqr(), c = rz.sqr();
  Fq2 e = B2 * (c.dbl() + c), f = e.dbl() + e;
  Fq2 g = (b + f).scale(TWO_INV), h = (ry + rz).sqr() - (b + c), i = e - b, j = rx.sqr(), e2 = e.sqr();
  rx = a * (b - f);
  ry = g.sqr() - (e2.dbl() + e2);
  rz = b * h;
  return {i, j.dbl() + j, -h};  // TwistType::M
}
static Ell addition_step(Fq2& rx, Fq2& ry, Fq2& rz, const Aff<Fq2>& q) {
  Fq2 theta = ry - q.y * rz, lambda = rx - q.x * rz;
  Fq2 c = theta.sqr(), d = lambda.sqr(), e = lambda * d, f = rz * c, g = rx * d;
  Fq2 h = e + f - g.dbl();
  rx = lambda * h;
  ry = theta * (g - h) - e * ry;
  rz = rz * e;
  Fq2 j = theta * q.x - lambda * q.y;
  return {j, -theta, lambda};
}
static G2Prepared prepare_g2(const Aff<Fq2>& q) {
  G2Prepared p;
  p.inf = q.is_inf();
  if (p.inf) return p;
  Fq2 rx = q.x, ry = q.y, rz = Fq2::one();
  p.coeffs.reserve(68);
  for (int i = 62; i >= 0; i--) {
    p.coeffs.push_back(doubling_step(rx, ry, rz));
    if ((X_ABS >> i) & 1) p.coeffs.push_back(addition_step(rx, ry, rz, q));
  }
  return p;
}
static inline void ell(Fq12& f, const Ell& c, const Aff<Fq>& p) {
  f = f.mul_by_014(c.c0, c.c1.scale(p.x), c.c2.scale(p.y));
}
// multi_miller_loop over a chunk: all pairs share one accumulator, squared once per bit
static Fq12 multi_miller_loop(const Aff<Fq>* ps, const G2Prepared* qs, size_t n) {
  std::vector<size_t> live;
  for (size_t i = 0; i < n; i++)
    if (!ps[i].is_inf() && !qs[i].inf) live.push_back(i);
  Fq12 f = Fq12::one();
  size_t idx = 0;
  for (int i = 62; i >= 0; i--) {
    f = f.sqr();
    for (size_t k : live) ell(f, qs[k].coeffs[idx], ps[k]);
    idx++;
    if ((X_ABS >> i) & 1) {
      for (size_t k : live) ell(f, qs[k].coeffs[idx], ps[k]);
      idx++;
    }
  }
  return f.conj();  // x < 0
}
static Fq12 exp_by_x(const Fq12& a) {
  Fq12 r = a;
  for (int i = 62; i >= 0; i--) {
    r = r.cyclotomic_sqr();
    if ((X_ABS >> i) & 1) r = r * a;
  }
  return r.conj();
}
static Fq12 final_exponentiation(const Fq12& f) {
  Fq12 r = f.conj() * f.inv();
  r = r.frobenius(2) * r;
  Fq12 y0 = r.cyclotomic_sqr(), y1 = exp_by_x(r), y2 = r.conj();
  y1 = y1 * y2;
  y2 = exp_by_x(y1);
  y1 = y1.conj() * y2;
  y2 = exp_by_x(y1);
  y1 = y1.frobenius(1) * y2;
  r = r * y0;
  y0 = exp_by_x(y1);
  y2 = exp_by_x(y0);
  y0 = y1.frobenius(2);
  y1 = y1.conj() * y2 * y0;
  return r * y1;
}

// ------------------------------------------------------------------------------------------------
// MSM: ark-ec VariableBaseMSM::msm_bigint (window c = ln(n) + 2, windows in parallel, running sum)
// ------------------------------------------------------------------------------------------------
template <class F>
static Jac<F> msm(const Aff<F>* bases, const u64 (*scalars)[4], size_t n) {
  if (n == 0) return Jac<F>::inf();
  int c = 3;  // ark-ec: if size < 32 { 3 } else { ln_without_floats(size) + 2 } = log2(size) * 69 / 100 + 2
  if (n >= 32) {
    int lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    c = lg * 69 / 100 + 2;
  }
  int nw = (255 + c - 1) / c;
  std::vector<Jac<F>> window_sums(nw);
#pragma omp parallel for schedule(dynamic)
  for (int w = 0; w < nw; w++) {
    int start = w * c;
    std::vector<Jac<F>> buckets(((size_t)1 << c) - 1, Jac<F>::inf());
    Jac<F> res = Jac<F>::inf();
    for (size_t i = 0; i < n; i++) {
      const u64* s = scalars[i];
      int wi = start / 64, sh = start % 64;
      u64 v = s[wi] >> sh;
      if (sh && wi + 1 < 4) v |= s[wi + 1] << (64 - sh);
      u64 d = v & (((u64)1 << c) - 1);
      if (d) buckets[d - 1] = buckets[d - 1].add_affine(bases[i]);
    }
    Jac<F> run = Jac<F>::inf();
    for (size_t b = buckets.size(); b-- > 0;) {
      run = run.add(buckets[b]);
      res = res.add(run);
    }
    window_sums[w] = res;
  }
  Jac<F> total = window_sums[nw - 1];
  for (int w = nw - 2; w >= 0; w--) {
    for (int j = 0; j < c; j++) total = total.dbl();
    total = total.add(window_sums[w]);
  }
  return total;
}

template <class F>
static Aff<F> to_affine(const Jac<F>& p) {
  Aff<F> o;
  normalize_batch<F>(&p, 1, &o);
  return o;
}

// ------------------------------------------------------------------------------------------------
// C API (arrays of Montgomery limbs; affine identity = all zero)
// ------------------------------------------------------------------------------------------------
// out[i] = sc[i] * p[i] (+ add[i] when add != NULL); stride_sc = 0 broadcasts one scalar
template <class F>
static void scale_add(const Aff<F>* p, const Fr* sc, size_t stride_sc, const Aff<F>* add, size_t n, Aff<F>* out) {
  init_fields();
  std::vector<Jac<F>> tmp(n);
#pragma omp parallel for schedule(dynamic, 16)
  for (size_t i = 0; i < n; i++) {
    Fr c = sc[i * stride_sc].from_mont();
    Jac<F> r = mul_scalar<F>(p[i], c.l);
    if (add) r = r.add_affine(add[i]);
    tmp[i] = r;
  }
  // normalize_batch per chunk, in parallel
  int nt = omp_get_max_threads();
  size_t chunk = (n + nt - 1) / nt;
  if (chunk == 0) chunk = 1;
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < (n + chunk - 1) / chunk; k++) {
    size_t lo = k * chunk, hi = lo + chunk < n ? lo + chunk : n;
    normalize_batch<F>(tmp.data() + lo, hi - lo, out + lo);
  }
}
extern "C" {

int cpu_init(void) {
  init_fields();
  return omp_get_max_threads();
}
void cpu_set_threads(int t) { omp_set_num_threads(t); }

// cfg_multi_pairing (inner_products/src/lib.rs:77-116) on affine inputs
void cpu_pairing_product(const Aff<Fq>* g1, const Aff<Fq2>* g2, size_t n, Fq12* out) {
  init_fields();
  std::vector<G2Prepared> prep(n);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) prep[i] = prepare_g2(g2[i]);  // lib.rs:86-88
  int nt = omp_get_max_threads();
  size_t chunk = (n + nt - 1) / (nt ? nt : 1);  // lib.rs:91-106: one chunk per thread
  if (chunk == 0) chunk = 1;
  size_t nchunks = (n + chunk - 1) / chunk;
  std::vector<Fq12> parts(nchunks ? nchunks : 1, Fq12::one());
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < nchunks; k++) {
    size_t lo = k * chunk, hi = lo + chunk < n ? lo + chunk : n;
    parts[k] = multi_miller_loop(g1 + lo, prep.data() + lo, hi - lo);  // lib.rs:110-112
  }
  Fq12 f = Fq12::one();
  for (size_t k = 0; k < nchunks; k++) f = f * parts[k];  // lib.rs:113
  *out = final_exponentiation(f);                          // lib.rs:115
}

void cpu_normalize_g1(const Jac<Fq>* in, size_t n, Aff<Fq>* out) {
  init_fields();
  normalize_batch<Fq>(in, n, out);
}

void cpu_msm_g1(const Aff<Fq>* bases, const Fr* sc, size_t n, Aff<Fq>* out) {
  init_fields();
  std::vector<u64> canon(4 * n);
#pragma omp parallel for
  for (size_t i = 0; i < n; i++) {
    Fr c = sc[i].from_mont();
    memcpy(&canon[4 * i], c.l, 32);
  }
  *out = to_affine(msm<Fq>(bases, (const u64(*)[4])canon.data(), n));
}
void cpu_msm_g2(const Aff<Fq2>* bases, const Fr* sc, size_t n, Aff<Fq2>* out) {
  init_fields();
  std::vector<u64> canon(4 * n);
#pragma omp parallel for
  for (size_t i = 0; i < n; i++) {
    Fr c = sc[i].from_mont();
    memcpy(&canon[4 * i], c.l, 32);
  }
  *out = to_affine(msm<Fq2>(bases, (const u64(*)[4])canon.data(), n));
}

void cpu_scale_g1(const Aff<Fq>* p, const Fr* sc, size_t n, Aff<Fq>* out) { scale_add<Fq>(p, sc, 1, nullptr, n, out); }
void cpu_scale_g2(const Aff<Fq2>* p, const Fr* sc, size_t n, Aff<Fq2>* out) { scale_add<Fq2>(p, sc, 1, nullptr, n, out); }
void cpu_fold_g1(const Aff<Fq>* hi, const Aff<Fq>* lo, const Fr* c, size_t n, Aff<Fq>* out) { scale_add<Fq>(hi, c, 0, lo, n, out); }
void cpu_fold_g2(const Aff<Fq2>* hi, const Aff<Fq2>* lo, const Fr* c, size_t n, Aff<Fq2>* out) { scale_add<Fq2>(hi, c, 0, lo, n, out); }
void cpu_fold_fr(const Fr* hi, const Fr* lo, const Fr* c, size_t n, Fr* out) {
  init_fields();
#pragma omp parallel for
  for (size_t i = 0; i < n; i++) out[i] = hi[i] * *c + lo[i];
}

// op counters are not needed here: bench.py's roofline model takes its counts from tests/hostsim
}
