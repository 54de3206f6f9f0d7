// Host-simulation build of the device headers (TEST INFRASTRUCTURE ONLY).
// The CUDA product never links this file: it exists so that the exact limb sequences, tower, curve
// and pairing code that the kernels inline can be unit-tested against the Python oracle on a
// CPU-only box (the carry flag is emulated, see ripp_b200/csrc/limb.cuh).  Built by tests/conftest.py.
#define RIPP_HOSTSIM 1
#include "../../ripp_b200/csrc/l6.cuh"
#include "../../ripp_b200/csrc/x3.cuh"
#include "../../ripp_b200/csrc/xt.cuh"
#include "../../ripp_b200/csrc/endo.cuh"
#include <pthread.h>
#include <thread>
#include <vector>
#include <string.h>
using namespace ripp;

template <class T>
static T ld(const uint32_t* p) { T x; memcpy(&x, p, sizeof(T)); return x; }
template <class T>
static void st(uint32_t* p, const T& x) { memcpy(p, &x, sizeof(T)); }

#define BIN(name, T, expr) void name(const uint32_t* a_, const uint32_t* b_, uint32_t* r_) { T a = ld<T>(a_), b = ld<T>(b_); (void)b; st<T>(r_, expr); }

extern "C" {
BIN(hs_fq_mul, Fq, a * b)
BIN(hs_fq_add, Fq, a + b)
BIN(hs_fq_sub, Fq, a - b)
BIN(hs_fq_inv, Fq, a.inv())
BIN(hs_fq_half, Fq, a.half())
BIN(hs_fr_mul, Fr, a * b)
BIN(hs_fr_add, Fr, a + b)
BIN(hs_fr_sub, Fr, a - b)
BIN(hs_fr_inv, Fr, a.inv())
BIN(hs_fr_half, Fr, a.half())
BIN(hs_fq2_mul, Fq2, a * b)
BIN(hs_fq2_sqr, Fq2, a.sqr())
BIN(hs_fq2_inv, Fq2, a.inv())
BIN(hs_fq6_mul, Fq6, a * b)
BIN(hs_fq6_sqr, Fq6, a.sqr())
BIN(hs_fq6_inv, Fq6, a.inv())
BIN(hs_fq12_mul, Fq12, a * b)
BIN(hs_fq12_sqr, Fq12, a.sqr())
BIN(hs_fq12_inv, Fq12, a.inv())
BIN(hs_fq12_cyc_sqr, Fq12, a.cyclotomic_sqr())
BIN(hs_fq12_frob1, Fq12, a.frob<1>())
BIN(hs_fq12_frob2, Fq12, a.frob<2>())
BIN(hs_fq12_frob3, Fq12, a.frob<3>())
BIN(hs_fq12_exp_by_x, Fq12, exp_by_x(a))
BIN(hs_final_exp, Fq12, final_exponentiation(a))
void hs_fq12_mul_by_014(const uint32_t* f, const uint32_t* d0, const uint32_t* d1, const uint32_t* d4, uint32_t* r) {
  st<Fq12>(r, ld<Fq12>(f).mul_by_014(ld<Fq2>(d0), ld<Fq2>(d1), ld<Fq2>(d4)));
}
// group ops on packed affine inputs (identity = (0,0)); outputs affine
void hs_g1_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, G1Jac::from_affine(ld<G1Aff>(a)).add(G1Jac::from_affine(ld<G1Aff>(b))).to_affine()); }
void hs_g1_add_mixed(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, G1Jac::from_affine(ld<G1Aff>(a)).add_mixed(ld<G1Aff>(b)).to_affine()); }
void hs_g1_dbl(const uint32_t* a, const uint32_t*, uint32_t* r) { st(r, G1Jac::from_affine(ld<G1Aff>(a)).dbl().to_affine()); }
void hs_g1_mul(const uint32_t* a, const uint32_t* bits, int nbits, uint32_t* r) { st(r, scalar_mul(ld<G1Aff>(a), bits, nbits).to_affine()); }
void hs_g2_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, G2Jac::from_affine(ld<G2Aff>(a)).add(G2Jac::from_affine(ld<G2Aff>(b))).to_affine()); }
void hs_g2_add_mixed(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, G2Jac::from_affine(ld<G2Aff>(a)).add_mixed(ld<G2Aff>(b)).to_affine()); }
void hs_g2_dbl(const uint32_t* a, const uint32_t*, uint32_t* r) { st(r, G2Jac::from_affine(ld<G2Aff>(a)).dbl().to_affine()); }
void hs_g2_mul(const uint32_t* a, const uint32_t* bits, int nbits, uint32_t* r) { st(r, scalar_mul(ld<G2Aff>(a), bits, nbits).to_affine()); }
// lazy arithmetic: r = REDC(sum_i a_i * b_i) for k products
void hs_fq_dot_redc(const uint32_t* a, const uint32_t* b, int k, int subs, uint32_t* r) {
  uint32_t acc[24] = {0}, t[24];
  for (int i = 0; i < k; i++) {
    detail::wide_mul<FqParams>(t, a + 12 * i, b + 12 * i);
    detail::wide_add<24>(acc, t);
  }
  detail::redc_wide<FqParams>(r, acc, subs);
}
// l6.cuh's lazy sums: T (13 limbs, < 2^386) -> T mod p
void hs_lz_reduce13(const uint32_t* T, uint32_t* r) { st(r, l6::lz_reduce13(T)); }
// one lane's share of a pass: slots = 8 canonical Fq values (12 words each): P = slots 0.., M = slots 4..;
// shape 0: <4, 3, 1, HALF> (pass B), 1: <3, 3, 1> (pass F), 2: <2, 1, 12> (pass E)
void hs_lz_pass(int shape, int half, const uint32_t* slots, uint32_t* r) {
  const int P[4] = {0, 12, 24, 36}, M[3] = {48, 60, 72};
  Fq o;
  if (shape == 0) o = l6::lz_pass<4, 3, 1, true>(slots, P, M, half != 0);
  else if (shape == 1) o = l6::lz_pass<3, 3, 1, false>(slots, P, M, false);
  else o = l6::lz_pass<2, 1, 12, false>(slots, P, M, false);
  st(r, o);
}
uint64_t hs_mul_count(int which) { return detail::mul_count_[which]; }
void hs_mul_count_reset() { detail::mul_count_[0] = detail::mul_count_[1] = 0; }
// k * P through the GLV / GLS decomposition (endo.cuh); k = 8 canonical words
void hs_g1_endo_mul(const uint32_t* p, const uint32_t* k, uint32_t* r) { EndoBits b; endo_decompose<1>(k, b); st(r, endo_mul<Fq>(ld<G1Aff>(p), b).to_affine()); }
void hs_g2_endo_mul(const uint32_t* p, const uint32_t* k, uint32_t* r) { EndoBits b; endo_decompose<2>(k, b); st(r, endo_mul<Fq2>(ld<G2Aff>(p), b).to_affine()); }
void hs_g1_endo(const uint32_t* a, const uint32_t*, uint32_t* r) { st(r, endo_map(ld<G1Aff>(a))); }
void hs_g2_endo(const uint32_t* a, const uint32_t*, uint32_t* r) { st(r, endo_map(ld<G2Aff>(a))); }
void hs_g1_gen(uint32_t* r) { st(r, g1_generator()); }
void hs_g2_gen(uint32_t* r) { st(r, g2_generator()); }
void hs_miller(const uint32_t* p, const uint32_t* q, uint32_t* r) { st(r, miller_loop(ld<G1Aff>(p), ld<G2Aff>(q))); }
void hs_pairing_product(int n, const uint32_t* ps, const uint32_t* qs, uint32_t* r) {
  Fq12 f = Fq12::one();
  for (int i = 0; i < n; i++) f = f * miller_loop(ld<G1Aff>(ps + 24 * i), ld<G2Aff>(qs + 48 * i));
  st(r, final_exponentiation(f));
}
}

// ---- L6 (six-lane Fq12) code paths: each lane is a host thread, sync() a pthread barrier ----------
namespace ripp { namespace l6 {
void host_barrier(void* bar) { pthread_barrier_wait((pthread_barrier_t*)bar); }
}}
static inline int tower_slot(int k) { return (k & 1) * 3 + (k >> 1); }

// W lanes per coefficient: 6 (W = 1) or 18 (W = 3) host threads per group
template <int W, class Fn>
static void run_group(Fn fn, int nreg = 8) {
  std::vector<uint32_t> sm(l6::group_words(nreg, 4) + l6::BUS_TOTAL, 0);
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, 6 * W);
  std::vector<std::thread> th;
  for (int t = 0; t < 6 * W; t++)
    th.emplace_back([&, t] {
      l6::CtxT<W> c{t / W, sm.data(), &bar, t % W, sm.data() + l6::group_words(nreg, 4), 0,
                    sm.data() + l6::group_words(nreg, 4) + l6::BUS_ZERO};
      fn(c);
    });
  for (auto& t : th) t.join();
  pthread_barrier_destroy(&bar);
}
template <class C>
static void l6_load_reg(const C& c, int reg, const uint32_t* fq12_tower) {
  l6::st2(l6::freg(c, reg) + c.k * l6::FQ2W, ld<Fq2>(fq12_tower + 24 * tower_slot(c.k)));
  l6::sync(c);
}
template <class C>
static void l6_store_reg(const C& c, int reg, uint32_t* fq12_tower) {
  if (c.role == 0) st<Fq2>(fq12_tower + 24 * tower_slot(c.k), l6::ld2(l6::freg(c, reg) + c.k * l6::FQ2W));
  l6::sync(c);
}
// op: 0 mul, 1 sqr, 2 conj, 3 frob1, 4 frob2, 5 inv, 6 exp_by_x, 7 final_exp, 8 cyc_sqr, 9 pow_fr
template <int W>
static void l6_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
  run_group<W>([&](const l6::CtxT<W>& c) {
    l6_load_reg(c, 0, a);
    l6_load_reg(c, 1, b);
    switch (op) {
      case 0: l6::mul(c, 2, 0, 1); break;
      case 1: l6::sqr(c, 2, 0); break;
      case 2: l6::conj(c, 2, 0); break;
      case 3: l6::frob(c, 2, 0, 1); break;
      case 4: l6::frob(c, 2, 0, 2); break;
      case 5: l6::inv(c, 2, 0, 3, 4, 5); break;
      case 6: l6::exp_by_x(c, 2, 0); break;
      case 7: l6::final_exp(c); l6::copy(c, 2, 0); break;
      case 8: l6::cyc_sqr(c, 2, 0); break;
      case 9: {  // pow_fr: exponent = first 8 words of b (canonical)
        if (c.k == 0) for (int i = 0; i < 8; i++) c.sm[l6::OFF_LINE + i] = b[i];
        l6::sync(c);
        l6::copy(c, 1, 0);
        l6::pow_fr(c, 2, 1, 3, 4, 5, c.sm + l6::OFF_LINE);
        break;
      }
    }
    l6_store_reg(c, 2, r);
  });
}
template <int W>
static void l6_mul_line(const uint32_t* f, const uint32_t* d0, const uint32_t* d1, const uint32_t* d4, uint32_t* r) {
  run_group<W>([&](const l6::CtxT<W>& c) {
    l6_load_reg(c, 0, f);
    if (c.k == 0) {
      l6::st2(c.sm + l6::OFF_LINE, ld<Fq2>(d0));
      l6::st2(c.sm + l6::OFF_LINE + 24, ld<Fq2>(d1));
      l6::st2(c.sm + l6::OFF_LINE + 48, ld<Fq2>(d4));
    }
    l6::sync(c);
    l6::mul_line(c, 0, 0);
    l6_store_reg(c, 0, r);
  });
}
// Miller loop (+ optional final exponentiation) over npairs pairs sharing one accumulator; valid[j] = 0 masks a pair
template <int W>
static void l6_miller(const uint32_t* p, const uint32_t* q, const int* valid, int npairs, int with_final_exp, uint32_t* r) {
  run_group<W>([&](const l6::CtxT<W>& c) {
    uint32_t* pairs = c.sm + l6::OFF_F + 8 * l6::F12W;
    if (c.k == 0) {
      for (int j = 0; j < npairs; j++) {
        uint32_t* pb = pairs + j * l6::PAIR_WORDS;
        G1Aff P = ld<G1Aff>(p + 24 * j);
        l6::stage_pair_p(pb, P);
        l6::set_pair_q(pb, q + 48 * j);  // Q is read through its address (same words as a packed G2Aff)
        pb[l6::PB_VALID] = valid[j];
      }
    }
    l6::sync(c);
    l6::miller(c, pairs, npairs);
    if (with_final_exp) l6::final_exp(c);
    l6_store_reg(c, 0, r);
  });
}
extern "C" {
void hs_l6_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) { l6_op<1>(op, a, b, r); }
void hs_l18_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) { l6_op<3>(op, a, b, r); }
void hs_l6_mul_line(const uint32_t* f, const uint32_t* d0, const uint32_t* d1, const uint32_t* d4, uint32_t* r) { l6_mul_line<1>(f, d0, d1, d4, r); }
void hs_l18_mul_line(const uint32_t* f, const uint32_t* d0, const uint32_t* d1, const uint32_t* d4, uint32_t* r) { l6_mul_line<3>(f, d0, d1, d4, r); }
void hs_l6_miller(const uint32_t* p, const uint32_t* q, const int* valid, int npairs, int with_final_exp, uint32_t* r) { l6_miller<1>(p, q, valid, npairs, with_final_exp, r); }
void hs_l18_miller(const uint32_t* p, const uint32_t* q, const int* valid, int npairs, int with_final_exp, uint32_t* r) { l6_miller<3>(p, q, valid, npairs, with_final_exp, r); }
}

// ---- x3 (three lanes per point): each lane a host thread, gather3 through a shared bus -------------
namespace {
struct X3Bus {
  pthread_barrier_t bar;
  Fq slot[3];
};
thread_local int x3_r = 0;
thread_local X3Bus* x3_bus = nullptr;
}
namespace ripp { namespace x3 {
int lane_r() { return x3_r; }
void gather3(const Fq& mine, Fq& t0, Fq& t1, Fq& t2) {
  x3_bus->slot[x3_r] = mine;
  pthread_barrier_wait(&x3_bus->bar);
  t0 = x3_bus->slot[0];
  t1 = x3_bus->slot[1];
  t2 = x3_bus->slot[2];
  pthread_barrier_wait(&x3_bus->bar);
}
}}
template <class Fn>
static void run_x3(Fn fn) {
  X3Bus bus;
  pthread_barrier_init(&bus.bar, nullptr, 3);
  std::vector<std::thread> th;
  for (int r = 0; r < 3; r++) th.emplace_back([&, r] { x3_r = r; x3_bus = &bus; fn(r); });
  for (auto& t : th) t.join();
  pthread_barrier_destroy(&bus.bar);
}
extern "C" {
// out = k * p + lo with k given as NAF bitmaps (9 words each); group 1 = G1, 2 = G2
// out = k * p + lo exactly as k_fold_w3 computes it: endomorphism digits of the canonical scalar k (8 words),
// joint double-and-add with the team formulas, barrier-safe mixed additions, branch-free normalisation
void hs_x3_endo_fold(int group, const uint32_t* p, const uint32_t* lo, const uint32_t* k, uint32_t* out) {
  EndoBits c;
  if (group == 1) endo_decompose<1>(k, c); else endo_decompose<2>(k, c);
  run_x3([&](int r) {
    if (group == 1) {
      Jac<Fq> acc = x3::endo_mul<Fq>(ld<G1Aff>(p), c);
      acc = x3::Ops<Fq>::madd(acc, ld<G1Aff>(lo));
      G1Aff o = x3::to_affine(acc);
      if (r == 0) st(out, o);
    } else {
      typedef x3::Fq2x3 F;
      Jac<F> acc = x3::endo_mul<F>(ld<Aff<F>>(p), c);
      acc = x3::Ops<F>::madd(acc, ld<Aff<F>>(lo));
      Aff<F> o = x3::to_affine(acc);
      if (r == 0) st(out, o);
    }
  });
}
void hs_x3_fold(int group, const uint32_t* p, const uint32_t* lo, const uint32_t* pos, const uint32_t* neg, int nd, uint32_t* out) {
  run_x3([&](int r) {
    if (group == 1) {
      Jac<Fq> acc = x3::mul_naf<Fq>(ld<G1Aff>(p), pos, neg, nd);
      acc = x3::g1_madd(acc, ld<G1Aff>(lo));
      G1Aff o = x3::to_affine(acc);
      if (r == 0) st(out, o);
    } else {
      typedef x3::Fq2x3 F;
      Jac<F> acc = x3::mul_naf<F>(ld<Aff<F>>(p), pos, neg, nd);
      acc = x3::g2_madd(acc, ld<Aff<F>>(lo));
      Aff<F> o = x3::to_affine(acc);
      if (r == 0) st(out, o);
    }
  });
}
}

// ---- xt (lane teams: 3 lanes per G1 point, 9 per G2 point): each lane a host thread ----------------
namespace ripp { namespace xt {
void host_barrier(void* bar) { pthread_barrier_wait((pthread_barrier_t*)bar); }
}}
template <class F, class Fn>
static void run_xt(Fn fn) {
  constexpr int L = xt::TeamOf<F>::LANES;
  std::vector<uint32_t> bus(2 * xt::TeamOf<F>::BUS_WORDS + sizeof(Aff<F>), 0);
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, L);
  std::vector<std::thread> th;
  for (int t = 0; t < L; t++) th.emplace_back([&, t] { xt::Team tm{t, bus.data(), 0, &bar}; fn(tm); });
  for (auto& t : th) t.join();
  pthread_barrier_destroy(&bar);
}
// the same fold as k_fold4_xp computes it: one team per endomorphism part (here one after the other), parts brought to
// affine form, then lo + the parts on one team
template <class F, int M>
static void xp_fold(const uint32_t* p, const uint32_t* lo, const EndoBits& c, uint32_t* out) {
  std::vector<Aff<F>> parts(M);
  for (int part = 0; part < M; part++)
    run_xt<F>([&](const xt::Team& tm) {
      Jac<F> acc = xt::part_mul<F>(tm, ld<Aff<F>>(p), c, M, part, tm.bus + 2 * xt::TeamOf<F>::BUS_WORDS);
      Aff<F> pa = xt::to_affine<F>(tm, acc);
      if (tm.t == 0) parts[part] = pa;
    });
  run_xt<F>([&](const xt::Team& tm) {
    Jac<F> s = Jac<F>::from_affine(ld<Aff<F>>(lo));
    for (int t = 0; t < M; t++) s = xt::madd<F>(tm, s, parts[t]);
    Aff<F> o = xt::to_affine<F>(tm, s);
    if (tm.t == 0) st(out, o);
  });
}
extern "C" {
// out = k * p + lo exactly as k_fold_xt computes it (k = 8 canonical words; group 1 = G1, 2 = G2)
void hs_xt_endo_fold(int group, const uint32_t* p, const uint32_t* lo, const uint32_t* k, uint32_t* out) {
  EndoBits c;
  if (group == 1) endo_decompose<1>(k, c); else endo_decompose<2>(k, c);
  if (group == 1) {
    run_xt<Fq>([&](const xt::Team& tm) {
      Jac<Fq> acc = xt::endo_mul<Fq>(tm, ld<G1Aff>(p), c, tm.bus + 2 * xt::TeamOf<Fq>::BUS_WORDS);
      acc = xt::madd<Fq>(tm, acc, ld<G1Aff>(lo));
      G1Aff o = xt::to_affine<Fq>(tm, acc);
      if (tm.t == 0) st(out, o);
    });
  } else {
    run_xt<Fq2>([&](const xt::Team& tm) {
      Jac<Fq2> acc = xt::endo_mul<Fq2>(tm, ld<G2Aff>(p), c, tm.bus + 2 * xt::TeamOf<Fq2>::BUS_WORDS);
      acc = xt::madd<Fq2>(tm, acc, ld<G2Aff>(lo));
      G2Aff o = xt::to_affine<Fq2>(tm, acc);
      if (tm.t == 0) st(out, o);
    });
  }
}
void hs_xp_endo_fold(int group, const uint32_t* p, const uint32_t* lo, const uint32_t* k, uint32_t* out) {
  EndoBits c;
  if (group == 1) {
    endo_decompose<1>(k, c);
    xp_fold<Fq, 2>(p, lo, c, out);
  } else {
    endo_decompose<2>(k, c);
    xp_fold<Fq2, 4>(p, lo, c, out);
  }
}
}
