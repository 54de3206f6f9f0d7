// Host-simulation build of the device headers (TEST INFRASTRUCTURE ONLY).
// The CUDA product never links this file: it exists so that the exact limb sequences, tower, curve
// and pairing code that the kernels inline can be unit-tested against the Python oracle on a
// CPU-only box (the carry flag is emulated, see ripp_b200/csrc/limb.cuh).  Built by tests/conftest.py.
#define RIPP_HOSTSIM 1
#include "../../ripp_b200/csrc/pairing.cuh"
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
uint64_t hs_mul_count(int which) { return detail::mul_count_[which]; }
void hs_mul_count_reset() { detail::mul_count_[0] = detail::mul_count_[1] = 0; }
void hs_g1_gen(uint32_t* r) { st(r, g1_generator()); }
void hs_g2_gen(uint32_t* r) { st(r, g2_generator()); }
void hs_miller(const uint32_t* p, const uint32_t* q, uint32_t* r) { st(r, miller_loop(ld<G1Aff>(p), ld<G2Aff>(q))); }
void hs_pairing_product(int n, const uint32_t* ps, const uint32_t* qs, uint32_t* r) {
  Fq12 f = Fq12::one();
  for (int i = 0; i < n; i++) f = f * miller_loop(ld<G1Aff>(ps + 24 * i), ld<G2Aff>(qs + 48 * i));
  st(r, final_exponentiation(f));
}
}
