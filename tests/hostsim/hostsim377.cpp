// The BLS12-377 device headers (ripp_b200/csrc/bls377.cuh on fp.cuh / curve.cuh) compiled for the HOST with the emulated
// carry flag of limb.cuh: the exact limb sequences, tower, D-twist Miller loop, final exponentiation and group law the
// kernels of sipp377.cu inline, unit-tested against oracle/bls12_377.py without a GPU.  Test-only; never linked into the product.
#define RIPP_HOSTSIM 1
#include <stdint.h>
#include <string.h>

#include "../../ripp_b200/csrc/fp.cuh"
#include "../../ripp_b200/csrc/tower.cuh"
#include "../../ripp_b200/csrc/bls377.cuh"

using namespace ripp;

extern "C" {
// r (144 words, tower order) = final_exponentiation(miller_loop(P, Q)) if fe else the Miller value
void hs377_pairing(const uint32_t* p, const uint32_t* q, int fe, uint32_t* r) {
  b377::G1Aff P;
  b377::G2Aff Q;
  memcpy(&P, p, 96);
  memcpy(&Q, q, 192);
  b377::F12 f = b377::miller_loop(P, Q);
  if (fe) f = b377::final_exponentiation(f);
  b377::Fq2* out = reinterpret_cast<b377::Fq2*>(r);
  for (int k = 0; k < 6; k++) out[b377::tower_slot377(k)] = f.c[k];
}
void hs377_g1_mul(const uint32_t* p, const uint32_t* s_canon, uint32_t* r) {
  b377::G1Aff P;
  memcpy(&P, p, 96);
  b377::G1Aff o = b377::scalar_mul_words<b377::Fq>(P, s_canon).to_affine();
  memcpy(r, &o, 96);
}
void hs377_g2_mul(const uint32_t* p, const uint32_t* s_canon, uint32_t* r) {
  b377::G2Aff P;
  memcpy(&P, p, 192);
  b377::G2Aff o = b377::scalar_mul_words<b377::Fq2>(P, s_canon).to_affine();
  memcpy(r, &o, 192);
}
// r = a^2 by the cyclotomic (Granger-Scott) squaring and by the generic product, both in tower order (2 x 144 words)
void hs377_cyc_sqr(const uint32_t* a, uint32_t* r) {
  b377::F12 f;
  const b377::Fq2* in = reinterpret_cast<const b377::Fq2*>(a);
  for (int k = 0; k < 6; k++) f.c[k] = in[b377::tower_slot377(k)];
  b377::F12 c = b377::cyc_sqr(f), g = f.sqr();
  b377::Fq2* out = reinterpret_cast<b377::Fq2*>(r);
  for (int k = 0; k < 6; k++) {
    out[b377::tower_slot377(k)] = c.c[k];
    out[6 + b377::tower_slot377(k)] = g.c[k];
  }
}
void hs377_fr_inv(const uint32_t* a_mont, uint32_t* r) {
  b377::Fr x;
  memcpy(x.v, a_mont, 32);
  b377::Fr y = x.inv();
  memcpy(r, y.v, 32);
}
}
