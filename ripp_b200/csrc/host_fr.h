// Host-side Fr arithmetic on 64-bit limbs (unsigned __int128 products) for the per-proof scalar work the provers keep
// on the CPU: the 2n - 1 coefficients of the KZG quotient polynomials (tipa/mod.rs:304-337, 393-422).  The
// carry-flag emulation of limb.cuh (kept bit-identical to the device sequences for tests/hostsim) costs ~270 ns per
// product on the host; this is ~10x faster and produces the same Montgomery representation (R = 2^256), so values
// move between the two freely.  Little-endian hosts only (Fr's eight 32-bit limbs are read as four 64-bit limbs).
#pragma once
#include <stdint.h>
#include <string.h>

#include "fp.cuh"

namespace ripp {
namespace hostfr {

typedef unsigned __int128 u128;

struct Mod {
  uint64_t p[4];
  uint64_t inv;  // -p^-1 mod 2^64
};
inline const Mod& mod() {
  static const Mod m = [] {
    Mod r;
    for (int i = 0; i < 4; i++) r.p[i] = (uint64_t)FrParams::p(2 * i) | ((uint64_t)FrParams::p(2 * i + 1) << 32);
    uint64_t x = 1;  // Newton: x <- x (2 - p0 x) doubles the number of correct low bits
    for (int i = 0; i < 6; i++) x *= 2 - r.p[0] * x;
    r.inv = 0 - x;
    return r;
  }();
  return m;
}
inline void load(uint64_t* d, const Fr& a) { memcpy(d, a.v, 32); }
inline Fr store(const uint64_t* s) {
  Fr r;
  memcpy(r.v, s, 32);
  return r;
}
inline bool geq_p(const uint64_t* a, const uint64_t* p) {
  for (int i = 3; i >= 0; i--)
    if (a[i] != p[i]) return a[i] > p[i];
  return true;
}
inline void sub_p(uint64_t* a, const uint64_t* p) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] - p[i] - (uint64_t)b;
    a[i] = (uint64_t)t;
    b = (t >> 64) & 1;
  }
}
// Montgomery product (CIOS), operands and result in [0, p)
inline Fr mul(const Fr& x, const Fr& y) {
  const Mod& M = mod();
  uint64_t a[4], b[4], t[6] = {0, 0, 0, 0, 0, 0};
  load(a, x);
  load(b, y);
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a[j] * b[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * M.inv;
    c = ((u128)m * M.p[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * M.p[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  if (t[4] || geq_p(t, M.p)) sub_p(t, M.p);
  return store(t);
}
inline Fr add(const Fr& x, const Fr& y) {
  const Mod& M = mod();
  uint64_t a[4], b[4];
  load(a, x);
  load(b, y);
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a[i] + b[i];
    a[i] = (uint64_t)c;
    c >>= 64;
  }
  if ((uint64_t)c || geq_p(a, M.p)) sub_p(a, M.p);
  return store(a);
}

}  // namespace hostfr
}  // namespace ripp
