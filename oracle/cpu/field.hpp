// ORACLE (test / baseline infrastructure, NOT product code): CPU restatement of the arithmetic the
// reference gets from ark-ff / ark-ec / ark-bls12-381 0.4 (third-party, not in /root/reference).
// PARITY UNPINNED against arkworks itself (cannot be built here); validated against oracle/*.py.
//
// 64-bit limbs, Montgomery form with R = 2^(64 N) (ark-ff `MontBackend`),
// CIOS multiplication on unsigned __int128.  All constants are derived at start-up from the modulus.
#pragma once
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

template <int N>
struct BigInt {
  u64 l[N];
};

template <int N>
static inline bool geq(const u64* a, const u64* b) {
  for (int i = N - 1; i >= 0; i--) {
    if (a[i] > b[i]) return true;
    if (a[i] < b[i]) return false;
  }
  return true;
}
template <int N>
static inline u64 add_n(u64* r, const u64* a, const u64* b) {
  u64 c = 0;
  for (int i = 0; i < N; i++) {
    u128 t = (u128)a[i] + b[i] + c;
    r[i] = (u64)t;
    c = (u64)(t >> 64);
  }
  return c;
}
template <int N>
static inline u64 sub_n(u64* r, const u64* a, const u64* b) {
  u64 br = 0;
  for (int i = 0; i < N; i++) {
    u128 t = (u128)a[i] - b[i] - br;
    r[i] = (u64)t;
    br = (u64)(t >> 64) & 1;
  }
  return br;
}

// Tag selects the modulus (0 = Fq, 1 = Fr); parameters are filled by init_fields().
template <int N, int Tag>
struct Fp {
  u64 l[N];
  static u64 P[N], ONE[N], R2[N], INV;

  static Fp zero() {
    Fp r;
    memset(r.l, 0, sizeof(r.l));
    return r;
  }
  static Fp one() {
    Fp r;
    memcpy(r.l, ONE, sizeof(r.l));
    return r;
  }
  bool is_zero() const {
    u64 o = 0;
    for (int i = 0; i < N; i++) o |= l[i];
    return o == 0;
  }
  bool operator==(const Fp& b) const { return memcmp(l, b.l, sizeof(l)) == 0; }
  bool operator!=(const Fp& b) const { return !(*this == b); }
  Fp operator+(const Fp& b) const {
    Fp r;
    add_n<N>(r.l, l, b.l);
    if (geq<N>(r.l, P)) sub_n<N>(r.l, r.l, P);
    return r;
  }
  Fp operator-(const Fp& b) const {
    Fp r;
    if (sub_n<N>(r.l, l, b.l)) add_n<N>(r.l, r.l, P);
    return r;
  }
  Fp operator-() const { return zero() - *this; }
  Fp dbl() const { return *this + *this; }
  // CIOS Montgomery product (ark-ff `mul_assign` without the no-carry shortcut)
  Fp operator*(const Fp& b) const {
    u64 t[N + 2];
    memset(t, 0, sizeof(t));
    for (int i = 0; i < N; i++) {
      u64 c = 0;
      for (int j = 0; j < N; j++) {
        u128 s = (u128)l[j] * b.l[i] + t[j] + c;
        t[j] = (u64)s;
        c = (u64)(s >> 64);
      }
      u128 s = (u128)t[N] + c;
      t[N] = (u64)s;
      t[N + 1] = (u64)(s >> 64);
      u64 m = t[0] * INV;
      s = (u128)m * P[0] + t[0];
      c = (u64)(s >> 64);
      for (int j = 1; j < N; j++) {
        s = (u128)m * P[j] + t[j] + c;
        t[j - 1] = (u64)s;
        c = (u64)(s >> 64);
      }
      s = (u128)t[N] + c;
      t[N - 1] = (u64)s;
      t[N] = t[N + 1] + (u64)(s >> 64);
    }
    Fp r;
    memcpy(r.l, t, sizeof(r.l));
    if (t[N] || geq<N>(r.l, P)) sub_n<N>(r.l, r.l, P);
    return r;
  }
  Fp sqr() const { return *this * *this; }
  Fp pow(const u64* e, int words) const {
    Fp r = one();
    for (int i = words * 64 - 1; i >= 0; i--) {
      r = r.sqr();
      if ((e[i / 64] >> (i % 64)) & 1) r = r * *this;
    }
    return r;
  }
  Fp inv() const {  // Fermat
    u64 e[N];
    u64 two[N] = {2};
    sub_n<N>(e, P, two);
    return pow(e, N);
  }
  Fp from_mont() const {
    Fp o = zero();
    o.l[0] = 1;
    return *this * o;
  }
  Fp to_mont() const {
    Fp r2;
    memcpy(r2.l, R2, sizeof(r2.l));
    return *this * r2;
  }
  static Fp from_u64(u64 v) {
    Fp r = zero();
    r.l[0] = v;
    return r.to_mont();
  }

  static void init(const u64* modulus) {
    memcpy(P, modulus, sizeof(P));
    u64 inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - P[0] * inv;  // Newton: p^-1 mod 2^64
    INV = (u64)0 - inv;
    // R mod p and R^2 mod p by repeated doubling of 1
    u64 x[N];
    memset(x, 0, sizeof(x));
    x[0] = 1;
    for (int i = 0; i < 2 * 64 * N; i++) {
      u64 c = add_n<N>(x, x, x);
      if (c || geq<N>(x, P)) sub_n<N>(x, x, P);
      if (i == 64 * N - 1) memcpy(ONE, x, sizeof(ONE));
    }
    memcpy(R2, x, sizeof(R2));
  }
};
template <int N, int Tag>
u64 Fp<N, Tag>::P[N];
template <int N, int Tag>
u64 Fp<N, Tag>::ONE[N];
template <int N, int Tag>
u64 Fp<N, Tag>::R2[N];
template <int N, int Tag>
u64 Fp<N, Tag>::INV;

typedef Fp<6, 0> Fq;
typedef Fp<4, 1> Fr;
