// SIPP over BLS12-377 + Blake2s: the reference's own instantiation (sipp/src/lib.rs:228-254 `SIPP<Bls12_377, Blake2s>`,
// sipp/examples/scaling-ipp.rs:10) -- SURVEY.md §8f-4.  One thread per pairing / per element on the BLS12-377 parameter
// set of bls377.cuh; host side: the Fiat-Shamir RNG (sipp/src/rng.rs:12-73: Blake2s seed, ChaCha20 stream) over ark-ec's
// DEFAULT short-Weierstrass serialisation (little-endian coordinates, flags in the top bits of the last byte; SURVEY.md
// App. A-4), which is what ark-bls12-377 uses (ark-bls12-381 overrides it with the Zcash format: gipa.cu).
//
// Entry points (host pointers, Montgomery limbs, affine points packed x | y with identity = all zero, GT in tower order):
//   ripp377_pairing_ip_affine          prod_i e(a_i, b_i)                         inner_products/src/lib.rs:52-116 on E = Bls12_377
//   ripp377_sipp_product_with_coeffs   prod_i e(r_i a_i, b_i)                     sipp/src/lib.rs:184-217
//   ripp377_sipp_prove / _verify                                                  sipp/src/lib.rs:42-180
#include "common.cuh"
#include "hash.h"
#include "bls377.cuh"

namespace {
using ripp::b377::F12;
typedef ripp::b377::Fq Fq7;
typedef ripp::b377::Fr Fr7;
typedef ripp::b377::Fq2 Fq27;
typedef ripp::b377::G1Aff G1A7;
typedef ripp::b377::G2Aff G2A7;
typedef ripp::b377::G1Jac G1J7;
typedef ripp::b377::G2Jac G2J7;
typedef std::vector<uint8_t> Bytes;

// ---- kernels --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k377_miller(const G1A7* __restrict__ p, const G2A7* __restrict__ q, size_t n, F12* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = ripp::b377::miller_loop(p[i], q[i]);
}
// out[t] = prod in[t R .. min(m, t R + R))
__global__ void __launch_bounds__(32) k377_f12_reduce(const F12* __restrict__ in, size_t m, int R, F12* __restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t lo = t * R;
  if (lo >= m) return;
  size_t hi = lo + R < m ? lo + R : m;
  F12 f = in[lo];
  for (size_t j = lo + 1; j < hi; j++) f = f * in[j];
  out[t] = f;
}
// tower-order output: slot 3 i + j <- flat coefficient 2 j + i
__global__ void k377_final_exp(const F12* __restrict__ in, Fq27* __restrict__ out_tower, int with_final_exp) {
  if (threadIdx.x || blockIdx.x) return;
  F12 f = with_final_exp ? ripp::b377::final_exponentiation(in[0]) : in[0];
  for (int k = 0; k < 6; k++) out_tower[ripp::b377::tower_slot377(k)] = f.c[k];
}
template <class F>
__global__ void __launch_bounds__(32) k377_scale(const ripp::Aff<F>* __restrict__ pts, const Fr7* __restrict__ sc, size_t n,
                                                ripp::Aff<F>* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr7 s = sc[i].from_mont();
  out[i] = ripp::b377::scalar_mul_words<F>(pts[i], s.v).to_affine();
}
// out[i] = hi[i] * c + lo[i]
template <class F>
__global__ void __launch_bounds__(32) k377_fold(const ripp::Aff<F>* __restrict__ hi, const ripp::Aff<F>* __restrict__ lo, Fr7 c_canon,
                                               size_t n, ripp::Aff<F>* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = ripp::b377::scalar_mul_words<F>(hi[i], c_canon.v).add_mixed(lo[i]).to_affine();
}
// out[t] = sum in[t R ..) (affine in, affine out)
template <class F>
__global__ void __launch_bounds__(32) k377_sum(const ripp::Aff<F>* __restrict__ in, size_t m, int R, ripp::Aff<F>* __restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t lo = t * R;
  if (lo >= m) return;
  size_t hi = lo + R < m ? lo + R : m;
  ripp::Jac<F> acc = ripp::Jac<F>::inf();
  for (size_t j = lo; j < hi; j++) acc = acc.add_mixed(in[j]);
  out[t] = acc.to_affine();
}
// out[i] = in[i]^(e[i]) (tower order in and out)
__global__ void __launch_bounds__(32) k377_gt_pow(const Fq27* __restrict__ in_tower, const Fr7* __restrict__ e, size_t n, F12* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F12 f;
  for (int k = 0; k < 6; k++) f.c[k] = in_tower[6 * i + ripp::b377::tower_slot377(k)];
  Fr7 s = e[i].from_mont();
  out[i] = ripp::b377::pow_words(f, s.v);
}
__global__ void k377_tower_to_flat(const Fq27* __restrict__ in_tower, F12* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  for (int k = 0; k < 6; k++) out->c[k] = in_tower[ripp::b377::tower_slot377(k)];
}

// ---- device helpers ----------------------------------------------------------------------------------------------
// prod_i e(p_i, q_i) over device vectors -> 576 B (tower order) at out_dev
int pairing_product(ripp_ctx* ctx, const G1A7* p, const G2A7* q, size_t n, void* out_dev, bool with_final_exp = true) {
  void *bufA, *bufB;
  size_t m = n ? n : 1;
  OK(scratch(ctx, 2, m * sizeof(F12) + 4096, &bufA));
  OK(scratch(ctx, 3, (m / 8 + 2) * sizeof(F12) + 4096, &bufB));
  F12 *src = (F12*)bufA, *dst = (F12*)bufB;
  if (n == 0) {
    F12 one = F12::one();
    CU(cudaMemcpyAsync(src, &one, sizeof(one), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  } else {
    k377_miller<<<(unsigned)((n + 31) / 32), 32, 0, ctx->stream>>>(p, q, n, src);
    LAUNCHED(ctx);
  }
  while (m > 1) {
    size_t mo = (m + 7) / 8;
    k377_f12_reduce<<<(unsigned)((mo + 31) / 32), 32, 0, ctx->stream>>>(src, m, 8, dst);
    LAUNCHED(ctx);
    F12* t = src;
    src = dst;
    dst = t;
    m = mo;
  }
  k377_final_exp<<<1, 1, 0, ctx->stream>>>(src, (Fq27*)out_dev, with_final_exp ? 1 : 0);
  LAUNCHED(ctx);
  return RIPP_OK;
}
// ---- host serialisation: ark-serialize 0.4 defaults ------------------------------------------
void put_fq(Bytes& o, const Fq7& m) {
  Fq7 c = m.from_mont();
  const uint8_t* p = (const uint8_t*)c.v;
  o.insert(o.end(), p, p + 48);
}
void put_fr(Bytes& o, const Fr7& m) {
  Fr7 c = m.from_mont();
  const uint8_t* p = (const uint8_t*)c.v;
  o.insert(o.end(), p, p + 32);
}
bool canon_gt(const Fq7& a, const Fq7& b) {  // canonical integers a > b
  Fq7 x = a.from_mont(), y = b.from_mont();
  for (int i = 11; i >= 0; i--)
    if (x.v[i] != y.v[i]) return x.v[i] > y.v[i];
  return false;
}
void put_g1(Bytes& o, const G1A7& p) {
  if (p.is_inf()) {
    o.insert(o.end(), 95, 0);
    o.push_back(0x40);
    return;
  }
  put_fq(o, p.x);
  put_fq(o, p.y);
  if (canon_gt(p.y, -p.y)) o.back() |= 0x80;  // SWFlags::YIsNegative: y > -y
}
void put_g2(Bytes& o, const G2A7& p) {
  if (p.is_inf()) {
    o.insert(o.end(), 191, 0);
    o.push_back(0x40);
    return;
  }
  put_fq(o, p.x.c0);
  put_fq(o, p.x.c1);
  put_fq(o, p.y.c0);
  put_fq(o, p.y.c1);
  Fq27 ny = -p.y;  // Fq2 ordering in ark-ff: lexicographic on (c1, c0)
  bool neg = (p.y.c1 != ny.c1) ? canon_gt(p.y.c1, ny.c1) : canon_gt(p.y.c0, ny.c0);
  if (neg) o.back() |= 0x80;
}
void put_gt(Bytes& o, const void* tower576) {
  const Fq7* c = (const Fq7*)tower576;
  for (int i = 0; i < 12; i++) put_fq(o, c[i]);
}
void put_u64_le(Bytes& o, uint64_t v) {
  for (int i = 0; i < 8; i++) o.push_back((uint8_t)(v >> (8 * i)));
}

struct Rng377 {  // sipp/src/rng.rs:12-73 with D = Blake2s
  uint8_t seed[32];
  void init(const Bytes& m) { ripp_hash::blake2s256(m.data(), m.size(), seed); }
  void absorb(const Bytes& fresh) {
    Bytes b(fresh);
    b.insert(b.end(), seed, seed + 32);
    ripp_hash::blake2s256(b.data(), b.size(), seed);
  }
  Fr7 next_u128() const {  // u128::rand = the first 16 keystream bytes, little-endian (lib.rs:85)
    uint8_t ks[64];
    ripp_hash::chacha20_block(seed, 0, ks);
    Fr7 c = Fr7::zero();
    memcpy(c.v, ks, 16);
    return c.to_mont();
  }
};
Rng377 seed_rng(const G1A7* a, const G2A7* b, const Fr7* r, size_t n, const void* value_gt) {
  Bytes s;  // lib.rs:56-60: (a, b, r, value) serialised uncompressed
  put_u64_le(s, n);
  for (size_t i = 0; i < n; i++) put_g1(s, a[i]);
  put_u64_le(s, n);
  for (size_t i = 0; i < n; i++) put_g2(s, b[i]);
  put_u64_le(s, n);
  for (size_t i = 0; i < n; i++) put_fr(s, r[i]);
  put_gt(s, value_gt);
  Rng377 g;
  g.init(s);
  return g;
}
}  // namespace

extern "C" int ripp377_pairing_ip_affine(ripp_ctx* ctx, const void* g1_aff, size_t n_left, const void* g2_aff, size_t n_right, void* gt_out) {
  if (!ctx || !gt_out) return fail(RIPP_ERR_ARG, "null argument");
  if (n_left != n_right)
    return fail(RIPP_ERR_LEN_MISMATCH, "left length, right length: " + std::to_string(n_left) + ", " + std::to_string(n_right));
  size_t n = n_left;
  if (n && (!g1_aff || !g2_aff)) return fail(RIPP_ERR_ARG, "null argument");
  CU(cudaSetDevice(ctx->device));
  void* d;
  OK(scratch(ctx, 14, n * 288 + 2048, &d));
  char* A = (char*)d;
  char* B = A + n * 96;
  char* out = B + n * 192;
  if (n) {
    CU(cudaMemcpyAsync(A, g1_aff, n * 96, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(B, g2_aff, n * 192, cudaMemcpyHostToDevice, ctx->stream));
  }
  OK(pairing_product(ctx, (const G1A7*)A, (const G2A7*)B, n, out));
  CU(cudaMemcpyAsync(gt_out, out, 576, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return RIPP_OK;
}

extern "C" int ripp377_sipp_product_with_coeffs(ripp_ctx* ctx, const void* a_aff, const void* b_aff, const void* r, size_t n, void* gt_out) {
  if (!ctx || !gt_out || (n && (!a_aff || !b_aff || !r))) return fail(RIPP_ERR_ARG, "null argument");
  CU(cudaSetDevice(ctx->device));
  void* d;
  OK(scratch(ctx, 14, n * (96 + 192 + 32 + 96) + 2048, &d));
  char* A = (char*)d;
  char* B = A + n * 96;
  char* Rp = B + n * 192;
  char* AR = Rp + n * 32;
  char* out = AR + n * 96;
  CU(cudaMemcpyAsync(A, a_aff, n * 96, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(B, b_aff, n * 192, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(Rp, r, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (n) {
    k377_scale<Fq7><<<(unsigned)((n + 31) / 32), 32, 0, ctx->stream>>>((const G1A7*)A, (const Fr7*)Rp, n, (G1A7*)AR);
    LAUNCHED(ctx);
  }
  OK(pairing_product(ctx, (const G1A7*)AR, (const G2A7*)B, n, out));
  CU(cudaMemcpyAsync(gt_out, out, 576, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return RIPP_OK;
}

extern "C" int ripp377_sipp_prove(ripp_ctx* ctx, const void* a_aff, const void* b_aff, const void* r, size_t n, const void* value_gt,
                                  uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
  if (!ctx || !a_aff || !b_aff || !r || !value_gt) return fail(RIPP_ERR_ARG, "null argument");
  if (n == 0 || (n & (n - 1))) return fail(RIPP_ERR_NOT_POW2, "SIPP needs a power-of-two length");
  CU(cudaSetDevice(ctx->device));
  Rng377 rng = seed_rng((const G1A7*)a_aff, (const G2A7*)b_aff, (const Fr7*)r, n, value_gt);
  void* d;
  OK(scratch(ctx, 14, n * (96 + 192 + 32 + 96) + 4096, &d));
  char* A0 = (char*)d;
  char* Bv = A0 + n * 96;
  char* Rp = Bv + n * 192;
  char* Av = Rp + n * 32;
  char* res = Av + n * 96;
  CU(cudaMemcpyAsync(A0, a_aff, n * 96, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(Bv, b_aff, n * 192, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(Rp, r, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  k377_scale<Fq7><<<(unsigned)((n + 31) / 32), 32, 0, ctx->stream>>>((const G1A7*)A0, (const Fr7*)Rp, n, (G1A7*)Av);  // lib.rs:61-66
  LAUNCHED(ctx);
  Bytes proof;
  size_t len = n;
  while (len != 1) {
    len /= 2;
    // lib.rs:77-78: z_l = prod e(a_R, b_L), z_r = prod e(a_L, b_R)
    // the two products run concurrently (second one on a child context: own stream and scratch)
    ripp_ctx* kid = ripp_child(ctx, 0);
    if (!kid) return fail(RIPP_ERR_CUDA, "child context");
    OK(ripp_fork(ctx, kid));
    OK(pairing_product(kid, (const G1A7*)Av, (const G2A7*)(Bv + len * 192), len, res + 576));
    OK(pairing_product(ctx, (const G1A7*)(Av + len * 96), (const G2A7*)Bv, len, res));
    OK(ripp_join(ctx, kid));
    uint8_t z[2 * 576];
    CU(cudaMemcpyAsync(z, res, 2 * 576, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    Bytes buf;
    put_gt(buf, z);
    put_gt(buf, z + 576);
    proof.insert(proof.end(), buf.begin(), buf.end());
    rng.absorb(buf);          // lib.rs:80-84
    Fr7 x = rng.next_u128();  // lib.rs:85
    Fr7 x_inv = x.inv();
    // lib.rs:87-100: a <- a_R x + a_L ; b <- b_R x^-1 + b_L
    k377_fold<Fq7><<<(unsigned)((len + 31) / 32), 32, 0, ctx->stream>>>((const G1A7*)(Av + len * 96), (const G1A7*)Av, x.from_mont(), len,
                                                                        (G1A7*)Av);
    LAUNCHED(ctx);
    k377_fold<Fq27><<<(unsigned)((len + 31) / 32), 32, 0, ctx->stream>>>((const G2A7*)(Bv + len * 192), (const G2A7*)Bv, x_inv.from_mont(),
                                                                         len, (G2A7*)Bv);
    LAUNCHED(ctx);
  }
  CU(cudaStreamSynchronize(ctx->stream));
  if (proof_len) *proof_len = proof.size();
  if (!proof_out || proof_cap < proof.size()) return fail(RIPP_ERR_ARG, "output buffer too small: need " + std::to_string(proof.size()));
  memcpy(proof_out, proof.data(), proof.size());
  return RIPP_OK;
}

// SIPP::verify (lib.rs:109-180): z' = value * prod_j z_l[j]^(x_j) z_r[j]^(x_j^-1); a' = MSM(a, s o r), b' = MSM(b, s^-1);
// accept iff e(a', b') == z'.  Proof elements are decoded as canonical little-endian Fq coefficients.
extern "C" int ripp377_sipp_verify(ripp_ctx* ctx, const void* a_aff, const void* b_aff, const void* r, size_t n, const void* value_gt,
                                   const uint8_t* proof, size_t proof_len, int* accept) {
  if (!ctx || !a_aff || !b_aff || !r || !value_gt || !proof || !accept) return fail(RIPP_ERR_ARG, "null argument");
  if (n < 2 || (n & (n - 1))) return fail(RIPP_ERR_NOT_POW2, "SIPP needs a power-of-two length of at least 2");
  size_t k = 0;
  while (((size_t)1 << k) < n) k++;
  if (proof_len != k * 2 * 576) return fail(RIPP_ERR_ARG, "proof has the wrong length for this instance");
  CU(cudaSetDevice(ctx->device));
  *accept = 0;
  Rng377 rng = seed_rng((const G1A7*)a_aff, (const G2A7*)b_aff, (const Fr7*)r, n, value_gt);
  std::vector<Fr7> xs(k), xinv(k);
  // decode the 2 k GT elements (canonical LE -> Montgomery) and derive the challenges
  std::vector<Fq7> zs(2 * k * 12);
  for (size_t j = 0; j < k; j++) {
    Bytes buf(proof + j * 1152, proof + (j + 1) * 1152);
    for (int c = 0; c < 24; c++) {
      Fq7 v;
      memcpy(v.v, buf.data() + 48 * c, 48);
      for (int i = 11; i >= 0; i--) {  // canonical: < p
        if (v.v[i] < ripp::b377::FqP::p(i)) break;
        if (v.v[i] > ripp::b377::FqP::p(i) || i == 0) return fail(RIPP_ERR_ARG, "non-canonical field element in the proof");
      }
      zs[j * 24 + c] = v.to_mont();
    }
    rng.absorb(buf);
    xs[j] = rng.next_u128();
    xinv[j] = xs[j].inv();
  }
  // exponent vectors (lib.rs:140-160)
  std::vector<Fr7> s(n, Fr7::one()), sinv(n, Fr7::one()), e(2 * k);
  const Fr7* rr = (const Fr7*)r;
  for (size_t j = 0; j < k; j++) {
    for (size_t i = 0; i < n; i++)
      if (i & ((size_t)1 << (k - j - 1))) {
        s[i] = s[i] * xs[j];
        sinv[i] = sinv[i] * xinv[j];
      }
    e[2 * j] = xs[j];
    e[2 * j + 1] = xinv[j];
  }
  for (size_t i = 0; i < n; i++) s[i] = s[i] * rr[i];
  void* d;
  size_t o_b = n * 96, o_s = o_b + n * 192, o_si = o_s + n * 32, o_as = o_si + n * 32, o_bs = o_as + n * 96, o_t1 = o_bs + n * 192,
         o_t2 = o_t1 + n * 96, o_z = o_t2 + n * 192, o_e = o_z + 2 * k * 576, o_p = o_e + 2 * k * 32, o_val = o_p + (2 * k + 2) * 576,
         o_res = o_val + 576;
  OK(scratch(ctx, 14, o_res + 4096, &d));
  char* D = (char*)d;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(D, a_aff, n * 96, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(D + o_b, b_aff, n * 192, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(D + o_s, s.data(), n * 32, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(D + o_si, sinv.data(), n * 32, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(D + o_z, zs.data(), 2 * k * 576, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(D + o_e, e.data(), 2 * k * 32, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(D + o_val, value_gt, 576, cudaMemcpyHostToDevice, st));
  unsigned nb = (unsigned)((n + 31) / 32);
  k377_scale<Fq7><<<nb, 32, 0, st>>>((const G1A7*)D, (const Fr7*)(D + o_s), n, (G1A7*)(D + o_as));
  LAUNCHED(ctx);
  k377_scale<Fq27><<<nb, 32, 0, st>>>((const G2A7*)(D + o_b), (const Fr7*)(D + o_si), n, (G2A7*)(D + o_bs));
  LAUNCHED(ctx);
  // tree sums
  {
    G1A7 *src = (G1A7*)(D + o_as), *dst = (G1A7*)(D + o_t1);
    for (size_t m = n; m > 1;) {
      size_t mo = (m + 7) / 8;
      k377_sum<Fq7><<<(unsigned)((mo + 31) / 32), 32, 0, st>>>(src, m, 8, dst);
      LAUNCHED(ctx);
      G1A7* t = src;
      src = dst;
      dst = t;
      m = mo;
    }
    CU(cudaMemcpyAsync(D + o_res + 1024, src, 96, cudaMemcpyDeviceToDevice, st));
    G2A7 *s2 = (G2A7*)(D + o_bs), *d2 = (G2A7*)(D + o_t2);
    for (size_t m = n; m > 1;) {
      size_t mo = (m + 7) / 8;
      k377_sum<Fq27><<<(unsigned)((mo + 31) / 32), 32, 0, st>>>(s2, m, 8, d2);
      LAUNCHED(ctx);
      G2A7* t = s2;
      s2 = d2;
      d2 = t;
      m = mo;
    }
    CU(cudaMemcpyAsync(D + o_res + 2048, s2, 192, cudaMemcpyDeviceToDevice, st));
  }
  // z' : value and the 2 k powers multiplied together (flat F12 values at o_p)
  F12* pw = (F12*)(D + o_p);
  k377_tower_to_flat<<<1, 1, 0, st>>>((const Fq27*)(D + o_val), pw);
  LAUNCHED(ctx);
  k377_gt_pow<<<(unsigned)((2 * k + 31) / 32), 32, 0, st>>>((const Fq27*)(D + o_z), (const Fr7*)(D + o_e), 2 * k, pw + 1);
  LAUNCHED(ctx);
  k377_f12_reduce<<<1, 32, 0, st>>>(pw, 2 * k + 1, (int)(2 * k + 1), pw + 2 * k + 1);
  LAUNCHED(ctx);
  k377_final_exp<<<1, 1, 0, st>>>(pw + 2 * k + 1, (Fq27*)(D + o_res), 0);
  LAUNCHED(ctx);
  // e(a', b')
  OK(pairing_product(ctx, (const G1A7*)(D + o_res + 1024), (const G2A7*)(D + o_res + 2048), 1, D + o_res + 576 + 2048));
  uint8_t lhs[576], rhs[576];
  CU(cudaMemcpyAsync(rhs, D + o_res, 576, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(lhs, D + o_res + 576 + 2048, 576, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  *accept = memcmp(lhs, rhs, 576) == 0 ? 1 : 0;
  return RIPP_OK;
}
