// Mock of the Rust shim (tools/ripp-b200) in C++: a host program that holds its values the way arkworks 0.4 holds
// them -- Fp = BigInt<N>([u64; N]) in Montgomery form, Projective { x, y, z } (Jacobian), Affine { x, y, infinity } --
// in structs that are NOT the ABI's packed layout, copies them FIELD BY FIELD into packed limb arrays exactly as
// tools/ripp-b200/src/pack.rs does, and calls the C ABI of include/ripp_b200.h.  It links libripp_b200.so and nothing
// else (no torch, no Python): what a maintainer's `cargo test` of the shim would exercise.
//
// in.bin (written by tests/test_gpu_abi_host.py from oracle values):
//   u64 n | n x G1 Jacobian (18 u64) | n x G2 Jacobian (36) | n x Fr (4) | n x Fr (4)          -- L1 inner products
//   u64 m | alpha (4) | beta (4) | m x proof.a (G1 affine x, y: 12) | proof.b (24) | proof.c (12), infinity flag bytes
// out.bin: GT (72 u64) | G1 Jacobian (18) | G2 Jacobian (36) | Fr (4) | i64 status of the length-mismatch call |
//          u64 proof_len | proof bytes | u8 accept
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ripp_b200.h"

struct ArkFq { uint64_t limbs[6]; };
struct ArkFr { uint64_t limbs[4]; };
struct ArkFq2 { ArkFq c0, c1; };
struct ArkG1Projective { ArkFq x, y, z; };
struct ArkG2Projective { ArkFq2 x, y, z; };
struct ArkG1Affine { ArkFq x, y; bool infinity; };   // trailing padding: sizeof != 96, so no memcpy of arrays
struct ArkG2Affine { ArkFq2 x, y; bool infinity; };
struct ArkProof { ArkG1Affine a; ArkG2Affine b; ArkG1Affine c; };  // ark_groth16::Proof<Bls12_381>

static void put(std::vector<uint64_t>& o, const ArkFq& f) { o.insert(o.end(), f.limbs, f.limbs + 6); }
static void put(std::vector<uint64_t>& o, const ArkFq2& f) { put(o, f.c0); put(o, f.c1); }
static void put(std::vector<uint64_t>& o, const ArkFr& f) { o.insert(o.end(), f.limbs, f.limbs + 4); }
static void put_aff(std::vector<uint64_t>& o, const ArkG1Affine& p) {
  if (p.infinity) { o.insert(o.end(), 12, 0); return; }
  put(o, p.x); put(o, p.y);
}
static void put_aff(std::vector<uint64_t>& o, const ArkG2Affine& p) {
  if (p.infinity) { o.insert(o.end(), 24, 0); return; }
  put(o, p.x); put(o, p.y);
}

#define CHECK(call)                                                                      \
  do {                                                                                   \
    int s_ = (call);                                                                     \
    if (s_ != RIPP_OK) {                                                                 \
      fprintf(stderr, "%s -> %d: %s\n", #call, s_, ripp_last_error_string());            \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)

template <class T>
static bool rd(FILE* f, T* dst, size_t count) { return fread(dst, sizeof(T), count, f) == count; }

int main(int argc, char** argv) {
  if (argc != 3) return 1;
  FILE* in = fopen(argv[1], "rb");
  FILE* out = fopen(argv[2], "wb");
  if (!in || !out) return 1;
  static_assert(sizeof(ArkG1Affine) != 96, "the affine struct is deliberately not the packed layout");
  ripp_ctx* ctx = nullptr;
  CHECK(ripp_ctx_create(0, &ctx));

  // ---- L1: the three InnerProduct impls ------------------------------------------------------------------
  uint64_t n = 0;
  if (!rd(in, &n, 1)) return 1;
  std::vector<ArkG1Projective> g1(n);
  std::vector<ArkG2Projective> g2(n);
  std::vector<ArkFr> s(n), t(n);
  for (auto& p : g1) if (!rd(in, &p.x, 1) || !rd(in, &p.y, 1) || !rd(in, &p.z, 1)) return 1;
  for (auto& p : g2) if (!rd(in, &p.x, 1) || !rd(in, &p.y, 1) || !rd(in, &p.z, 1)) return 1;
  for (auto& x : s) if (!rd(in, &x, 1)) return 1;
  for (auto& x : t) if (!rd(in, &x, 1)) return 1;
  std::vector<uint64_t> a, b, sw, tw;
  for (auto& p : g1) { put(a, p.x); put(a, p.y); put(a, p.z); }
  for (auto& p : g2) { put(b, p.x); put(b, p.y); put(b, p.z); }
  for (auto& x : s) put(sw, x);
  for (auto& x : t) put(tw, x);
  uint64_t gt[72], r1[18], r2[36], rs[4];
  CHECK(ripp_pairing_ip(ctx, a.data(), n, b.data(), n, gt));         // GpuPairingInnerProduct::inner_product
  CHECK(ripp_msm_g1(ctx, a.data(), n, sw.data(), n, r1));            // GpuMultiexponentiationInnerProductG1
  CHECK(ripp_msm_g2(ctx, b.data(), n, sw.data(), n, r2));            // ...G2
  CHECK(ripp_scalar_ip(ctx, sw.data(), n, tw.data(), n, rs));        // GpuScalarInnerProduct
  int64_t mismatch = ripp_pairing_ip(ctx, a.data(), n, b.data(), n - 1, gt);  // Err(MessageLengthInvalid(n, n - 1))
  CHECK(ripp_pairing_ip(ctx, a.data(), n, b.data(), n, gt));
  fwrite(gt, 8, 72, out);
  fwrite(r1, 8, 18, out);
  fwrite(r2, 8, 36, out);
  fwrite(rs, 8, 4, out);
  fwrite(&mismatch, 8, 1, out);

  // ---- setup + aggregate_proofs + verify_aggregate_proof --------------------------------------------------
  uint64_t m = 0;
  ArkFr alpha, beta;
  if (!rd(in, &m, 1) || !rd(in, &alpha, 1) || !rd(in, &beta, 1)) return 1;
  std::vector<ArkProof> proofs(m);
  for (auto& p : proofs) {
    uint8_t inf[3];
    if (!rd(in, &p.a.x, 1) || !rd(in, &p.a.y, 1) || !rd(in, &p.b.x, 1) || !rd(in, &p.b.y, 1) || !rd(in, &p.c.x, 1) ||
        !rd(in, &p.c.y, 1) || !rd(in, inf, 3))
      return 1;
    p.a.infinity = inf[0];
    p.b.infinity = inf[1];
    p.c.infinity = inf[2];
  }
  uint64_t vk_words = 0, mi = 0;
  if (!rd(in, &mi, 1) || !rd(in, &vk_words, 1)) return 1;
  std::vector<uint64_t> vk(vk_words), inputs(m * mi * 4);
  if (!rd(in, vk.data(), vk_words) || !rd(in, inputs.data(), inputs.size())) return 1;
  void *srs1 = nullptr, *srs2 = nullptr;
  CHECK(ripp_dev_alloc(ctx, (2 * m - 1) * 96, &srs1));
  CHECK(ripp_dev_alloc(ctx, (2 * m - 1) * 192, &srs2));
  uint64_t g_beta[12], h_alpha[24];
  CHECK(ripp_tipa_setup_dev(ctx, alpha.limbs, beta.limbs, m, srs1, srs2, g_beta, h_alpha));  // GpuSrs::setup
  std::vector<uint64_t> pa, pb, pc;
  for (auto& p : proofs) { put_aff(pa, p.a); put_aff(pb, p.b); put_aff(pc, p.c); }
  std::vector<uint8_t> proof(1 << 20);
  size_t plen = 0;
  CHECK(ripp_tipp_aggregate(ctx, srs1, srs2, pa.data(), pb.data(), pc.data(), m, proof.data(), proof.size(), &plen));
  // VerifierSRS { g, h, g_beta, h_alpha }: the generators are power 0 of each tower
  std::vector<uint64_t> vsrs(12 + 24 + 12 + 24);
  CHECK(ripp_dev_download(ctx, vsrs.data(), srs1, 96));
  CHECK(ripp_dev_download(ctx, vsrs.data() + 12, srs2, 192));
  memcpy(vsrs.data() + 36, g_beta, 96);
  memcpy(vsrs.data() + 48, h_alpha, 192);
  int accept = 0;
  CHECK(ripp_tipp_verify_aggregate(ctx, vsrs.data(), vk.data(), mi, inputs.data(), m, proof.data(), plen, &accept));
  uint64_t pl = plen;
  fwrite(&pl, 8, 1, out);
  fwrite(proof.data(), 1, plen, out);
  uint8_t acc = (uint8_t)accept;
  fwrite(&acc, 1, 1, out);
  ripp_dev_free(ctx, srs1);
  ripp_dev_free(ctx, srs2);
  ripp_ctx_destroy(ctx);
  fclose(out);
  return 0;
}
