// Sharded provers: ONE instance partitioned over the ranks of the context's communicator (comm.cu), one process per
// GPU.  Included by gipa.cu (shares its host-side serialisation, challenge derivation and the resident prover).
//
// SURVEY.md §8e.  Leaf inner products: contiguous slices, one partial per rank (a Miller value BEFORE the final
// exponentiation, or one affine point), ONE all-gather, combined on every rank in rank order.  GIPA rounds: cyclic
// partition -- rank k holds global indices j g + k at local index j -- so that a round's six products are products
// over LOCAL halves and its folds are local (gipa.rs:209-217 pairs index i with i + n'; g | n' keeps both on one
// rank); per round the only exchange is one all-gather of the partials of ALL instances proved together (the two
// recursions of aggregate_proofs advance in lock step: 12 partials, one collective, two hashes).  Below `tail_len`
// global elements every kernel of a round is a latency chain of a few warps: the vectors are all-gathered once and
// every rank finishes with the resident prover (gipa_prove) continuing the same transcript, so all ranks emit the
// bytes one GPU emits.  KZG openings: every rank runs the opening MSM over its contiguous slice of the SRS powers.
#pragma once

static ripp_ctx* top_ctx(ripp_ctx* c) {
  while (c->parent) c = c->parent;
  return c;
}

// in: [world][nblob] blobs of `bytes` -> out: [nblob][world]
__global__ void k_blob_transpose(const uint32_t* __restrict__ in, int world, int nblob, int words, uint32_t* __restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)world * nblob * words;
  if (t >= total) return;
  int w = (int)(t % words);
  size_t e = t / words;
  int b = (int)(e % nblob), r = (int)(e / nblob);
  out[((size_t)b * world + r) * words + w] = in[t];
}

// ---- leaf inner products over contiguous slices ---------------------------------------------------------------
extern "C" int ripp_pairing_ip_sharded_dev(ripp_ctx* ctx, const void* g1_slice_dev, const void* g2_slice_dev, size_t n_local,
                                           void* gt_out_dev) {
  if (!ctx || !gt_out_dev || (n_local && (!g1_slice_dev || !g2_slice_dev))) return fail(RIPP_ERR_ARG, "null argument");
  CU(cudaSetDevice(ctx->device));
  const int world = top_ctx(ctx)->world;
  if (world == 1) return ripp_pairing_ip_dev(ctx, g1_slice_dev, g2_slice_dev, n_local, gt_out_dev);
  void* buf;
  OK(scratch(ctx, 27, (size_t)(world + 1) * 576 + 256, &buf));
  char* part = (char*)buf;
  char* gath = part + 576;
  const void* g1[1] = {g1_slice_dev};
  const void* g2[1] = {g2_slice_dev};
  OK(ripp_pairing_batch_l6(ctx, 1, g1, g2, n_local, part, false));
  OK(ripp_all_gather_internal(ctx, part, 576, gath));
  return ripp_final_exp_l6(ctx, gath, (uint32_t)world, gt_out_dev, 1);
}

template <class F>
static int msm_sharded(ripp_ctx* ctx, const void* bases, const void* sc, size_t n_local, void* out) {
  if (!ctx || !out || (n_local && (!bases || !sc))) return fail(RIPP_ERR_ARG, "null argument");
  CU(cudaSetDevice(ctx->device));
  const int world = top_ctx(ctx)->world;
  const bool g1 = sizeof(F) == sizeof(Fq);
  if (world == 1) return g1 ? ripp_msm_g1_dev(ctx, bases, sc, n_local, out) : ripp_msm_g2_dev(ctx, bases, sc, n_local, out);
  void* buf;
  OK(scratch(ctx, 27, (size_t)(world + 1) * sizeof(Aff<F>) + 256, &buf));
  char* part = (char*)buf;
  char* gath = part + sizeof(Aff<F>);
  OK(g1 ? ripp_msm_g1_dev(ctx, bases, sc, n_local, part) : ripp_msm_g2_dev(ctx, bases, sc, n_local, part));
  OK(ripp_all_gather_internal(ctx, part, sizeof(Aff<F>), gath));
  return ripp_seg_sum_dev(ctx, g1 ? 1 : 2, gath, world, 1, out);
}
extern "C" int ripp_msm_g1_sharded_dev(ripp_ctx* ctx, const void* bases, const void* sc, size_t n_local, void* out) {
  return msm_sharded<Fq>(ctx, bases, sc, n_local, out);
}
extern "C" int ripp_msm_g2_sharded_dev(ripp_ctx* ctx, const void* bases, const void* sc, size_t n_local, void* out) {
  return msm_sharded<Fq2>(ctx, bases, sc, n_local, out);
}

// ---- GIPA rounds over the cyclic partition ----------------------------------------------------------------------
struct ShardedInst {
  GipaSpec sp;
  ripp_ctx* cx;  // the instance's own context (stream + children): the instances' kernels overlap
  char *A, *B, *V, *W;
  std::vector<std::vector<Val>> steps;
  std::vector<Fr> transcript;
};

// queue the six partial products of one round of `in` (local halves) into slots[0..6) (576 B each); no synchronisation
static int queue_round_partials(ShardedInst& in, size_t split, char* slots, char* tmp_gt) {
  const GipaSpec& sp = in.sp;
  ripp_ctx* cx = in.cx;
  Slice xs[6] = {at(sp.a, in.A, split), at(sp.w, in.W, split), at(sp.a, in.A, split), at(sp.a, in.A, 0), at(sp.w, in.W, 0), at(sp.a, in.A, 0)};
  Slice ys[6] = {at(sp.v, in.V, 0), at(sp.b, in.B, 0), at(sp.b, in.B, 0), at(sp.v, in.V, split), at(sp.b, in.B, split), at(sp.b, in.B, split)};
  if (sp.w == VT_NONE) xs[1].t = xs[4].t = VT_NONE;
  const void *g1[6], *g2[6];
  int pair_slot[6], np = 0, nk = 0;
  ripp_ctx* kids[6];
  for (int j = 0; j < 6; j++)
    if (cx->child[j]) OK(ripp_fork(cx, cx->child[j]));
  for (int i = 0; i < 6; i++) {
    int a = xs[i].t, b = ys[i].t, o = ip_out_type(a, b);
    char* dst = slots + 576 * i;
    if (o == VT_GT) {
      bool xg1 = a == VT_G1;
      g1[np] = xg1 ? xs[i].p : ys[i].p;
      g2[np] = xg1 ? ys[i].p : xs[i].p;
      pair_slot[np++] = i;
      continue;
    }
    if (a == VT_NONE || b == VT_NONE) continue;  // placeholder: the slot stays zero
    if (a == VT_FR && b == VT_FR) {
      OK(ripp_scalar_ip_dev(cx, xs[i].p, ys[i].p, split, dst));
      continue;
    }
    ripp_ctx* kid = ripp_child(cx, nk);
    if (!kid) return fail(RIPP_ERR_CUDA, "child context");
    OK(ripp_fork(cx, kid));
    kids[nk++] = kid;
    const char* pts = a == VT_FR ? ys[i].p : xs[i].p;
    const char* sc = a == VT_FR ? xs[i].p : ys[i].p;
    OK(o == VT_G1 ? ripp_msm_g1_dev(kid, pts, sc, split, dst) : ripp_msm_g2_dev(kid, pts, sc, split, dst));
  }
  if (np) {
    OK(ripp_pairing_batch_l6(cx, np, g1, g2, split, tmp_gt, false));
    for (int j = 0; j < np; j++)
      CU(cudaMemcpyAsync(slots + 576 * pair_slot[j], tmp_gt + 576 * j, 576, cudaMemcpyDeviceToDevice, cx->stream));
  }
  for (int j = 0; j < nk; j++) OK(ripp_join(cx, kids[j]));
  return RIPP_OK;
}

// Rounds of `ninst` instances of equal local length m in lock step while the global length exceeds tail_len; then the
// all-gather of what is left into global order.  On return in[i].A.. point to the gathered tail vectors (n_tail
// elements, in scratch of in[i].cx) and steps / transcript hold the rounds done.
static int sharded_rounds(ripp_ctx* ctx, ShardedInst* in, int ninst, size_t m, size_t tail_len, size_t* n_tail) {
  ripp_ctx* top = top_ctx(ctx);
  const int world = top->world;
  const size_t nslot = (size_t)6 * ninst;
  void* buf;
  OK(scratch(ctx, 28, (2 * world + 3) * nslot * 576 + 4096, &buf));
  char* parts = (char*)buf;                       // [ninst][6] x 576
  char* gath = parts + nslot * 576;               // [world][ninst * 6]
  char* tr = gath + (size_t)world * nslot * 576;  // [ninst * 6][world]
  char* comb = tr + (size_t)world * nslot * 576;  // [ninst * 6] combined values
  char* tmp_gt = comb + nslot * 576;              // pairing batch output
  uint8_t* pin;
  OK(pinned(ctx, &pin));
  while (m > 1 && m * world > (tail_len > (size_t)world ? tail_len : (size_t)world)) {
    const size_t split = m / 2;
    CU(cudaMemsetAsync(parts, 0, nslot * 576, ctx->stream));
    for (int i = 0; i < ninst; i++) {
      OK(ripp_fork(ctx, in[i].cx));
      OK(queue_round_partials(in[i], split, parts + (size_t)i * 6 * 576, tmp_gt + (size_t)i * 6 * 576));
      OK(ripp_join(ctx, in[i].cx));
    }
    OK(ripp_all_gather_internal(ctx, parts, nslot * 576, gath));
    {
      size_t total = (size_t)world * nslot * 144;
      k_blob_transpose<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t*)gath, world, (int)nslot, 144,
                                                                                 (uint32_t*)tr);
      LAUNCHED(ctx);
    }
    // combine slot by slot type: GT slots share one product + final exponentiation launch per contiguous run
    for (size_t s = 0; s < nslot;) {
      const GipaSpec& sp = in[s / 6].sp;
      int j = (int)(s % 6);
      int xt_ = (j % 3 == 1) ? sp.w : sp.a, yt_ = (j % 3 == 0) ? sp.v : sp.b;
      int o = ip_out_type(xt_, yt_);
      size_t run = 1;
      if (o == VT_GT) {  // extend over adjacent GT slots
        while (s + run < nslot) {
          const GipaSpec& sp2 = in[(s + run) / 6].sp;
          int j2 = (int)((s + run) % 6);
          if (ip_out_type((j2 % 3 == 1) ? sp2.w : sp2.a, (j2 % 3 == 0) ? sp2.v : sp2.b) != VT_GT) break;
          run++;
        }
        OK(ripp_final_exp_l6(ctx, tr + s * world * 576, (uint32_t)world, comb + s * 576, (int)run));
      } else if (xt_ == VT_NONE || yt_ == VT_NONE) {
        CU(cudaMemsetAsync(comb + s * 576, 0, 576, ctx->stream));
      } else {
        // blobs are 576 B apart whatever the element size: compact this slot's `world` elements first
        size_t es = vt_size(o);
        for (int r = 0; r < world; r++)
          CU(cudaMemcpyAsync(tmp_gt + r * es, tr + (s * world + r) * 576, es, cudaMemcpyDeviceToDevice, ctx->stream));
        OK(ripp_seg_sum_dev(ctx, o == VT_G1 ? 1 : (o == VT_G2 ? 2 : 3), tmp_gt, world, 1, comb + s * 576));
      }
      s += run;
    }
    CU(cudaMemcpyAsync(pin, comb, nslot * 576, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < ninst; i++) {
      const GipaSpec& sp = in[i].sp;
      std::vector<Val> com(6);
      for (int j = 0; j < 6; j++) {
        int xt_ = (j % 3 == 1) ? sp.w : sp.a, yt_ = (j % 3 == 0) ? sp.v : sp.b;
        com[j].t = ip_out_type(xt_, yt_);
        memset(com[j].raw, 0, sizeof(com[j].raw));
        memcpy(com[j].raw, pin + ((size_t)i * 6 + j) * 576, vt_size(com[j].t));
      }
      Fr c, c_inv;
      gipa_challenge(in[i].transcript.empty() ? Fr::zero() : in[i].transcript.back(), com.data(), &c, &c_inv);
      ripp_ctx* cx = in[i].cx;
      OK(ripp_fork(ctx, cx));
      {
        const int types[4] = {sp.a, sp.b, sp.v, sp.w};
        char* const bases[4] = {in[i].A, in[i].B, in[i].V, in[i].W};
        OK(fold_round(cx, types, bases, split, c, c_inv));
      }
      OK(ripp_join(ctx, cx));
      in[i].steps.push_back(com);
      in[i].transcript.push_back(c);
    }
    m = split;
  }
  // tail: gather the m local elements of every vector into global order (global index j g + k = local j of rank k)
  *n_tail = m * world;
  for (int i = 0; i < ninst; i++) {
    const GipaSpec& sp = in[i].sp;
    int types[4] = {sp.a, sp.b, sp.v, sp.w};
    char** vecs[4] = {&in[i].A, &in[i].B, &in[i].V, &in[i].W};
    size_t total = 0;
    for (int t = 0; t < 4; t++) total += ((m * world * vt_size(types[t]) + 255) & ~(size_t)255);
    void* tb;
    OK(scratch(in[i].cx, 29, 2 * total + 1024, &tb));
    char* dst = (char*)tb;
    char* stage = dst + total;
    for (int t = 0; t < 4; t++) {
      if (types[t] == VT_NONE || !*vecs[t]) continue;
      size_t es = vt_size(types[t]), bytes = m * es;
      if (world == 1) {
        CU(cudaMemcpyAsync(dst, *vecs[t], bytes, cudaMemcpyDeviceToDevice, ctx->stream));
      } else {
        OK(ripp_all_gather_internal(ctx, *vecs[t], bytes, stage));  // [world][m] -> [m][world]
        size_t tw = (size_t)world * m * (es / 4);
        k_blob_transpose<<<(unsigned)((tw + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t*)stage, world, (int)m, (int)(es / 4),
                                                                               (uint32_t*)dst);
        LAUNCHED(ctx);
      }
      *vecs[t] = dst;
      dst += (m * world * es + 255) & ~(size_t)255;
    }
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return RIPP_OK;
}

// GIPA::prove_with_aux for ONE instance whose four vectors are partitioned cyclically over the ranks.
extern "C" int ripp_gipa_prove_sharded_dev(ripp_ctx* ctx, int kind, const void* a_dev, const void* b_dev, const void* v_dev,
                                           const void* w_dev, size_t n_local, size_t tail_len, uint8_t* proof_out, size_t proof_cap,
                                           size_t* proof_len, void* transcript_out, uint8_t* ck_base_out, size_t ck_cap, size_t* ck_len) {
  GipaSpec sp;
  if (!ctx || !gipa_spec(kind, &sp)) return fail(RIPP_ERR_ARG, "bad context or GIPA kind");
  if (!a_dev || !b_dev || !v_dev || (sp.w != VT_NONE && !w_dev)) return fail(RIPP_ERR_ARG, "null vector");
  if (n_local == 0 || (n_local & (n_local - 1)))
    return fail(RIPP_ERR_NOT_POW2, "left length, right length: " + std::to_string(n_local) + ", " + std::to_string(n_local));
  CU(cudaSetDevice(ctx->device));
  const size_t m = n_local;
  size_t sa = m * vt_size(sp.a), sb = m * vt_size(sp.b), sv = m * vt_size(sp.v), sw = m * vt_size(sp.w);
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  ripp_ctx* cx = ripp_child(ctx, 6);
  if (!cx) return fail(RIPP_ERR_CUDA, "child context");
  void* work;
  OK(scratch(cx, 11, up(sa) + up(sb) + up(sv) + up(sw) + 1024, &work));
  ShardedInst in;
  in.sp = sp;
  in.cx = cx;
  in.A = (char*)work;
  in.B = in.A + up(sa);
  in.V = in.B + up(sb);
  in.W = sp.w == VT_NONE ? nullptr : in.V + up(sv);
  CU(cudaMemcpyAsync(in.A, a_dev, sa, cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaMemcpyAsync(in.B, b_dev, sb, cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaMemcpyAsync(in.V, v_dev, sv, cudaMemcpyDeviceToDevice, ctx->stream));
  if (in.W) CU(cudaMemcpyAsync(in.W, w_dev, sw, cudaMemcpyDeviceToDevice, ctx->stream));
  size_t n_tail = 0;
  OK(sharded_rounds(ctx, &in, 1, m, tail_len ? tail_len : ((size_t)1 << 12), &n_tail));
  GipaOut g;
  OK(gipa_prove(ctx, sp, in.A, in.B, in.V, in.W, n_tail, &g, nullptr, &in.steps, &in.transcript));
  OK(copy_out(g.proof, proof_out, proof_cap, proof_len));
  if (transcript_out) memcpy(transcript_out, g.transcript.data(), g.transcript.size() * sizeof(Fr));
  Bytes ck;
  put_val(ck, g.v0);
  if (sp.w != VT_NONE) put_val(ck, g.w0);
  return copy_out(ck, ck_base_out, ck_cap, ck_len);
}

// ---- KZG openings over contiguous slices of the SRS powers --------------------------------------------------------
// openings of the final keys of one TIPA instance (tipa/mod.rs:191-229); every rank holds the full SRS (2 n - 1
// points: 37 MB at n = 2^16) and multiplies its slice [lo, hi).  Both opening MSMs run concurrently (two streams),
// their partial points travel in ONE all-gather.
static int tipa_open_sharded(ripp_ctx* ctx, const GipaSpec& sp, const GipaOut& g, const void* srs_g1, const void* srs_g2, size_t n,
                             const Fr& r_shift, Bytes* proof) {
  ripp_ctx* top = top_ctx(ctx);
  const int world = top->world, rank = top->rank;
  const size_t n_srs = 2 * n - 1;
  std::vector<Fr> tinv(g.transcript.size());
  for (size_t i = 0; i < tinv.size(); i++) tinv[i] = g.transcript[i].inv();
  Bytes parts;
  put_fr(parts, g.transcript[0]);
  put_val(parts, g.v0);
  if (sp.w != VT_NONE) put_val(parts, g.w0);
  Fr z = challenge_from_random_bytes(parts);
  Fr shift_a = sp.w != VT_NONE ? r_shift.inv() : Fr::one();
  size_t base = n_srs / world, rem = n_srs % world;
  size_t lo = rank * base + ((size_t)rank < rem ? rank : rem), cnt = base + ((size_t)rank < rem ? 1 : 0);
  std::vector<Fr> qa = kzg_quotient(tinv, shift_a, z, n_srs);
  std::vector<Fr> qb;
  if (sp.w != VT_NONE) qb = kzg_quotient(g.transcript, Fr::one(), z, n_srs);
  void* d;
  OK(scratch(ctx, 12, 2 * n_srs * sizeof(Fr) + (size_t)(2 * world + 4) * 288 + 2048, &d));
  char* qa_d = (char*)d;
  char* qb_d = qa_d + ((n_srs * sizeof(Fr) + 255) & ~(size_t)255);
  char* part = qb_d + ((n_srs * sizeof(Fr) + 255) & ~(size_t)255);  // [G2 192 | G1 96]
  char* gath = part + 288;                                          // [world][288]
  char* cmp = gath + (size_t)world * 288;                           // compact [world] G2, then [world] G1
  char* res = cmp + (size_t)world * 288;
  CU(cudaMemsetAsync(part, 0, 288, ctx->stream));
  CU(cudaMemcpyAsync(qa_d, qa.data() + lo, cnt * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  ripp_ctx* kid = nullptr;
  if (sp.w != VT_NONE) {
    kid = ripp_child(ctx, 5);
    if (!kid) return fail(RIPP_ERR_CUDA, "child context");
    CU(cudaMemcpyAsync(qb_d, qb.data() + lo, cnt * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    OK(ripp_fork(ctx, kid));
    OK(ripp_msm_g1_dev(kid, (const char*)srs_g1 + lo * 96, qb_d, cnt, part + 192));
  }
  OK(ripp_msm_g2_dev(ctx, (const char*)srs_g2 + lo * 192, qa_d, cnt, part));
  if (kid) OK(ripp_join(ctx, kid));
  OK(ripp_all_gather_internal(ctx, part, 288, gath));
  for (int r = 0; r < world; r++) {
    CU(cudaMemcpyAsync(cmp + r * 192, gath + (size_t)r * 288, 192, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaMemcpyAsync(cmp + (size_t)world * 192 + r * 96, gath + (size_t)r * 288 + 192, 96, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  OK(ripp_seg_sum_dev(ctx, 2, cmp, world, 1, res));
  if (kid) OK(ripp_seg_sum_dev(ctx, 1, cmp + (size_t)world * 192, world, 1, res + 192));
  G2Aff open_a;
  G1Aff open_b;
  CU(cudaMemcpyAsync(&open_a, res, 192, cudaMemcpyDeviceToHost, ctx->stream));
  if (kid) CU(cudaMemcpyAsync(&open_b, res + 192, 96, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *proof = g.proof;
  put_val(*proof, g.v0);
  if (sp.w != VT_NONE) {
    put_val(*proof, g.w0);
    put_g2(*proof, open_a);
    put_g1(*proof, open_b);
  } else {
    put_g2(*proof, open_a);
  }
  return RIPP_OK;
}

// ---- aggregate_proofs for ONE batch of n proofs partitioned over the ranks -------------------------------------------
// a, b, c: this rank's CYCLIC shares (n / world elements); srs_g1 / srs_g2: the full 2 n - 1 powers on every rank.
// Products of the prologue: Miller partials of the local shares, one all-gather each; the two recursions advance in
// lock step (sharded_rounds with two instances), their tails finish concurrently on two host threads, the KZG
// openings are sharded by slices.  Same bytes as ripp_tipp_aggregate_dev on one GPU.
static int sharded_products(ripp_ctx* ctx, int k, const void* const* g1, const void* const* g2, size_t m, Val* out) {
  ripp_ctx* top = top_ctx(ctx);
  const int world = top->world;
  void* buf;
  OK(scratch(ctx, 27, (size_t)(2 * world + 2) * k * 576 + 1024, &buf));
  char* part = (char*)buf;
  char* gath = part + (size_t)k * 576;
  char* tr = gath + (size_t)world * k * 576;
  char* res = tr + (size_t)world * k * 576;
  OK(ripp_pairing_batch_l6(ctx, k, g1, g2, m, part, false));
  OK(ripp_all_gather_internal(ctx, part, (size_t)k * 576, gath));
  size_t total = (size_t)world * k * 144;
  k_blob_transpose<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t*)gath, world, k, 144, (uint32_t*)tr);
  LAUNCHED(ctx);
  OK(ripp_final_exp_l6(ctx, tr, (uint32_t)world, res, k));
  uint8_t* pin;
  OK(pinned(ctx, &pin));
  CU(cudaMemcpyAsync(pin, res, (size_t)k * 576, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < k; i++) {
    out[i].t = VT_GT;
    memcpy(out[i].raw, pin + (size_t)i * 576, 576);
  }
  return RIPP_OK;
}

extern "C" int ripp_tipp_aggregate_sharded_dev(ripp_ctx* ctx, const void* srs_g1_dev, const void* srs_g2_dev, const void* a_dev,
                                               const void* b_dev, const void* c_dev, size_t n_total, size_t tail_len,
                                               uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
  if (!ctx || !srs_g1_dev || !srs_g2_dev || !a_dev || !b_dev || !c_dev) return fail(RIPP_ERR_ARG, "null argument");
  const size_t n = n_total;
  if (n < 2 || (n & (n - 1))) return fail(RIPP_ERR_NOT_POW2, "number of proofs must be a power of two, at least 2");
  CU(cudaSetDevice(ctx->device));
  ripp_ctx* top = top_ctx(ctx);
  const int world = top->world, rank = top->rank;
  if (n % world || n / world < 1) return fail(RIPP_ERR_ARG, "the number of proofs must be a multiple of the world size");
  const size_t m = n / world;
  cudaStream_t st = ctx->stream;
  void* ws;
  size_t o_ck1 = 0, o_ck2 = o_ck1 + m * 192, o_ar = o_ck2 + m * 96, o_ck1r = o_ar + m * 96, o_pw = o_ck1r + m * 192,
         o_pwi = o_pw + m * 32, o_res = o_pwi + m * 32;
  OK(scratch(ctx, 13, o_res + 4096, &ws));
  char* W = (char*)ws;
  unsigned nb = (unsigned)((m + 127) / 128);
  // commitment keys = even SRS powers (tipa/mod.rs:114-118), this rank's cyclic share: global index j g + k -> power 2 (j g + k)
  k_gather_stride2<G2Aff><<<nb, 128, 0, st>>>((const G2Aff*)srs_g2_dev, m, (G2Aff*)(W + o_ck1), 2 * (size_t)world, 2 * (size_t)rank);
  LAUNCHED(ctx);
  k_gather_stride2<G1Aff><<<nb, 128, 0, st>>>((const G1Aff*)srs_g1_dev, m, (G1Aff*)(W + o_ck2), 2 * (size_t)world, 2 * (size_t)rank);
  LAUNCHED(ctx);
  const void* ck1 = W + o_ck1;
  const void* ck2 = W + o_ck2;
  Val com[3];
  {
    const void* g1[3] = {a_dev, ck2, c_dev};
    const void* g2[3] = {ck1, b_dev, ck1};
    OK(sharded_products(ctx, 3, g1, g2, m, com));
  }
  Bytes parts;
  for (int i = 0; i < 3; i++) put_val(parts, com[i]);
  Fr r = challenge_from_random_bytes(parts);
  Fr r_inv = r.inv();
  k_fr_powers<<<nb, 128, 0, st>>>(r, m, (Fr*)(W + o_pw), (Fr*)(W + o_pwi), r_inv, (size_t)world, (size_t)rank);
  LAUNCHED(ctx);
  ripp_ctx* pk = ripp_child(ctx, 5);
  ripp_ctx* mk = ripp_child(ctx, 4);
  ripp_ctx* ka = ripp_child(ctx, 6);
  ripp_ctx* kc = ripp_child(ctx, 7);
  if (!pk || !mk || !ka || !kc) return fail(RIPP_ERR_CUDA, "child context");
  OK(ripp_fork(ctx, pk));
  OK(ripp_g2_scale_dev(pk, ck1, W + o_pwi, m, W + o_ck1r));
  OK(ripp_g1_scale_dev(ctx, a_dev, W + o_pw, m, W + o_ar));
  OK(ripp_fork(ctx, mk));
  OK(ripp_msm_g1_dev(mk, c_dev, W + o_pw, m, W + o_res));  // agg_c partial
  OK(ripp_join(ctx, pk));
  Val ipv[2];
  {
    const void* g1[2] = {W + o_ar, W + o_ar};
    const void* g2[2] = {b_dev, W + o_ck1r};
    OK(sharded_products(ctx, 2, g1, g2, m, ipv));
  }
  if (memcmp(ipv[1].raw, com[0].raw, 576) != 0)
    return fail(RIPP_ERR_INNER_PRODUCT, "com_a != IP(a_r, ck_1_r) (groth16_aggregation.rs:133-136)");
  // agg_c = MSM(c, r_vec): partial points, one all-gather
  Val agg_c;
  agg_c.t = VT_G1;
  memset(agg_c.raw, 0, 576);
  {
    OK(ripp_join(ctx, mk));
    void* gb;
    OK(scratch(ctx, 27, (size_t)(world + 2) * 96 + 256, &gb));
    OK(ripp_all_gather_internal(ctx, W + o_res, 96, gb));
    OK(ripp_seg_sum_dev(ctx, 1, gb, world, 1, (char*)gb + (size_t)world * 96));
    CU(cudaMemcpyAsync(agg_c.raw, (char*)gb + (size_t)world * 96, 96, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  // the two recursions in lock step
  ShardedInst in[2];
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  gipa_spec(RIPP_GIPA_PAIRING, &in[0].sp);
  gipa_spec(RIPP_GIPA_MULTIEXP_SSM, &in[1].sp);
  in[0].cx = ka;
  in[1].cx = kc;
  const void* srcs[2][4] = {{W + o_ar, b_dev, W + o_ck1r, ck2}, {c_dev, W + o_pw, ck1, nullptr}};
  for (int i = 0; i < 2; i++) {
    const GipaSpec& sp = in[i].sp;
    size_t sz[4] = {m * vt_size(sp.a), m * vt_size(sp.b), m * vt_size(sp.v), m * vt_size(sp.w)};
    void* work;
    OK(scratch(in[i].cx, 11, up(sz[0]) + up(sz[1]) + up(sz[2]) + up(sz[3]) + 1024, &work));
    char* p = (char*)work;
    char** dst[4] = {&in[i].A, &in[i].B, &in[i].V, &in[i].W};
    for (int t = 0; t < 4; t++) {
      if (!srcs[i][t]) {
        *dst[t] = nullptr;
        continue;
      }
      *dst[t] = p;
      CU(cudaMemcpyAsync(p, srcs[i][t], sz[t], cudaMemcpyDeviceToDevice, st));
      p += up(sz[t]);
    }
  }
  size_t n_tail = 0;
  OK(sharded_rounds(ctx, in, 2, m, tail_len ? tail_len : ((size_t)1 << 12), &n_tail));
  // tails: latency-bound, replicated on every rank, the two recursions concurrently from two host threads
  GipaOut g[2];
  int st_c = RIPP_OK, st_a = RIPP_OK;
  std::string err_c;
  OK(ripp_fork(ctx, ka));
  OK(ripp_fork(ctx, kc));
  {
    std::thread tc([&] {
      cudaSetDevice(ctx->device);
      st_c = gipa_prove(kc, in[1].sp, in[1].A, in[1].B, in[1].V, in[1].W, n_tail, &g[1], nullptr, &in[1].steps, &in[1].transcript);
      if (st_c != RIPP_OK) err_c = ripp_err_slot();
    });
    st_a = gipa_prove(ka, in[0].sp, in[0].A, in[0].B, in[0].V, in[0].W, n_tail, &g[0], nullptr, &in[0].steps, &in[0].transcript);
    tc.join();
  }
  if (st_a != RIPP_OK) return st_a;
  if (st_c != RIPP_OK) return fail(st_c, err_c);
  OK(ripp_join(ctx, ka));
  OK(ripp_join(ctx, kc));
  Bytes proof_ab, proof_c;
  OK(tipa_open_sharded(ctx, in[0].sp, g[0], srs_g1_dev, srs_g2_dev, n, r, &proof_ab));
  OK(tipa_open_sharded(ctx, in[1].sp, g[1], srs_g1_dev, srs_g2_dev, n, Fr::one(), &proof_c));
  Bytes out;
  for (int i = 0; i < 3; i++) put_val(out, com[i]);
  put_val(out, ipv[0]);
  put_val(out, agg_c);
  out.insert(out.end(), proof_ab.begin(), proof_ab.end());
  out.insert(out.end(), proof_c.begin(), proof_c.end());
  return copy_out(out, proof_out, proof_cap, proof_len);
}
