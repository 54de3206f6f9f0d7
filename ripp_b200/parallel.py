"""Multi-GPU plumbing (one process per GPU, torch.distributed): contiguous input slices per rank and
the gather-then-combine of the per-rank partials.  GT products and EC sums are not NCCL
reduction ops, so partials are all-gathered and combined locally in rank order."""
import numpy as np


def shard_bounds(n, rank, world):
    """[lo, hi) of rank's contiguous slice; slice sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_pairing_ip(ctx, g1_dev, g2_dev, n_local, out_ptr, scratch):
    """Product of pairings over vectors sharded across ranks.  g1_dev / g2_dev: this rank's slice (device);
    scratch = (partial, gathered) torch int32 tensors of 144 and world*144 elements on the same device."""
    import torch.distributed as dist

    partial, gathered = scratch
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        ctx.pairing_ip_dev(g1_dev, g2_dev, n_local, out_ptr)
        return
    ctx.miller_partial_dev(g1_dev, g2_dev, n_local, partial.data_ptr())
    dist.all_gather_into_tensor(gathered, partial)
    ctx.gt_combine_dev(gathered.data_ptr(), world, out_ptr)


# =====================================================================================================
# Sharded GIPA / TIPA provers (SURVEY.md §8e "GIPA rounds", "KZG openings")
# =====================================================================================================
# Cyclic partition: with g ranks, rank k holds the elements with global index i = j g + k (local index j) of
# all four vectors.  A round pairs global index i of the left half with i + n' of the right half
# (gipa.rs:209-217); while g divides n', both live on the same rank and at local indices j and j + n'/g, so the
# six products of the round are products over LOCAL halves, the four folds are local, and the folded element i
# stays on rank i mod g: the only exchange per round is one all-gather of six fixed-size partials (<= 576 B
# each), combined on every rank in rank order (GT: product of Miller values, then ONE final exponentiation per
# commitment; points / scalars: sum).  When one element per rank is left the vectors (g elements) are
# all-gathered once and every rank finishes the last log2(g) rounds redundantly.  The Fiat-Shamir hash is
# computed identically on every rank from the combined values, so all ranks emit the same proof bytes -- the
# bytes ripp_gipa_prove_dev emits on one GPU.
#
# Host side (this file): transcript, serialisation, hashing, torch.distributed.  Every group / field operation
# on vectors runs in the CUDA library through the C ABI (no CPU fallback; the context fails without a GPU).
import hashlib

from . import codec

_KIND_TYPES = {0: ("G1", "G2", "G2", "G1"), 1: ("G1", "Fr", "G2", "G1"), 2: ("G1", "Fr", "G2", None),
               3: ("Fr", "Fr", "G2", "G2"), 4: ("Fr", "Fr", "G2", "G1"), 5: ("Fr", "Fr", "G2", None),
               6: ("Fr", "Fr", "G1", None)}
_WORDS = {"G1": 24, "G2": 48, "Fr": 8, "GT": 144}
_SUM_ID = {"G1": 1, "G2": 2, "Fr": 3}
_DEC = {"G1": codec.g1_dec, "G2": codec.g2_dec, "Fr": codec.fr_dec, "GT": codec.gt_dec}
_SER = {"G1": codec.ser_g1, "G2": codec.ser_g2, "Fr": codec.ser_fr, "GT": codec.ser_gt}


def ip_out_type(x, y):
    """Output type of IP(x, y): pairing -> GT, placeholder -> Fr (zero), scalar x scalar -> Fr, else the point type."""
    if {x, y} == {"G1", "G2"}:
        return "GT"
    if x is None or y is None or (x == "Fr" and y == "Fr"):
        return "Fr"
    return y if x == "Fr" else x


def cyclic_share(vec, rank, world):
    """Rank's share of a global vector under the cyclic partition (elements rank, rank + world, ...)."""
    return vec[rank::world]


def gipa_challenge(prev_c, com_bytes):
    """gipa.rs:235-258: Blake2b-512(nonce_be || prev c || six commitments); x = u128_be(digest[..16]);
    returns (c, c_inv) = (x^-1, x) -- swapped as gipa.rs:253-255."""
    nonce = 0
    while True:
        d = hashlib.blake2b(nonce.to_bytes(8, "big") + codec.ser_fr(prev_c) + com_bytes, digest_size=64).digest()
        x = int.from_bytes(d[:16], "big")
        if x % codec.R:
            return pow(x, -1, codec.R), x
        nonce += 1


def challenge_from_random_bytes(parts):
    """tipa/mod.rs:195-209: hash(nonce_be || parts) until Fp::from_random_bytes accepts (32 bytes LE, top bit cleared, < r)."""
    nonce = 0
    while True:
        d = hashlib.blake2b(nonce.to_bytes(8, "big") + parts, digest_size=64).digest()
        v = int.from_bytes(d[:32], "little") & ((1 << 255) - 1)
        if v < codec.R:
            return v
        nonce += 1


class Comm:
    """All-gather of one fixed-size device blob per rank.  NCCL: device tensors straight over NVLink.  Any other
    backend (gloo in the single-GPU tests): staged through the host."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist, self.group = dist, group
        self.on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0
        self.device_native = self.on and dist.get_backend(group) == "nccl"

    def all_gather(self, t):
        """t: contiguous device tensor -> (world, *t.shape) device tensor, rank order."""
        import torch

        if self.world == 1:
            return t.unsqueeze(0)
        if self.device_native:
            out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
            self.dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1), group=self.group)
            return out
        torch.cuda.current_stream().synchronize()
        host = t.cpu()
        outs = [torch.empty_like(host) for _ in range(self.world)]
        self.dist.all_gather(outs, host, group=self.group)
        return torch.stack(outs).to(t.device)


class ShardedGIPA:
    """GIPA::prove_with_aux (gipa.rs:162-312) over vectors partitioned cyclically across the ranks of `comm`.
    Vectors are torch int32 CUDA tensors (n_local, words) in the C ABI's packed layout (affine points / Fr,
    Montgomery); the context must run on torch's current stream (ctx.set_stream)."""

    def __init__(self, kind, ctx, comm=None):
        self.kind, self.ctx, self.comm = kind, ctx, comm or Comm()
        self.ta, self.tb, self.tv, self.tw = _KIND_TYPES[kind]
        # (x type, y type) of the three products of a commitment triple: IP(A, v), IP(w, B), IP(A, B)
        self.prod_types = [(self.ta, self.tv), (self.tw, self.tb), (self.ta, self.tb)]
        self.out_types = [ip_out_type(x, y) for x, y in self.prod_types]
        self.tail_len = 1 << 12
        self._helpers = None  # extra contexts (own stream + scratch) so MSM-type products and folds overlap

    # -- stream fork / join between the main context's stream (torch's current stream) and helper contexts ----
    def _helper(self, i):
        import torch

        if self._helpers is None:
            self._helpers = []
            for _ in range(3):
                c = type(self.ctx)(self.ctx.device)
                st = torch.cuda.Stream()
                c.set_stream(st.cuda_stream)
                self._helpers.append((c, st))
        return self._helpers[i % 3]

    def _fork(self, st):
        import torch

        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        st.wait_event(ev)

    def _join(self, st):
        import torch

        ev = torch.cuda.Event()
        ev.record(st)
        torch.cuda.current_stream().wait_event(ev)

    # -- six partial products over local halves -> combined values on every rank ------------------------
    def _round_products(self, A, B, V, W, split, count_ranks):
        import torch

        ctx = self.ctx
        dev = A.device
        parts = torch.zeros((6, 144), dtype=torch.int32, device=dev)
        lo = lambda t: None if t is None else t[:split]
        hi = lambda t: None if t is None else t[split:2 * split]
        # gipa.rs:220-231: com_1 = (IP(A_R, v_L), IP(w_R, B_L), IP(A_R, B_L)); com_2 = (IP(A_L, v_R), IP(w_L, B_R), IP(A_L, B_R))
        xs = [hi(A), hi(W), hi(A), lo(A), lo(W), lo(A)]
        ys = [lo(V), lo(B), lo(B), hi(V), hi(B), hi(B)]
        types = self.prod_types * 2
        outs = self.out_types * 2
        gt_slots = [i for i in range(6) if outs[i] == "GT"]
        # MSM-type products first, each on a helper stream; then the pairing batch on the main stream (they overlap)
        used, nh = [], 0
        for i in range(6):
            tx, ty = types[i]
            if outs[i] == "GT" or tx is None or ty is None:
                continue  # placeholder commitment: Fr zero
            if tx == "Fr" and ty == "Fr":
                ctx.scalar_ip_dev(xs[i].data_ptr(), ys[i].data_ptr(), split, parts[i].data_ptr())
                continue
            hc, hst = self._helper(nh)
            nh += 1
            if hst not in used:
                self._fork(hst)
                used.append(hst)
            pts, sc = (ys[i], xs[i]) if tx == "Fr" else (xs[i], ys[i])
            fn = hc.msm_g1_dev if outs[i] == "G1" else hc.msm_g2_dev
            fn(pts.data_ptr(), sc.data_ptr(), split, parts[i].data_ptr())
        tmp = None
        if gt_slots:
            g1p, g2p = [], []
            for i in gt_slots:
                x_is_g1 = types[i][0] == "G1"
                g1p.append((xs[i] if x_is_g1 else ys[i]).data_ptr())
                g2p.append((ys[i] if x_is_g1 else xs[i]).data_ptr())
            tmp = torch.zeros((len(gt_slots), 144), dtype=torch.int32, device=dev)
            ctx.miller_partial_batch_dev(g1p, g2p, split, tmp.data_ptr())
        for hst in used:
            self._join(hst)
        if tmp is not None:
            parts[gt_slots] = tmp
        if count_ranks > 1:
            gathered = self.comm.all_gather(parts).permute(1, 0, 2).contiguous()  # (6, world, 144)
        else:
            gathered = parts.unsqueeze(1).contiguous()
        combined = torch.zeros((6, 144), dtype=torch.int32, device=dev)
        cnt = gathered.shape[1]
        if gt_slots:
            src = gathered[gt_slots].contiguous()
            dst = torch.zeros((len(gt_slots), 144), dtype=torch.int32, device=dev)
            ctx.gt_combine_batch_dev(src.data_ptr(), cnt, len(gt_slots), dst.data_ptr())
            combined[gt_slots] = dst
        for t in ("G1", "G2", "Fr"):
            slots = [i for i in range(6) if outs[i] == t and types[i][0] is not None and types[i][1] is not None]
            if not slots:
                continue
            w = _WORDS[t]
            src = gathered[slots][:, :, :w].contiguous()
            dst = torch.zeros((len(slots), w), dtype=torch.int32, device=dev)
            ctx.seg_sum_dev(_SUM_ID[t], src.data_ptr(), cnt, len(slots), dst.data_ptr())
            combined[slots, :w] = dst
        torch.cuda.current_stream().synchronize()
        host = combined.cpu().numpy().view(np.uint32)
        return [_DEC[outs[i]](host[i, :_WORDS[outs[i]]]) for i in range(6)], outs

    def _ser_triples(self, vals, outs):
        b = b""
        for i in range(6):
            s = _SER[outs[i]](vals[i])
            b += codec.ser_identity_output(s) if i % 3 == 2 else s
        return b

    def _fold(self, t, typ, split, c, ctx=None):
        if t is None:
            return
        ctx = ctx or self.ctx
        cw = codec.fr_enc(c).copy()
        fn = {"G1": ctx.g1_fold_dev, "G2": ctx.g2_fold_dev, "Fr": ctx.fr_fold_dev}[typ]
        fn(t[split:2 * split].data_ptr(), t[:split].data_ptr(), cw, split, t[:split].data_ptr())

    def prove_with_aux_dev(self, a, b, v, w=None, tail_len=None):
        """-> (GIPAProof bytes, r_transcript ints (reversed, as GIPAAux), ck_base bytes).

        Rounds run partitioned while the GLOBAL vector is longer than tail_len (default 2^12: below that every kernel
        of a round is a latency chain of a few warps and partitioning buys nothing); then the remaining vectors are
        all-gathered once and every rank finishes with the resident single-GPU prover continuing the same
        transcript (ripp_gipa_prove_resume_dev), which overlaps its MSMs, folds and pairing batches on child streams."""
        import torch

        world = self.comm.world
        m = a.shape[0]
        if m == 0 or m & (m - 1) or world & (world - 1):
            raise ValueError("local length and world size must be powers of two (gipa.rs:116-122)")
        tail_len = self.tail_len if tail_len is None else tail_len
        A, B, V = a.clone(), b.clone(), v.clone()  # gipa.rs:175-176 clones all four vectors
        W = None if self.tw is None else w.clone()
        steps, transcript = [], []
        while m > 1 and m * world > max(tail_len, world):
            split = m // 2
            vals, outs = self._round_products(A, B, V, W, split, world)
            com_bytes = self._ser_triples(vals, outs)
            c, c_inv = gipa_challenge(transcript[-1] if transcript else 0, com_bytes)
            # gipa.rs:261-291: A <- A_R c + A_L, B <- B_R c^-1 + B_L, v <- v_R c^-1 + v_L, w <- w_R c + w_L
            for i, (t, typ, sc) in enumerate(((A, self.ta, c), (B, self.tb, c_inv), (V, self.tv, c_inv), (W, self.tw, c))):
                if i == 0 or t is None:
                    self._fold(t, typ, split, sc)          # main stream
                else:
                    hc, hst = self._helper(i - 1)
                    self._fork(hst)
                    self._fold(t, typ, split, sc, hc)
            for i in range(3):
                self._join(self._helper(i)[1])
            steps.append(com_bytes)
            transcript.append(c)
            m = split

        state = self._gather_tail((A, B, V, W), m, steps, transcript)
        return state if getattr(self, "_defer", False) else self._finish(state)

    def _gather_tail(self, vecs, m, steps, transcript):
        """all-gather what is left into global order (global index j g + k = local index j of rank k): collective."""
        import torch

        world = self.comm.world

        def whole(t):
            if t is None:
                return None
            t = t[:m].contiguous()
            return t if world == 1 else self.comm.all_gather(t).permute(1, 0, 2).reshape(m * world, t.shape[1]).contiguous()

        vecs = tuple(whole(t) for t in vecs)
        torch.cuda.current_stream().synchronize()
        return vecs, m * world, steps, transcript

    def _finish(self, state, ctx=None):
        """... and finish on every rank with the resident prover, continuing the transcript: local, no collective (so two
        instances may finish concurrently from two host threads on two contexts)."""
        (A, B, V, W), n_tail, steps, transcript = state
        ctx = ctx or self.ctx
        prev = codec.fr_enc(transcript[-1]).copy() if transcript else None
        tail, tail_tr, ck_base = ctx.gipa_prove_resume_dev(self.kind, A.data_ptr(), B.data_ptr(), V.data_ptr(),
                                                           None if W is None else W.data_ptr(), n_tail, prev)
        k_tail = int.from_bytes(tail[:8], "little")
        if not steps:  # everything was proved by the resident prover
            return tail, codec.fr_vec_dec(tail_tr), ck_base
        base_len = len(tail) - 8 - len(steps[0]) * k_tail
        tail_steps, r_base = tail[8:len(tail) - base_len], tail[len(tail) - base_len:]
        proof = (k_tail + len(steps)).to_bytes(8, "little") + tail_steps + b"".join(reversed(steps)) + r_base
        return proof, codec.fr_vec_dec(tail_tr) + transcript[::-1], ck_base

    def rounds_dev(self, a, b, v, w=None, tail_len=None):
        """The collective part only: partitioned rounds + the gather of the tail vectors.  -> state for finish()."""
        self._defer = True
        try:
            return self.prove_with_aux_dev(a, b, v, w, tail_len)
        finally:
            self._defer = False

    def finish(self, state, ctx=None):
        return self._finish(state, ctx)


class ShardedTIPA:
    """TIPA::prove_with_srs_shift (tipa/mod.rs:176-231) / TIPAWithSSM::prove_with_structured_scalar_message
    (structured_scalar_message.rs:211-268) with the GIPA state partitioned cyclically and the two KZG opening MSMs
    (tipa/mod.rs:304-337) sharded by contiguous slices of the SRS powers."""

    def __init__(self, kind, ctx, comm=None):
        self.gipa = ShardedGIPA(kind, ctx, comm)
        self.ctx, self.comm = ctx, self.gipa.comm

    def _open(self, group, srs_slice, lo, n_srs, transcript, r_shift, z):
        import torch

        q = self.ctx.kzg_quotient(codec.fr_vec_enc(transcript), codec.fr_enc(r_shift).copy(), codec.fr_enc(z).copy(), n_srs)
        n_loc = srs_slice.shape[0]
        w = 24 if group == 1 else 48
        part = torch.zeros(w, dtype=torch.int32, device=srs_slice.device)
        if n_loc:
            qd = torch.from_numpy(q[lo:lo + n_loc].view(np.int32).copy()).to(srs_slice.device)
            (self.ctx.msm_g1_dev if group == 1 else self.ctx.msm_g2_dev)(srs_slice.data_ptr(), qd.data_ptr(), n_loc, part.data_ptr())
        gathered = self.comm.all_gather(part).contiguous()
        out = torch.zeros(w, dtype=torch.int32, device=srs_slice.device)
        self.ctx.seg_sum_dev(group, gathered.data_ptr(), gathered.shape[0], 1, out.data_ptr())
        torch.cuda.current_stream().synchronize()
        return (codec.g1_dec if group == 1 else codec.g2_dec)(out.cpu().numpy().view(np.uint32))

    def prove_with_srs_shift(self, srs_g1_slice, srs_g2_slice, slice_lo, n_srs, a, b, v, w=None, r_shift=1):
        """srs_g{1,2}_slice: this rank's CONTIGUOUS slice [slice_lo, slice_lo + len) of g^(alpha^i) / h^(beta^i),
        i < n_srs = 2 n - 1; a, b, v, w: this rank's cyclic shares.  -> TIPAProof / TIPAWithSSMProof bytes."""
        return self.open_keys(srs_g1_slice, srs_g2_slice, slice_lo, n_srs, self.gipa.prove_with_aux_dev(a, b, v, w), r_shift)

    def open_keys(self, srs_g1_slice, srs_g2_slice, slice_lo, n_srs, gipa_out, r_shift=1):
        """The KZG challenge and the sharded openings of the final commitment keys (tipa/mod.rs:191-229), given the GIPA output."""
        g = self.gipa
        proof, transcript, ck_base = gipa_out
        ssm = g.tw is None
        tinv = [pow(x, -1, codec.R) for x in transcript]
        z = challenge_from_random_bytes(codec.ser_fr(transcript[0]) + ck_base)
        shift_a = 1 if ssm else pow(r_shift, -1, codec.R)
        open_a = self._open(2, srs_g2_slice, slice_lo, n_srs, tinv, shift_a, z)
        if ssm:
            return proof + ck_base + codec.ser_g2(open_a)
        open_b = self._open(1, srs_g1_slice, slice_lo, n_srs, transcript, 1, z)
        return proof + ck_base + codec.ser_g2(open_a) + codec.ser_g1(open_b)


def _sharded_products(ctx, comm, g1_list, g2_list, n_local):
    """prod_i e(g1[i], g2[i]) over vectors partitioned across the ranks, for several vector pairs at once:
    one Miller launch on the local shares, one all-gather, one combine + final-exponentiation launch."""
    import torch

    k = len(g1_list)
    dev = g1_list[0].device
    part = torch.zeros((k, 144), dtype=torch.int32, device=dev)
    ctx.miller_partial_batch_dev([t.data_ptr() for t in g1_list], [t.data_ptr() for t in g2_list], n_local, part.data_ptr())
    gathered = comm.all_gather(part).permute(1, 0, 2).contiguous()  # (k, world, 144)
    out = torch.zeros((k, 144), dtype=torch.int32, device=dev)
    ctx.gt_combine_batch_dev(gathered.data_ptr(), gathered.shape[1], k, out.data_ptr())
    torch.cuda.current_stream().synchronize()
    host = out.cpu().numpy().view(np.uint32)
    return [codec.gt_dec(host[i]) for i in range(k)]


def sharded_aggregate_proofs(ctx, comm, srs_g1_slice, srs_g2_slice, slice_lo, n, ck1, ck2, a, b, c, tail_len=None):
    """aggregate_proofs (applications/groth16_aggregation.rs:77-160) for ONE batch of n Groth16 proofs partitioned
    over the ranks of `comm`.  a, b, c (the proofs' A, B, C) and ck1 = h^(beta^(2i)), ck2 = g^(alpha^(2i)) are this
    rank's CYCLIC shares (global index j g + k at local index j); srs_g{1,2}_slice its CONTIGUOUS slice
    [slice_lo, ...) of the 2 n - 1 SRS powers (for the KZG openings).  All ranks return the AggregateProof bytes
    ripp_tipp_aggregate_dev returns on one GPU."""
    import torch

    world, rank = comm.world, comm.rank
    m = a.shape[0]
    assert m * world == n and n & (n - 1) == 0
    # :100-102 com_a = IP(a, ck_1), com_b = IP(ck_2, b), com_c = IP(c, ck_1)
    com_a, com_b, com_c = _sharded_products(ctx, comm, [a, ck2, c], [ck1, b, ck1], m)
    # :105-116 r
    r = challenge_from_random_bytes(codec.ser_gt(com_a) + codec.ser_gt(com_b) + codec.ser_gt(com_c))
    r_inv = pow(r, -1, codec.R)
    # :118-131 r_vec and the rescaled vectors, for this rank's global indices j g + k
    step, step_inv = pow(r, world, codec.R), pow(r_inv, world, codec.R)
    rv, rvi = [pow(r, rank, codec.R)], [pow(r_inv, rank, codec.R)]
    for _ in range(m - 1):
        rv.append(rv[-1] * step % codec.R)
        rvi.append(rvi[-1] * step_inv % codec.R)
    r_vec = torch.from_numpy(codec.fr_vec_enc(rv).view(np.int32)).to(a.device)
    r_vec_inv = torch.from_numpy(codec.fr_vec_enc(rvi).view(np.int32)).to(a.device)
    a_r, ck1_r = torch.empty_like(a), torch.empty_like(ck1)
    ctx.g1_scale_dev(a.data_ptr(), r_vec.data_ptr(), m, a_r.data_ptr())
    ctx.g2_scale_dev(ck1.data_ptr(), r_vec_inv.data_ptr(), m, ck1_r.data_ptr())
    # :124 ip_ab and the :133-136 sanity product
    ip_ab, check = _sharded_products(ctx, comm, [a_r, a_r], [b, ck1_r], m)
    if check != com_a:
        raise _lib_error("com_a != IP(a_r, ck_1_r) (groth16_aggregation.rs:133-136)")
    # :125 agg_c = MSM(c, r_vec)
    part = torch.zeros(24, dtype=torch.int32, device=a.device)
    ctx.msm_g1_dev(c.data_ptr(), r_vec.data_ptr(), m, part.data_ptr())
    gathered = comm.all_gather(part).contiguous()
    agg = torch.zeros(24, dtype=torch.int32, device=a.device)
    ctx.seg_sum_dev(1, gathered.data_ptr(), gathered.shape[0], 1, agg.data_ptr())
    torch.cuda.current_stream().synchronize()
    agg_c = codec.g1_dec(agg.cpu().numpy().view(np.uint32))
    # :138-149 the two TIPA proofs
    n_srs = 2 * n - 1
    t_ab, t_c = ShardedTIPA(0, ctx, comm), ShardedTIPA(2, ctx, comm)
    if tail_len is not None:
        t_ab.gipa.tail_len = t_c.gipa.tail_len = tail_len
    # partitioned rounds of both recursions (collectives, one after the other), then the two latency-bound tails
    # concurrently: two host threads, two contexts (the library releases the GIL inside the C call)
    import threading

    st_ab = t_ab.gipa.rounds_dev(a_r, b, ck1_r, ck2)
    st_c = t_c.gipa.rounds_dev(c, r_vec, ck1, None)
    side_ctx, side_stream = t_c.gipa._helper(0)
    t_c.gipa._fork(side_stream)
    res = {}

    def run(name, gipa, state, cx):
        try:
            res[name] = gipa.finish(state, cx)
        except Exception as e:  # re-raised below on the calling thread
            res[name] = e

    th = threading.Thread(target=run, args=("c", t_c.gipa, st_c, side_ctx))
    th.start()
    run("ab", t_ab.gipa, st_ab, ctx)
    th.join()
    t_c.gipa._join(side_stream)
    for v in res.values():
        if isinstance(v, Exception):
            raise v
    proof_ab = t_ab.open_keys(srs_g1_slice, srs_g2_slice, slice_lo, n_srs, res["ab"], r_shift=r)
    proof_c = t_c.open_keys(srs_g1_slice, srs_g2_slice, slice_lo, n_srs, res["c"])
    return (codec.ser_gt(com_a) + codec.ser_gt(com_b) + codec.ser_gt(com_c) + codec.ser_gt(ip_ab) + codec.ser_g1(agg_c)
            + proof_ab + proof_c)


def _lib_error(msg):
    from . import _lib

    return _lib.RippError(_lib.RIPP_ERR_INNER_PRODUCT, msg)


def init_library_comm(ctx, group=None):
    """Gives `ctx` the library's own NCCL communicator (comm.cu) over the ranks of a torch.distributed group: rank 0
    draws the NCCL unique id, the bootstrap group (any backend) broadcasts its 128 bytes, every rank joins.  After this
    the sharded entry points of the C ABI (ripp_*_sharded_dev) run their collectives inside the library."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        ctx.comm_init(b"\0" * 128, 0, 1)
        return 0, 1
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init(box[0], rank, world)
    return rank, world
