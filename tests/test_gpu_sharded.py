"""Sharded GIPA / TIPA provers (ripp_b200/parallel.py): the proof bytes of the cyclically
partitioned prover must equal the single-GPU prover's (which equal the oracle's, tests/test_gpu_protocols.py).

* world = 1: the host-driven round loop over the batch primitives (Miller partials without final
  exponentiation, batched combine, segment sums) against ripp_gipa_prove_dev / ripp_tipa_prove_dev;
* world = 2: two processes, gloo all-gather staged through the host, both ranks on this box's GPU 0 (NCCL refuses
  two ranks on one device; with one GPU per rank the same code path gathers device tensors over NCCL)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

KINDS = [0, 1, 2, 3, 4, 5, 6]


def _vectors(ctx, kind, n):
    """Global input vectors as numpy words, generated on the GPU from the synthetic scalar streams."""
    from ripp_b200 import synth
    from ripp_b200.parallel import _KIND_TYPES

    out = []
    for t, tag in zip(_KIND_TYPES[kind], ("sh-a", "sh-b", "sh-v", "sh-w")):
        if t == "G1":
            out.append(synth.g1_points_dev(ctx, tag, n).download((n, 24)))
        elif t == "G2":
            out.append(synth.g2_points_dev(ctx, tag, n).download((n, 48)))
        elif t == "Fr":
            out.append(synth.scalars_mont(tag, n))
        else:
            out.append(None)
    return out


def _cuda(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


def _single_gpu_proof(ctx, kind, vecs, n):
    bufs = [None if v is None else ctx.to_device(v) for v in vecs]
    return ctx.gipa_prove_dev(kind, bufs[0], bufs[1], bufs[2], bufs[3], n)


def _with_stream(ctx):
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    ctx.set_stream(s.cuda_stream)
    return s


@pytest.mark.parametrize("kind", KINDS)
def test_host_driven_gipa_matches_device_prover(ctx, kind):
    from ripp_b200 import codec
    from ripp_b200.parallel import ShardedGIPA

    n = 8
    vecs = _vectors(ctx, kind, n)
    want_proof, want_tr, want_ck = _single_gpu_proof(ctx, kind, vecs, n)
    _with_stream(ctx)
    try:
        sg = ShardedGIPA(kind, ctx)
        proof, tr, ck = sg.prove_with_aux_dev(*[_cuda(v) for v in vecs], tail_len=2)  # two host-driven rounds + tail
        whole = sg.prove_with_aux_dev(*[_cuda(v) for v in vecs])                     # default: all in the resident prover
    finally:
        torch.cuda.synchronize()
        ctx.set_stream(None)
        torch.cuda.set_stream(torch.cuda.default_stream())
    assert proof == want_proof
    assert tr == codec.fr_vec_dec(want_tr)
    assert ck == want_ck
    assert whole == (proof, tr, ck)


def _srs(ctx, n):
    from ripp_b200 import codec, synth

    alpha, beta = synth.scalar("srs-alpha", 0), synth.scalar("srs-beta", 0)
    pa, pb = [1], [1]
    for _ in range(2 * n - 2):
        pa.append(pa[-1] * alpha % codec.R)
        pb.append(pb[-1] * beta % codec.R)
    return synth._gen_dev(ctx, 1, pa).download((2 * n - 1, 24)), synth._gen_dev(ctx, 2, pb).download((2 * n - 1, 48))


def _tipa_case(ctx, kind, n):
    """(vectors with the TIPA keys, srs_g1, srs_g2, r_shift, single-GPU proof bytes)"""
    from ripp_b200 import codec, synth

    vecs = _vectors(ctx, kind, n)
    s1, s2 = _srs(ctx, n)
    vecs[2] = np.ascontiguousarray(s2[::2])                      # ck_a = even powers of h^beta (tipa/mod.rs:114-118)
    if vecs[3] is not None:
        vecs[3] = np.ascontiguousarray(s1[::2])
    r_shift = 1 if vecs[3] is None else synth.scalar("shift", 0)
    if r_shift != 1:  # shifted key ck_a[i] * r^-i, as groth16_aggregation.rs:127-131 builds it
        sc = ctx.to_device(codec.fr_vec_enc([pow(r_shift, -i, codec.R) for i in range(n)]))
        kd, out = ctx.to_device(vecs[2]), ctx.alloc(n * 192)
        ctx.g2_scale_dev(kd, sc, n, out)
        ctx.sync()
        vecs[2] = out.download((n, 48))
    bufs = [None if v is None else ctx.to_device(v) for v in vecs]
    want = ctx.tipa_prove_dev(kind, ctx.to_device(s1), ctx.to_device(s2), bufs[0], bufs[1], bufs[2], bufs[3], n,
                              codec.fr_enc(r_shift).copy())
    return vecs, s1, s2, r_shift, want


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_host_driven_tipa_matches_device_prover(ctx, kind):
    from ripp_b200.parallel import ShardedTIPA

    n = 8
    vecs, s1, s2, r_shift, want = _tipa_case(ctx, kind, n)
    _with_stream(ctx)
    try:
        st = ShardedTIPA(kind, ctx)
        st.gipa.tail_len = 2
        got = st.prove_with_srs_shift(_cuda(s1), _cuda(s2), 0, 2 * n - 1, *[_cuda(v) for v in vecs], r_shift=r_shift)
    finally:
        torch.cuda.synchronize()
        ctx.set_stream(None)
        torch.cuda.set_stream(torch.cuda.default_stream())
    assert got == want


def _worker(rank, world, port, kind, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ripp_b200 import _lib
        from ripp_b200.parallel import Comm, ShardedTIPA, cyclic_share, shard_bounds

        torch.cuda.set_device(0)
        ctx = _lib.Context(0)
        vecs, s1, s2, r_shift, want = _tipa_case(ctx, kind, n)
        _with_stream(ctx)
        lo, hi = shard_bounds(2 * n - 1, rank, world)
        shares = [None if v is None else _cuda(np.ascontiguousarray(cyclic_share(v, rank, world))) for v in vecs]
        st = ShardedTIPA(kind, ctx, Comm())
        st.gipa.tail_len = 2  # partitioned rounds down to one element per rank, then the gathered tail
        got = st.prove_with_srs_shift(_cuda(s1[lo:hi]), _cuda(s2[lo:hi]), lo, 2 * n - 1, *shares, r_shift=r_shift)
        torch.cuda.synchronize()
        q.put((rank, got == want, len(got)))
    except Exception as e:  # surface the failure in the parent
        import traceback

        q.put((rank, False, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,world", [(0, 2), (1, 2), (2, 4)])
def test_sharded_tipa_matches_single_gpu(kind, world):
    import torch.multiprocessing as mp

    n = 16
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29600 + (os.getpid() + 7 * kind) % 300
    procs = [mpc.Process(target=_worker, args=(r, world, port, kind, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    for r in res:
        assert r[1] is True, r[2]


def _agg_case(ctx, n):
    from ripp_b200 import synth

    inst = synth.tipp_instance_dev(ctx, n)
    want = ctx.tipp_aggregate_dev(inst["srs_g1"], inst["srs_g2"], inst["a"], inst["b"], inst["c"], n)
    host = {"s1": inst["srs_g1"].download((2 * n - 1, 24)), "s2": inst["srs_g2"].download((2 * n - 1, 48)),
            "a": inst["a"].download((n, 24)), "b": inst["b"].download((n, 48)), "c": inst["c"].download((n, 24))}
    return host, want, inst


def _agg_worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ripp_b200 import _lib
        from ripp_b200.parallel import Comm, cyclic_share, shard_bounds, sharded_aggregate_proofs

        torch.cuda.set_device(0)
        ctx = _lib.Context(0)
        h, want, inst = _agg_case(ctx, n)
        _with_stream(ctx)
        lo, hi = shard_bounds(2 * n - 1, rank, world)
        sh = lambda v: _cuda(np.ascontiguousarray(cyclic_share(v, rank, world)))
        got = sharded_aggregate_proofs(ctx, Comm(), _cuda(h["s1"][lo:hi]), _cuda(h["s2"][lo:hi]), lo, n, sh(h["s2"][::2][:n]),
                                       sh(h["s1"][::2][:n]), sh(h["a"]), sh(h["b"]), sh(h["c"]), tail_len=2)
        torch.cuda.synchronize()
        ok = got == want and ctx.tipp_verify_aggregate(inst["vsrs"], inst["vk"], inst["inputs"], got)
        q.put((rank, bool(ok), len(got)))
    except Exception:
        import traceback

        q.put((rank, False, traceback.format_exc()))
    finally:
        if world > 1:
            dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_aggregate_matches_single_gpu(world):
    """aggregate_proofs of ONE batch partitioned over the ranks: same AggregateProof bytes, accepted by the GPU verifier."""
    import torch.multiprocessing as mp

    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29900 + os.getpid() % 90 + world
    procs = [mpc.Process(target=_agg_worker, args=(r, world, port, 16, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] is True, r[2]
