"""Multi-rank logic on CPU (gloo, world_size 2): sharding of the input vectors by contiguous slices,
all-gather of one fixed-size partial per rank, deterministic rank-order combine (SURVEY.md §8e,
DESIGN.md §5).  The per-rank leaf work is stood in for by the oracle so the test needs no GPU; the
`-m gpu` suite checks that device partials combine to the whole product (test_sharded_partials_combine)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np

    from oracle import bls12_381 as E
    from oracle import synth as OS
    from ripp_b200 import codec as C
    from ripp_b200.parallel import shard_bounds

    n = 6
    ps, qs = OS.g1_points("mg-a", n), OS.g2_points("mg-b", n)
    sc = OS.scalars("mg-s", n)
    lo, hi = shard_bounds(n, rank, world)
    # Miller partial of this rank's slice (no final exponentiation)
    f = E.F12_ONE
    for p, qq in zip(ps[lo:hi], qs[lo:hi]):
        f = E.f12_mul(f, E.miller_loop(p, qq))
    part = torch.from_numpy(C.gt_enc(f).astype(np.int64))
    gathered = [torch.zeros_like(part) for _ in range(world)]
    dist.all_gather(gathered, part)
    total = E.F12_ONE
    for g in gathered:  # rank order
        total = E.f12_mul(total, C.gt_dec(g.numpy().astype(np.uint32)))
    ok_pair = E.final_exponentiation(total) == E.multi_pairing(ps, qs)
    # MSM partial point of this rank's slice
    pt = E.msm(ps[lo:hi], sc[lo:hi], E.g1_add, E.g1_mul)
    part = torch.from_numpy(C.g1_enc(pt).astype(np.int64))
    gathered = [torch.zeros_like(part) for _ in range(world)]
    dist.all_gather(gathered, part)
    acc = None
    for g in gathered:
        acc = E.g1_add(acc, C.g1_dec(g.numpy().astype(np.uint32)))
    ok_msm = acc == E.msm(ps, sc, E.g1_add, E.g1_mul)
    q.put((rank, ok_pair, ok_msm, (lo, hi)))
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    from ripp_b200.parallel import shard_bounds

    for n in (0, 1, 5, 8, 4096, 65536 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_and_combine():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res)
