"""One rank of the multi-GPU parity run (launched by tests/test_gpu_multi.py through torch.distributed.run, one process
per GPU).  The library's own NCCL communicator carries the data path; gloo only broadcasts the NCCL unique id.
Every sharded entry point is compared BYTE FOR BYTE with the single-GPU entry point on the whole instance."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from ripp_b200 import _lib, synth
    from ripp_b200.parallel import init_library_comm, shard_bounds

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ctx = _lib.Context(local)
    r, w = init_library_comm(ctx)
    assert (r, w) == (rank, world) and ctx.comm_info() == (rank, world)
    checks = []

    def same(name, got, want):
        ok = bool(np.array_equal(np.asarray(got), np.asarray(want))) if not isinstance(got, bytes) else got == want
        checks.append((name, ok))
        assert ok, "rank %d: %s differs from the single-GPU result" % (rank, name)

    # ---- leaf inner products over contiguous slices (sizes that are not multiples of the world size included) ----
    for n in (1000, 4096):
        g1, g2 = synth.g1_points_dev(ctx, "mg-a", n), synth.g2_points_dev(ctx, "mg-b", n)
        sc = ctx.to_device(synth.scalars_mont("mg-s", n))
        lo, hi = shard_bounds(n, rank, world)
        whole, part = ctx.alloc(576), ctx.alloc(576)
        ctx.pairing_ip_dev(g1, g2, n, whole)
        ctx.pairing_ip_sharded_dev(g1.ptr + 96 * lo, g2.ptr + 192 * lo, hi - lo, part)
        ctx.sync()
        same("pairing_ip n=%d" % n, part.download(144), whole.download(144))
        ctx.msm_g1_dev(g1, sc, n, whole)
        ctx.msm_sharded_dev(1, g1.ptr + 96 * lo, sc.ptr + 32 * lo, hi - lo, part)
        ctx.sync()
        same("msm_g1 n=%d" % n, part.download(24), whole.download(24))
        ctx.msm_g2_dev(g2, sc, n, whole)
        ctx.msm_sharded_dev(2, g2.ptr + 192 * lo, sc.ptr + 32 * lo, hi - lo, part)
        ctx.sync()
        same("msm_g2 n=%d" % n, part.download(48), whole.download(48))

    # ---- GIPA over the cyclic partition: all-gathered rounds, then the resident tail ----
    def cyc(buf, n, words):
        h = buf.download((n, words))
        return ctx.to_device(np.ascontiguousarray(h[rank::world]))

    n = 256
    a, b = synth.g1_points_dev(ctx, "mg-ga", n), synth.g2_points_dev(ctx, "mg-gb", n)
    v, wv = synth.g2_points_dev(ctx, "mg-gv", n), synth.g1_points_dev(ctx, "mg-gw", n)
    s = ctx.to_device(synth.scalars_mont("mg-gs", n))
    for kind, vecs in ((0, ((a, 24), (b, 48), (v, 48), (wv, 24))), (1, ((a, 24), (s, 8), (v, 48), (wv, 24))),
                       (2, ((a, 24), (s, 8), (v, 48), None))):
        full = [x[0] if x else None for x in vecs]
        want = ctx.gipa_prove_dev(kind, full[0], full[1], full[2], full[3], n)
        sh = [cyc(x[0], n, x[1]) if x else None for x in vecs]
        for tail in (16, 2 * world):
            got = ctx.gipa_prove_sharded_dev(kind, sh[0], sh[1], sh[2], sh[3], n // world, world, tail_len=tail)
            same("gipa kind %d tail %d proof" % (kind, tail), got[0], want[0])
            same("gipa kind %d tail %d transcript" % (kind, tail), got[1], want[1])
            same("gipa kind %d tail %d ck_base" % (kind, tail), got[2], want[2])

    # ---- aggregate_proofs of ONE batch partitioned over the ranks ----
    for n, tail in ((64, 16), (256, 32)):
        inst = synth.tipp_instance_dev(ctx, n)
        want = ctx.tipp_aggregate_dev(inst["srs_g1"], inst["srs_g2"], inst["a"], inst["b"], inst["c"], n)
        sa, sb, sc_ = cyc(inst["a"], n, 24), cyc(inst["b"], n, 48), cyc(inst["c"], n, 24)
        got = ctx.tipp_aggregate_sharded_dev(inst["srs_g1"], inst["srs_g2"], sa, sb, sc_, n, tail_len=tail)
        same("aggregate n=%d" % n, got, want)
        assert ctx.tipp_verify_aggregate(inst["vsrs"], inst["vk"], inst["inputs"], got)

    dist.barrier()
    if rank == 0:
        print("MGPU_OK world=%d checks=%d" % (world, len(checks)))
    ctx.comm_destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
