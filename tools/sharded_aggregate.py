"""aggregate_proofs of ONE batch of 2^log_n proofs partitioned over the ranks (dev / evidence tool).
torchrun --nproc-per-node N tools/sharded_aggregate.py [log_n] [reps]"""
import hashlib
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist

from ripp_b200 import _lib, synth
from ripp_b200.parallel import Comm, cyclic_share, shard_bounds, sharded_aggregate_proofs

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 14
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
os.environ["NCCL_DEBUG_FILE"] = "/tmp/ripp_b200_nccl_%h_%p.log"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = _lib.Context(lr)
n = 1 << logn
inst = synth.tipp_instance_dev(ctx, n)  # every rank generates the same global instance, then keeps its shares
h = {"s1": inst["srs_g1"].download((2 * n - 1, 24)), "s2": inst["srs_g2"].download((2 * n - 1, 48)),
     "a": inst["a"].download((n, 24)), "b": inst["b"].download((n, 48)), "c": inst["c"].download((n, 24))}
if world == 1:  # the resident single-GPU prover on the same instance, for reference
    for _ in range(2):
        ctx.sync(); t0 = time.perf_counter()
        ref = ctx.tipp_aggregate_dev(inst["srs_g1"], inst["srs_g2"], inst["a"], inst["b"], inst["c"], n)
        t_ref = time.perf_counter() - t0
    print("resident single-GPU prover n=2^%d: %.1f ms  blake2b %s" % (logn, 1e3 * t_ref, hashlib.blake2b(ref, digest_size=16).hexdigest()))
for k in ("srs_g1", "srs_g2", "a", "b", "c"):
    inst[k].free()
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
cu = lambda x: torch.from_numpy(np.ascontiguousarray(x).view(np.int32)).cuda()
sh = lambda v: cu(cyclic_share(v, rank, world))
lo, hi = shard_bounds(2 * n - 1, rank, world)
args = (cu(h["s1"][lo:hi]), cu(h["s2"][lo:hi]), lo, n, sh(h["s2"][::2][:n]), sh(h["s1"][::2][:n]), sh(h["a"]), sh(h["b"]), sh(h["c"]))
comm = Comm()
proof = sharded_aggregate_proofs(ctx, comm, *args)
for _ in range(reps):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    proof = sharded_aggregate_proofs(ctx, comm, *args)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print("sharded aggregate n=2^%d world=%d: %.1f ms  proof %d B blake2b %s" % (logn, world, 1e3 * dt, len(proof), hashlib.blake2b(proof, digest_size=16).hexdigest()))
if rank == 0:
    print("verify on GPU:", ctx.tipp_verify_aggregate(inst["vsrs"], inst["vk"], inst["inputs"], proof))
if world > 1:
    dist.destroy_process_group()
