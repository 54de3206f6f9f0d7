"""GIPA prove of ONE global instance partitioned cyclically over the ranks (dev tool; bench.py reports the same
as sub_metrics.gipa_multiexp_prove_s).   torchrun --nproc-per-node N tools/sharded_gipa.py [kind] [log_n] [reps]"""
import hashlib
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist

from ripp_b200 import _lib, synth
from ripp_b200.parallel import Comm, ShardedGIPA, _KIND_TYPES

kind = int(sys.argv[1]) if len(sys.argv) > 1 else 1
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 18
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
os.environ["NCCL_DEBUG_FILE"] = "/tmp/ripp_b200_nccl_%h_%p.log"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = _lib.Context(lr)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ctx.set_stream(st.cuda_stream)
n = 1 << logn
nl = n // world


def share(tag, t):
    if t is None:
        return None
    sc = np.ascontiguousarray(synth.scalars_mont(tag, n)[rank::world])
    if t == "Fr":
        return torch.from_numpy(sc.view(np.int32)).cuda()
    d = ctx.to_device(sc)
    out = torch.empty((nl, 24 if t == "G1" else 48), dtype=torch.int32, device="cuda")
    (ctx.g1_scale_dev if t == "G1" else ctx.g2_scale_dev)(None, d, nl, out.data_ptr())
    ctx.sync()
    return out


vecs = [share(tag, t) for tag, t in zip(("cfg3-a", "cfg3-b", "cfg3-v", "cfg3-w"), _KIND_TYPES[kind])]
sg = ShardedGIPA(kind, ctx, Comm())
proof = sg.prove_with_aux_dev(*vecs)[0]
for _ in range(reps):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    proof = sg.prove_with_aux_dev(*vecs)[0]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print("kind %d n=2^%d world=%d: %.1f ms  proof %d B blake2b %s" % (kind, logn, world, 1e3 * dt, len(proof),
                                                                        hashlib.blake2b(proof, digest_size=16).hexdigest()))
if os.environ.get("RIPP_TRACE_ROUNDS") and rank == 0:
    print(getattr(sg, "round_ms", None))
if world > 1:
    dist.destroy_process_group()
