"""Per-round component timing (dev tool): six pairing products of length n' in one batch (Miller + product tree +
final exponentiation = what one GIPA round of the pairing instantiation launches), for n' = 1 .. 4096.
RIPP_B200_L18_WARPS=0 forces the six-lane throughput shape everywhere (A/B against the eighteen-lane shape)."""
import os, sys
sys.path.insert(0, ".")
import ctypes
import numpy as np, torch
from ripp_b200 import _lib, synth
ctx = _lib.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
N = 8192
g1 = synth.g1_points_dev(ctx, "tr-a", N); g2 = synth.g2_points_dev(ctx, "tr-b", N)
out = ctx.alloc(8 * 576)
L = _lib.lib()
for lg in range(0, 13):
    n = 1 << lg
    nseg = 6
    a1 = (ctypes.c_void_p * nseg)(*[g1.ptr + 96 * ((i * 577) % (N - n)) for i in range(nseg)])
    a2 = (ctypes.c_void_p * nseg)(*[g2.ptr + 192 * ((i * 911) % (N - n)) for i in range(nseg)])
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.ripp_pairing_ip_batch_dev(ctx.handle, nseg, a1, a2, ctypes.c_size_t(n), ctypes.c_void_p(out.ptr)))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ctx.set_timing(True); ctx.timing()
    _lib.check(L.ripp_pairing_ip_batch_dev(ctx.handle, nseg, a1, a2, ctypes.c_size_t(n), ctypes.c_void_p(out.ptr)))
    tm = ctx.timing(); ctx.set_timing(False)
    print("6 products x n'=%4d  L18_WARPS=%s: total %.3f ms (miller+tree %.3f, final exp %.3f)" % (
        n, os.environ.get("RIPP_B200_L18_WARPS", "default"), min(ts[1:]), tm["miller"][0], tm["final_exp"][0]))
