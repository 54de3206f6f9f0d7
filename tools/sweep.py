"""BASELINE configs[4]: scaling sweep n = 2^10 .. 2^max for the pairing and MSM inner products on this rank's GPU
(device resident, CUDA-event timed), plus SIPP prove at 2^10 (configs[0]).  Prints one JSON line per point.
    python tools/sweep.py [max_log_pairing] [max_log_msm]
"""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from ripp_b200 import _lib, codec, synth

ctx = _lib.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
max_p = int(sys.argv[1]) if len(sys.argv) > 1 else 20
max_m = int(sys.argv[2]) if len(sys.argv) > 2 else 22
imad_peak, _ = ctx.bench_imad(0, 4096)


def timed(fn, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for s, e in evs:
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    return min(s.elapsed_time(e) for s, e in evs)


out = torch.zeros(144, dtype=torch.int32, device="cuda")
a = synth.g1_points_dev(ctx, "cfg2-m", 1 << max_p)
b = synth.g2_points_dev(ctx, "cfg2-k", 1 << max_p)
for lg in range(10, max_p + 1, 2):
    n = 1 << lg
    ms = timed(lambda: ctx.pairing_ip_dev(a, b, n, out.data_ptr()))
    macs = n * 6700 * 300
    print(json.dumps({"op": "PairingInnerProduct", "log_n": lg, "ms": round(ms, 3), "pairs_per_s": round(n / ms * 1e3),
                      "frac_of_imad_wide_peak": round(macs / (ms * 1e-3) / imad_peak, 4)}))
a.free()
b.free()
bases = synth.g1_points_dev(ctx, "cfg3-a", 1 << max_m)
sc = ctx.to_device(synth.scalars_mont("cfg3-b", 1 << max_m))
pt = torch.zeros(24, dtype=torch.int32, device="cuda")
for lg in range(10, max_m + 1, 2):
    n = 1 << lg
    ms = timed(lambda: ctx.msm_g1_dev(bases, sc, n, pt.data_ptr()))
    print(json.dumps({"op": "MultiexponentiationInnerProduct<G1>", "log_n": lg, "ms": round(ms, 3),
                      "points_per_s": round(n / ms * 1e3)}))
bases.free()
sc.free()
# SIPP prove, n = 2^10 (BASELINE configs[0]) through the host-pointer entry point
n = 1 << 10
A = synth.g1_points_dev(ctx, "sipp-a", n).download((n, 24))
B = synth.g2_points_dev(ctx, "sipp-b", n).download((n, 48))
r = synth.scalars_mont("sipp-r", n)
t0 = time.perf_counter()
z = ctx.sipp_product_with_coeffs(A, B, r)
t_direct = time.perf_counter() - t0
ctx.sipp_prove(A, B, r, z)
t0 = time.perf_counter()
proof = ctx.sipp_prove(A, B, r, z)
t_prove = time.perf_counter() - t0
print(json.dumps({"op": "SIPP (BLS12-381, Blake2s)", "log_n": 10, "direct_s": round(t_direct, 4), "prove_s": round(t_prove, 4),
                  "proof_bytes": len(proof)}))
