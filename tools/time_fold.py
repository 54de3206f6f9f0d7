"""Fold kernel timing (dev tool): out = hi * c + lo for G1 / G2, shared scalar of 128 or 255 bits.
RIPP_B200_FOLD=endo|w3 selects the kernel."""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from ripp_b200 import _lib, codec, synth
ctx = _lib.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
N = 8192
g1 = synth.g1_points_dev(ctx, "tf-a", N); g2 = synth.g2_points_dev(ctx, "tf-b", N)
o1 = ctx.alloc(N * 96); o2 = ctx.alloc(N * 192)
for bits in (128, 255):
    c = codec.fr_enc(synth.scalar("tf-c", bits) % (1 << bits) | (1 << (bits - 1)) if bits < 255 else synth.scalar("tf-c", 0) | (1 << 247)).copy()
    for n in (32, 256, 2048, 4096):
        for name, fn, src, dst, w in (("G1", ctx.g1_fold_dev, g1, o1, 96), ("G2", ctx.g2_fold_dev, g2, o2, 192)):
            ts = []
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(src.ptr + n * w, src.ptr, c, n, dst); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print("%s mode=%s bits=%d n=%d: %.3f ms" % (name, os.environ.get("RIPP_B200_FOLD", "default"), bits, n, min(ts[1:])))
