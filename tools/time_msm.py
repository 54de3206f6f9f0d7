"""Quick timing of ripp_msm_g1_dev (dev tool)."""
import sys, time
sys.path.insert(0, ".")
from ripp_b200 import _lib, synth
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 18
ctx = _lib.Context(0)
n = 1 << logn
bases = synth.g1_points_dev(ctx, "cfg3-a", n)
sc = ctx.to_device(synth.scalars_mont("cfg3-b", n))
out = ctx.alloc(96)
for i in range(3):
    ctx.sync(); t0 = time.time(); ctx.msm_g1_dev(bases, sc, n, out); ctx.sync()
    print("msm g1 n=2^%d: %.2f ms" % (logn, 1e3 * (time.time() - t0)))
