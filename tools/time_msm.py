"""Quick timing of ripp_msm_g1_dev (dev tool)."""
import sys, time
sys.path.insert(0, ".")
from ripp_b200 import _lib, synth
ctx = _lib.Context(0)
for logn in [int(x) for x in sys.argv[1:]] or [18]:
    n = 1 << logn
    bases = synth.g1_points_dev(ctx, "cfg3-a", n)
    sc = ctx.to_device(synth.scalars_mont("cfg3-b", n))
    out = ctx.alloc(96)
    for i in range(3):
        ctx.sync(); t0 = time.time(); ctx.msm_g1_dev(bases, sc, n, out); ctx.sync()
        dt = time.time() - t0
    ctx.set_timing(True); ctx.timing(); l0 = ctx.launches
    ctx.msm_g1_dev(bases, sc, n, out); tm = ctx.timing(); ctx.set_timing(False)
    print("msm g1 n=2^%d: %.2f ms wall, %.2f ms device, %d launches" % (logn, 1e3 * dt, sum(tm[k][0] for k in ("msm", "msm_sort", "msm_reduce")), ctx.launches - l0))
