"""Quick timing of ripp_miller_partial_dev / pairing_ip_dev (dev tool)."""
import sys, time
sys.path.insert(0, ".")
from ripp_b200 import _lib, synth
ctx = _lib.Context(0)
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = 1 << logn
a = synth.g1_points_dev(ctx, "cfg2-m", n); b = synth.g2_points_dev(ctx, "cfg2-k", n)
out = ctx.alloc(576)
for i in range(3):
    ctx.sync(); t0 = time.time(); ctx.miller_partial_dev(a, b, n, out); ctx.sync()
    print("miller partial n=2^%d: %.2f ms" % (logn, 1e3 * (time.time() - t0)))
