"""Quick timing of ripp_tipp_aggregate_dev (dev tool; bench.py is the contract)."""
import sys, time
sys.path.insert(0, ".")
from ripp_b200 import _lib, synth
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 12
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = _lib.Context(0)
n = 1 << logn
t0 = time.time(); inst = synth.tipp_instance_dev(ctx, n); print("setup s", time.time() - t0)
ts = []
for i in range(reps):
    l0 = ctx.launches; t0 = time.time()
    proof = ctx.tipp_aggregate_dev(inst["srs_g1"], inst["srs_g2"], inst["a"], inst["b"], inst["c"], n)
    ts.append(1e3 * (time.time() - t0))
    print("aggregate n=%d: %.1f ms, %d launches, %d proof bytes" % (n, ts[-1], ctx.launches - l0, len(proof)))
if reps > 2:
    print("mean of last %d: %.1f ms (min %.1f)" % (reps - 1, sum(ts[1:]) / (reps - 1), min(ts[1:])))
ctx.set_timing(True); ctx.timing()
proof = ctx.tipp_aggregate_dev(inst["srs_g1"], inst["srs_g2"], inst["a"], inst["b"], inst["c"], n)
tm = ctx.timing()
print({k: (round(v[0], 2), v[1], round(v[0] / max(v[1], 1), 2)) for k, v in tm.items()})
