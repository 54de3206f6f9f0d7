"""SIPP<Bls12_377, Blake2s> timing on the GPU (dev tool): direct product, prove, verify at n = 2^logn (scaling-ipp.rs:72-82)."""
import sys, time, random
sys.path.insert(0, ".")
from ripp_b200 import _lib
from ripp_b200 import sipp_377 as G
from oracle import sipp_377 as S
from oracle import bls12_377 as E
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n = 1 << logn
ctx = _lib.Context(0)
rnd = random.Random(1)
# subgroup points from a few oracle points combined on the GPU would need more API; sizes here stay small enough for the oracle
base_a, base_b = S.points("t377-a", 8, 1), S.points("t377-b", 8, 2)
a = [base_a[i % 8] for i in range(n)]; b = [base_b[(3 * i) % 8] for i in range(n)]
r = [rnd.randrange(E.R) for _ in range(n)]
for rep in range(2):
    t0 = time.time(); z = G.product_of_pairings_with_coeffs(a, b, r, ctx); t1 = time.time()
    p = G.SIPP377.prove(a, b, r, z, ctx); t2 = time.time()
    ok = G.SIPP377.verify(a, b, r, z, p, ctx); t3 = time.time()
    print("SIPP BLS12-377 n=2^%d: direct %.1f ms, prove %.1f ms, verify %.1f ms, accepted=%s" % (logn, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), ok))
