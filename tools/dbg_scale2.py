import sys
sys.path.insert(0, ".")
import numpy as np, random
from ripp_b200 import _lib, synth, codec as C
from oracle import bls12_381 as E
ctx = _lib.Context(0)
X = 0xd201000000010000
rnd = random.Random(5)
for n in (37, 48, 63, 64, 65, 96):
    for t in range(4):
        s = [(rnd.randrange(1 << 60) * X ** t) % E.R for _ in range(n)]
        ds = ctx.to_device(C.fr_vec_enc(s)); out = ctx.alloc(n * 192)
        ctx.g2_scale_dev(None, ds, n, out); ctx.sync()
        got = C.g2_vec_dec(out.download((n, 48)))
        exp = [E.g2_mul(E.G2_GEN, k) for k in s[:4]] 
        print("n=%d only digit %d: first4 ok=%s" % (n, t, [g == e for g, e in zip(got, exp)]))
    s = [rnd.randrange(E.R) for _ in range(n)]
    ds = ctx.to_device(C.fr_vec_enc(s)); out = ctx.alloc(n * 192)
    ctx.g2_scale_dev(None, ds, n, out); ctx.sync()
    got = C.g2_vec_dec(out.download((n, 48)))
    idx = [0, 1, n // 2, n - 1]
    print("n=%d random: ok=%s" % (n, [got[i] == E.g2_mul(E.G2_GEN, s[i]) for i in idx]))
