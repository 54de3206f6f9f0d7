import sys
sys.path.insert(0, ".")
from ripp_b200 import _lib
ctx = _lib.Context(0)
for k, name in [(0, "IMAD.WIDE plain"), (1, "IMAD 32"), (2, "carry chain in+out"), (3, "carry out only + addc count"), (4, "carry in only + add.cc producer")]:
    macs, ms = ctx.bench_imad(k, 4096)
    print("kind %d %-34s %.2f TMAC/s (%.3f ms)" % (k, name, macs / 1e12, ms))
