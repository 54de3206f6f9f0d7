"""Repeated-run stability (dev tool): memory must not grow across aggregations / verifications / context churn, and two
contexts must be usable concurrently from two host threads."""
import sys, threading
sys.path.insert(0, ".")
import torch
from ripp_b200 import _lib, synth

ctx = _lib.Context(0)
n = 1 << 10
inst = synth.tipp_instance_dev(ctx, n)
def free_mb():
    return torch.cuda.mem_get_info()[0] / 2**20
proof = ctx.tipp_aggregate_dev(inst["srs_g1"], inst["srs_g2"], inst["a"], inst["b"], inst["c"], n)
assert ctx.tipp_verify_aggregate(inst["vsrs"], inst["vk"], inst["inputs"], proof)
base = free_mb()
for i in range(30):
    p2 = ctx.tipp_aggregate_dev(inst["srs_g1"], inst["srs_g2"], inst["a"], inst["b"], inst["c"], n)
    assert p2 == proof
    assert ctx.tipp_verify_aggregate(inst["vsrs"], inst["vk"], inst["inputs"], p2)
print("30 aggregations + verifications: free memory %.1f -> %.1f MiB" % (base, free_mb()))
assert base - free_mb() < 64
for i in range(20):  # context churn: create, use, destroy
    c2 = _lib.Context(0)
    d = c2.to_device(inst["a"].download((n, 24)))
    e = c2.to_device(inst["b"].download((n, 48)))
    out = c2.alloc(576)
    c2.pairing_ip_dev(d, e, n, out); c2.sync()
    d.free(); e.free(); out.free(); c2.close()
print("20 context create/destroy cycles: free memory %.1f MiB" % free_mb())
assert base - free_mb() < 64
# two contexts, two threads, same instance
res = {}
c3 = _lib.Context(0)
def run(name, c):
    res[name] = [c.tipp_aggregate_dev(inst["srs_g1"], inst["srs_g2"], inst["a"], inst["b"], inst["c"], n) for _ in range(5)]
t = threading.Thread(target=run, args=("b", c3)); t.start(); run("a", ctx); t.join()
assert all(p == proof for p in res["a"] + res["b"])
print("two contexts x two threads: 10 concurrent aggregations identical")
print("stress ok")
