"""ORACLE side (test / bench infrastructure): CPU timing of the reference path for bench.py's
`cpu_baseline` object and its `--impl reference` arm.

The reference (arkworks + rayon) cannot be built in this environment (no Rust toolchain, crates not
vendored), so the timed CPU implementation is the compiled restatement oracle/cpu (kind = "port"):
same algorithms and the same parallel structure (one Miller-loop chunk per thread, windows of the
MSM in parallel, element-wise maps in parallel), OpenMP on all host cores.
"""
import os
import time

from . import bls12_381 as E
from . import protocols as O
from . import synth


class TippWorkload:
    """aggregate_proofs (groth16_aggregation.rs:77-160) on the synthetic 2^k-proof instance of SURVEY.md §8d."""

    def __init__(self, n, seed=0, threads=None):
        from .cpu import binding

        self.n = n
        self.be = binding.CppBackend(threads)
        be = self.be
        self.srs = O.tipa_setup(n, synth.scalar("srs-alpha", 0, seed), synth.scalar("srs-beta", 0, seed), be)
        _, sc, _ = synth.groth16_instance_scalars(n, seed=seed)
        g1 = be.vec_g1([E.G1_GEN] * n)
        g2 = be.vec_g2([E.G2_GEN] * n)
        self.a = be.mul_vec_g1(g1, [p[0] for p in sc])
        self.b = be.mul_vec_g2(g2, [p[1] for p in sc])
        self.c = be.mul_vec_g1(g1, [p[2] for p in sc])
        self.proof = None

    def run(self):
        proofs = _Columns(self.a, self.b, self.c)
        t0 = time.perf_counter()
        self.proof = O.aggregate_proofs(self.srs, proofs, be=self.be)
        return time.perf_counter() - t0

    def info(self):
        return {"cores": self.be.cpu.threads, "kind": "port",
                "sample": "one full aggregation of %d proofs (not sampled), compiled CPU restatement with OpenMP on %d threads; "
                          "arkworks itself cannot be built here" % (self.n, self.be.cpu.threads)}


class _Columns:
    """Sequence of (A, B, C) triples backed by three packed vectors (avoids per-element decoding)."""

    def __init__(self, a, b, c):
        self.cols = (a, b, c)

    def __len__(self):
        return len(self.cols[0])

    def column(self, j):
        return self.cols[j]


def pairing_pairs_per_s(sample_pairs=2048, threads=None):
    from .cpu import binding

    be = binding.CppBackend(threads)
    n = sample_pairs
    g1 = be.mul_vec_g1(be.vec_g1([E.G1_GEN] * n), synth.scalars("cfg2-m", n))
    g2 = be.mul_vec_g2(be.vec_g2([E.G2_GEN] * n), synth.scalars("cfg2-k", n))
    t0 = time.perf_counter()
    be.pairing_product(g1, g2)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "pairs/s", "cores": be.cpu.threads, "kind": "port", "seconds": dt,
            "sample": "%d pairs" % n}


def msm_points_per_s(sample_points=1 << 14, threads=None):
    """G1 MSM (arkworks' bucket method as restated in oracle/cpu) on `sample_points` synthetic points."""
    from .cpu import binding

    be = binding.CppBackend(threads)
    n = sample_points
    sc = synth.scalars("cfg3-b", n)
    g1 = be.mul_vec_g1(be.vec_g1([E.G1_GEN] * n), synth.scalars("cfg3-a", n))
    fr = be.vec_fr(sc)
    t0 = time.perf_counter()
    be.msm_g1(g1, fr)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "points/s", "cores": be.cpu.threads, "kind": "port", "seconds": dt,
            "sample": "%d points" % n}
