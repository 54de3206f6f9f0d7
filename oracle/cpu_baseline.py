"""ORACLE side (test / bench infrastructure): CPU timing of the reference path for bench.py's
`cpu_baseline` object and `--impl reference` arm.

Until the compiled C++ restatement (oracle/cpu) is built this times the pure-Python big-int
restatement on one core; kind = "port".
"""
import os
import time

from . import bls12_381 as E
from . import synth

SAMPLE_PAIRS = 128


def _compiled():
    try:
        from .cpu import binding

        return binding.load()
    except Exception:
        return None


def pairing_pairs_per_s(sample_pairs=None):
    lib = _compiled()
    if lib is not None:
        return lib.pairing_pairs_per_s(sample_pairs)
    n = sample_pairs or SAMPLE_PAIRS
    ps, qs = synth.g1_points("cfg2-m", n), synth.g2_points("cfg2-k", n)
    t0 = time.perf_counter()
    E.multi_pairing(ps, qs)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "pairs/s", "cores": 1, "kind": "port", "seconds": dt,
            "sample": "%d pairs of the 2^16-pair workload, pure-Python big-int restatement, 1 core" % n}
