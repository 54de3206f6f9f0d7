import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


class HostSim:
    """ctypes view of tests/hostsim/_hostsim.so: the device headers compiled for the host."""

    def __init__(self, lib):
        self.lib = lib
        lib.hs_mul_count.restype = ctypes.c_uint64

    def call(self, fn, *args, out):
        o = np.zeros(out, dtype=np.uint32)
        cargs, keep = [], []
        for a in args:
            if isinstance(a, np.ndarray):
                a = np.ascontiguousarray(a, dtype=np.uint32)
                keep.append(a)
                cargs.append(ctypes.c_void_p(a.ctypes.data))
            else:
                cargs.append(a)
        getattr(self.lib, fn)(*cargs, ctypes.c_void_p(o.ctypes.data))
        return o


def _build_hostsim(flags):
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    import hashlib

    tag = "_" + hashlib.md5(" ".join(flags).encode()).hexdigest()[:8] if flags else ""  # stable across processes
    so = os.path.join(ROOT, "tests", "hostsim", "_hostsim%s.so" % tag)
    csrc = os.path.join(ROOT, "ripp_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC"] + flags + ["-o", so, src], check=True)
    return HostSim(ctypes.CDLL(so))


@pytest.fixture(scope="session")
def hostsim():
    # RIPP_HOSTSIM_FLAGS: extra -D flags for the default harness (development runs)
    return _build_hostsim(os.environ.get("RIPP_HOSTSIM_FLAGS", "").split())


@pytest.fixture(scope="session")
def ctx():
    from ripp_b200 import _lib

    c = _lib.Context(0)
    yield c
    c.close()
