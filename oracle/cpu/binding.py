"""ORACLE side: build + ctypes binding of the compiled CPU restatement (oracle/cpu/ripp_cpu.cpp) and a
`Backend` for oracle/protocols.py that keeps vectors as packed limb arrays, so the protocol layer can
run at BASELINE sizes and be timed as the CPU baseline (bench.py cpu_baseline / --impl reference)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ripp_cpu.cpp")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libripp_cpu.so")

_P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_RQ, _RR = 1 << 384, 1 << 256
_RQI, _RRI = pow(_RQ, -1, _P), pow(_RR, -1, _R)


def build(force=False):
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cpp", ".hpp"))]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    # -march=x86-64-v3 (not native): the .so built here must run on the GPU box's host CPU
    cmd = ["g++", "-O3", "-std=c++17", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC]
    subprocess.run(cmd, check=True)
    return LIB


class CpuLib:
    def __init__(self):
        build()
        self.lib = ctypes.CDLL(LIB)
        self.threads = self.lib.cpu_init()

    def set_threads(self, t):
        self.lib.cpu_set_threads(int(t))
        self.threads = int(t)


_cpu = None


def load():
    global _cpu
    if _cpu is None:
        _cpu = CpuLib()
    return _cpu


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _w(v, nbytes):
    return np.frombuffer(int(v).to_bytes(nbytes, "little"), dtype=np.uint64)


def _i(a):
    return int.from_bytes(np.ascontiguousarray(a).tobytes(), "little")


# ---- vector containers: (n, words) uint64 arrays of Montgomery limbs ------------------------------
class Vec:
    """Packed vector that slices like a list; indexing an element decodes it to the tuple form the
    pure-Python oracle uses (only done for the handful of values that get serialised)."""

    kind = None
    words = 0

    def __init__(self, arr):
        self.a = arr

    def __len__(self):
        return self.a.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return type(self)(self.a[i])
        return self.dec(self.a[i])

    def __iter__(self):
        return (self.dec(r) for r in self.a)


class G1Vec(Vec):
    words = 12

    @staticmethod
    def dec(r):
        if not r.any():
            return None
        return (_i(r[:6]) * _RQI % _P, _i(r[6:]) * _RQI % _P)

    @staticmethod
    def enc(pts):
        out = np.zeros((len(pts), 12), dtype=np.uint64)
        for k, p in enumerate(pts):
            if p is not None:
                out[k, :6] = _w(p[0] * _RQ % _P, 48)
                out[k, 6:] = _w(p[1] * _RQ % _P, 48)
        return G1Vec(out)


class G2Vec(Vec):
    words = 24

    @staticmethod
    def dec(r):
        if not r.any():
            return None
        v = [_i(r[6 * j : 6 * j + 6]) * _RQI % _P for j in range(4)]
        return ((v[0], v[1]), (v[2], v[3]))

    @staticmethod
    def enc(pts):
        out = np.zeros((len(pts), 24), dtype=np.uint64)
        for k, p in enumerate(pts):
            if p is not None:
                for j, v in enumerate((p[0][0], p[0][1], p[1][0], p[1][1])):
                    out[k, 6 * j : 6 * j + 6] = _w(v * _RQ % _P, 48)
        return G2Vec(out)


class FrVec(Vec):
    words = 4

    @staticmethod
    def dec(r):
        return _i(r) * _RRI % _R

    @staticmethod
    def enc(vals):
        out = np.zeros((len(vals), 4), dtype=np.uint64)
        for k, v in enumerate(vals):
            out[k] = _w(v % _R * _RR % _R, 32)
        return FrVec(out)


def _gt_dec(a):
    order = (0, 2, 4, 1, 3, 5)
    out = [None] * 6
    for slot, k in enumerate(order):
        out[k] = (_i(a[12 * slot : 12 * slot + 6]) * _RQI % _P, _i(a[12 * slot + 6 : 12 * slot + 12]) * _RQI % _P)
    return tuple(out)


def _as(vec, cls):
    if isinstance(vec, cls):
        return np.ascontiguousarray(vec.a)
    return np.ascontiguousarray(cls.enc(list(vec)).a)


class CppBackend:
    """Heavy ops of oracle/protocols.py on the compiled restatement (OpenMP)."""

    def __init__(self, threads=None):
        self.cpu = load()
        if threads:
            self.cpu.set_threads(threads)

    # containers
    def vec_g1(self, pts):
        return pts if isinstance(pts, G1Vec) else G1Vec.enc(list(pts))

    def vec_g2(self, pts):
        return pts if isinstance(pts, G2Vec) else G2Vec.enc(list(pts))

    def vec_fr(self, vals):
        return vals if isinstance(vals, FrVec) else FrVec.enc(list(vals))

    def pairing_product(self, g1s, g2s):
        a, b = _as(g1s, G1Vec), _as(g2s, G2Vec)
        out = np.zeros(72, dtype=np.uint64)
        self.cpu.lib.cpu_pairing_product(_ptr(a), _ptr(b), ctypes.c_size_t(len(a)), _ptr(out))
        return _gt_dec(out)

    def msm_g1(self, pts, scalars):
        a, s = _as(pts, G1Vec), _as(scalars, FrVec)
        out = np.zeros(12, dtype=np.uint64)
        self.cpu.lib.cpu_msm_g1(_ptr(a), _ptr(s), ctypes.c_size_t(len(a)), _ptr(out))
        return G1Vec.dec(out)

    def msm_g2(self, pts, scalars):
        a, s = _as(pts, G2Vec), _as(scalars, FrVec)
        out = np.zeros(24, dtype=np.uint64)
        self.cpu.lib.cpu_msm_g2(_ptr(a), _ptr(s), ctypes.c_size_t(len(a)), _ptr(out))
        return G2Vec.dec(out)

    def mul_vec_g1(self, pts, scalars):
        a, s = _as(pts, G1Vec), _as(scalars, FrVec)
        out = np.zeros_like(a)
        self.cpu.lib.cpu_scale_g1(_ptr(a), _ptr(s), ctypes.c_size_t(len(a)), _ptr(out))
        return G1Vec(out)

    def mul_vec_g2(self, pts, scalars):
        a, s = _as(pts, G2Vec), _as(scalars, FrVec)
        out = np.zeros_like(a)
        self.cpu.lib.cpu_scale_g2(_ptr(a), _ptr(s), ctypes.c_size_t(len(a)), _ptr(out))
        return G2Vec(out)

    def fold_g1(self, hi, lo, c):
        h, l = _as(hi, G1Vec), _as(lo, G1Vec)
        cc = np.ascontiguousarray(FrVec.enc([c]).a)
        out = np.zeros_like(h)
        self.cpu.lib.cpu_fold_g1(_ptr(h), _ptr(l), _ptr(cc), ctypes.c_size_t(len(h)), _ptr(out))
        return G1Vec(out)

    def fold_g2(self, hi, lo, c):
        h, l = _as(hi, G2Vec), _as(lo, G2Vec)
        cc = np.ascontiguousarray(FrVec.enc([c]).a)
        out = np.zeros_like(h)
        self.cpu.lib.cpu_fold_g2(_ptr(h), _ptr(l), _ptr(cc), ctypes.c_size_t(len(h)), _ptr(out))
        return G2Vec(out)

    def fold_fr(self, hi, lo, c):
        h, l = _as(hi, FrVec), _as(lo, FrVec)
        cc = np.ascontiguousarray(FrVec.enc([c]).a)
        out = np.zeros_like(h)
        self.cpu.lib.cpu_fold_fr(_ptr(h), _ptr(l), _ptr(cc), ctypes.c_size_t(len(h)), _ptr(out))
        return FrVec(out)
