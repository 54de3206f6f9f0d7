"""The C-ABI library loads and exports every symbol include/ripp_b200.h declares (no compute calls:
there is no GPU here), and refuses to run without a device instead of falling back to the CPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from ripp_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        from ripp_b200 import build

        build.build_lib()
    return _lib


def test_exports_every_declared_symbol():
    L = _lib().lib()
    hdr = open(os.path.join(ROOT, "include", "ripp_b200.h")).read()
    names = sorted(set(re.findall(r"\b(ripp_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), n


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = _lib()
    with pytest.raises(m.RippError) as e:
        m.Context(0)
    assert e.value.status == m.RIPP_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ripp_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"(from|import)\s+\.*oracle|#include[^\n]*oracle", src), os.path.join(dp, f)
