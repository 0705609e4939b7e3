"""The product library loads, exports every symbol include/relem.h declares, and refuses to work without a GPU
(no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

import rnaelem_b200 as rb
from rnaelem_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "relem.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(relem_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(binding.SYMBOLS)


def test_library_exports_every_declared_symbol():
    path = rb.lib_path()
    assert os.path.exists(path), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert b"sm_100a" in ctypes.c_char_p(ctypes.cast(lib.relem_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()).value


def test_library_contains_sm100a_code():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", rb.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rb.RelemError) as e:
        rb.Context(0)
    assert "no CPU path" in str(e.value) or "no CUDA device" in str(e.value)


def test_assigned_range_is_the_reference_sharding():
    """ArrayJobManager::assigned_range (arrayjob_manager.hpp:141-149)."""
    lib = rb.load_library()
    for total, n in [(10, 3), (7, 7), (5, 8), (100000, 8), (0, 4)]:
        covered = []
        for k in range(n):
            a, b = ctypes.c_int64(), ctypes.c_int64()
            lib.relem_assigned_range(total, n, k, ctypes.byref(a), ctypes.byref(b))
            covered += list(range(a.value, b.value))
            assert 0 <= b.value - a.value <= total // n + 1
        assert covered == list(range(total))
