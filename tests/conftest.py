import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "librelem_emu.so")
CSRC = os.path.join(ROOT, "rnaelem_b200", "csrc")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


@pytest.fixture(scope="session")
def emu_lib():
    """Single-threaded host emulation of the kernel source (debug aid for a GPU-less container; never part of
    the product): the same .cu/.cuh files compiled by g++ with -DRELEM_HOST_EMU."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if not _newer(EMU_LIB, srcs):
        subprocess.check_call(["make", "-C", EMU_DIR, "-s", "librelem_emu.so"])
    return EMU_LIB


@pytest.fixture(scope="session")
def emu_renorm_lib(emu_lib):
    """the emulation built with tiny exterior-row rescaling thresholds (see tests/emu/Makefile)"""
    p = os.path.join(EMU_DIR, "librelem_emu_renorm.so")
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if not _newer(p, srcs):
        subprocess.check_call(["make", "-C", EMU_DIR, "-s", "librelem_emu_renorm.so"])
    return p


@pytest.fixture(scope="session")
def gpu_lib():
    import rnaelem_b200 as rb
    p = rb.lib_path()
    assert os.path.exists(p), "librelem.so missing: run __graft_entry__.build()"
    return p
