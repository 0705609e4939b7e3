"""The drop-in claim proven on the reference itself: the reference's OWN main.cpp -- option parser, FastqBatchReader,
Adam and L-BFGS-B (optimizer.hpp), regulariser, model / interim / raw writers, all unmodified -- with only the two
thread fan-outs of the hot path (motif_trainer.hpp:617-621, motif_scanner.hpp:943-946) replaced by calls into the C ABI
of include/relem.h (integration/relem_host*.hpp, applied by integration/*.sed in oracle/Makefile target `gpu`), must
reproduce the golden outputs of the unmodified reference binary on every golden command line, including the
--no-shuffle / L-BFGS-B runs that the shipped rnaelem_b200/RNAelem refuses.

CPU half: the patched reference linked with the single-thread emulation of the kernel source (debug aid, tests/emu);
only where /root/reference exists (the build container).  GPU half: oracle/_ref/RNAelem_gpu (prebuilt in the build
container, travels with the snapshot) linked with the product librelem.so.
"""
import os
import subprocess

import pytest

import clilib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")
GPU_BIN = os.path.join(ORACLE, "_ref", "RNAelem_gpu")
EMU_BIN = os.path.join(ORACLE, "_ref", "RNAelem_ref_emu")
HAVE_REF = os.path.exists("/root/reference/RNAelem/main.cpp")
CASES = ["synth_adam", "trna_softmax", "ragged_train", "ragged_scan", "norss_adam", "ragged_likratio", "synth_scan",
         "synth_lbfgsb", "ragged_lbfgsb"]


@pytest.fixture(scope="session")
def dropin_emu(emu_lib):
    if not HAVE_REF:
        pytest.skip("reference sources not present: the patched reference cannot be built here")
    subprocess.check_call(["make", "-s", "-C", ORACLE, "gpu", "RELEM_LIBDIR=../tests/emu", "RELEM_LIB=relem_emu",
                           "GPU_BIN=_ref/RNAelem_ref_emu", "RPATH=$$ORIGIN/../../tests/emu"])
    return EMU_BIN


@pytest.mark.parametrize("name", CASES)
def test_patched_reference_emulated(name, dropin_emu, tmp_path):
    clilib.check_case(dropin_emu, name, str(tmp_path))


def test_sed_recipe_touches_only_the_two_fan_outs():
    """the edit scripts must change exactly the lines INTEGRATION.md names"""
    if not HAVE_REF:
        pytest.skip("reference sources not present")
    for script, header, gone, added in (
            ("trainer.sed", "motif_trainer.hpp", "ClassThread<RNAelemTrainDP>", "relem_host::estep("),
            ("scanner.sed", "motif_scanner.hpp", "ClassThread<RNAelemScanDP> ct(", "relem_host::scan(")):
        src = open(os.path.join("/root/reference/RNAelem", header)).read().split("\n")
        out = subprocess.run(["sed", "-f", os.path.join(ROOT, "integration", script),
                              os.path.join("/root/reference/RNAelem", header)], capture_output=True, text=True,
                             check=True).stdout.split("\n")
        removed = [l for l in src if l not in out]
        new = [l for l in out if l not in src]
        assert any(gone in l for l in removed) and any(added in l for l in new)
        assert len(removed) <= 6 and len(new) <= 3, (removed, new)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_patched_reference_gpu(name, tmp_path):
    assert os.path.exists(GPU_BIN), "oracle/_ref/RNAelem_gpu missing: run __graft_entry__.build() in the build container"
    clilib.check_case(GPU_BIN, name, str(tmp_path))
