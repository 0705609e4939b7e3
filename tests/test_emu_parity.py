"""CPU-side check of the kernel SOURCE logic through its single-thread host emulation (tests/emu): the same
.cu/.cuh files, compiled by g++, against the golden outputs of the unmodified reference.  This does not replace
the GPU parity tests (tests/test_gpu_parity.py); it exists so that a logic regression is caught in the GPU-less
build container.  Kept to the small cases so the CPU suite stays fast."""
import pytest

import caselib
import rnaelem_b200 as rb

SMALL = ["m0", "m1", "m3", "ragged"]


@pytest.mark.parametrize("name", SMALL)
def test_estep_emulated(name, emu_lib):
    case = caselib.load_case(name)
    ctx = caselib.make_ctx(case, lib=emu_lib)
    caselib.check_estep(case, ctx)


@pytest.mark.parametrize("name", SMALL)
def test_scan_emulated(name, emu_lib):
    case = caselib.load_case(name)
    ctx = caselib.make_ctx(case, lib=emu_lib)
    caselib.check_scan(case, ctx)


@pytest.mark.parametrize("name", ["m0", "ragged"])
def test_exterior_row_rescaling(name, emu_renorm_lib):
    """exterior rows are mantissa + power-of-two exponent per position; with the thresholds at 0.5 / 2 every column is
    rescaled and E-step and scan must still match the reference"""
    case = caselib.load_case(name)
    ctx = caselib.make_ctx(case, lib=emu_renorm_lib)
    caselib.check_estep(case, ctx)
    caselib.check_scan(case, ctx)


def test_no_rss_linear_model(emu_lib):
    """--no-rss (motif_model.hpp:170-219): exterior-row recursion only; fixture m2 comes from the reference"""
    case = caselib.load_case("m2")
    ctx = caselib.make_ctx(case, lib=emu_lib)
    caselib.check_estep(case, ctx)
    caselib.check_scan(case, ctx)
    seqs, wss = caselib.scan_inputs(case)
    sc, off, wc = rb.pack_batch(seqs, wss)
    with pytest.raises(rb.RelemError):
        ctx.bpp(ctx.batch(sc, off, wc))   # no base-pair filter in this mode
