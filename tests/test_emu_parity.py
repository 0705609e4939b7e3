"""CPU-side check of the kernel SOURCE logic through its single-thread host emulation (tests/emu): the same
.cu/.cuh files, compiled by g++, against the golden outputs of the unmodified reference.  This does not replace
the GPU parity tests (tests/test_gpu_parity.py); it exists so that a logic regression is caught in the GPU-less
build container.  Kept to the small cases so the CPU suite stays fast."""
import pytest

import caselib
import rnaelem_b200 as rb

SMALL = ["m0", "m1", "m3", "ragged", "nbases"]


@pytest.mark.parametrize("name", SMALL)
def test_estep_emulated(name, emu_lib):
    case = caselib.load_case(name)
    ctx = caselib.make_ctx(case, lib=emu_lib)
    caselib.check_estep(case, ctx)


@pytest.mark.parametrize("name", ["m1", "ragged"])
def test_estep_emulated_fused_diagonals(name, emu_lib, monkeypatch):
    """small chunks take one launch per diagonal and direction (relem_lin_diag_kernel); the emulation takes that path
    when RELEM_FUSE_CELLS says so"""
    monkeypatch.setenv("RELEM_FUSE_CELLS", "1000000")
    case = caselib.load_case(name)
    ctx = caselib.make_ctx(case, lib=emu_lib)
    caselib.check_estep(case, ctx)


@pytest.mark.parametrize("name", SMALL)
def test_scan_emulated(name, emu_lib):
    case = caselib.load_case(name)
    ctx = caselib.make_ctx(case, lib=emu_lib)
    caselib.check_scan(case, ctx)


def test_host_buffer_calls_reuse_the_staging_batch(emu_lib):
    """relem_scan / relem_estep keep one staging batch per context whose device buffers only grow: calls with a larger,
    then a smaller, then a different-model batch must each see their own data"""
    big, small = caselib.load_case("ragged"), caselib.load_case("m0")
    ctx = caselib.make_ctx(big, lib=emu_lib)
    caselib.check_scan(big, ctx, host_buffers=True)
    caselib.check_estep(big, ctx, via_host_call=True)
    ctx2 = caselib.make_ctx(small, lib=emu_lib)
    caselib.check_scan(small, ctx2, host_buffers=True)
    caselib.check_scan(small, ctx2, host_buffers=True)
    caselib.check_scan(big, ctx, host_buffers=True)


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


def _lik_ratio_signs(lib):
    """--lik-ratio kinds (include/relem.h): kinds 3 / 4 are the kind-1 terms with the opposite sign; 4 does not count
    towards sum_eff (motif_trainer.hpp:156-202)"""
    import numpy as np
    case = caselib.load_case("ragged")
    ctx = caselib.make_ctx(case, lib=lib)
    seqs, wss = caselib.scan_inputs(case)
    seqs, wss = seqs[3:], wss[3:]          # the reads long enough to hold the motif
    sc, off, wc = rb.pack_batch(seqs, wss)
    n = len(seqs)
    res = {}
    for kd in (rb.POS_WITH, rb.LR_WITHOUT, rb.LR_NEG):
        res[kd] = ctx.estep_run(ctx.batch(sc, off, wc, np.full(n, kd, np.uint8), np.full(n, -1, np.int32)))
    a, b, c = res[rb.POS_WITH], res[rb.LR_WITHOUT], res[rb.LR_NEG]
    assert a.n_skipped == 0 and a.fn != 0.
    for r in (b, c):
        assert caselib.close(r.fn, -a.fn)
        caselib.assert_close_vec(r.EN_diff, -np.asarray(a.EN_diff), "EN_diff sign")
        caselib.assert_close_vec(r.EH_diff, -np.asarray(a.EH_diff), "EH_diff sign")
    assert caselib.close(b.sum_eff, a.sum_eff) and c.sum_eff == 0.


def test_lik_ratio_kinds(emu_lib):
    _lik_ratio_signs(emu_lib)
