"""Size-independent properties of the CUDA path at BASELINE.json's shapes (configs[1] / configs[2]: thousands of 200-nt
sequences, max-span 50, ((.*.))), where no reference run is affordable: additivity and order independence of the batch
sums, normalisation of the posteriors, consistency of the Viterbi alignment with the posterior arg-max region, and the
two independent Viterbi implementations (values-only forward + recomputed traceback, dp_vit.cuh, against the
all-in-one log-space kernel with stored traces, dp_pass.cuh) byte for byte on random reads."""
import math

import numpy as np
import pytest

import caselib
import rnaelem_b200 as rb

pytestmark = pytest.mark.gpu
L = 200


def _ctx(gpu_lib):
    case = caselib.load_case("synth200")
    return caselib.make_ctx(case, lib=gpu_lib)


def _reads(n, seed):
    rng = np.random.RandomState(seed)
    return [rng.randint(1, 5, size=L).astype(np.uint8) for _ in range(n)]


def test_estep_sums_are_additive_over_a_full_size_batch(gpu_lib):
    """fn / EN_diff / EH_diff / sum_eff of 2 x 2048 (positive, negative) pairs = the sum over its two halves and over
    the per-sequence detail (the batch reduction, the chunking and the difference channel do not lose anything)"""
    ctx = _ctx(gpu_lib)
    n = 2048
    pos, neg = _reads(n, 1), _reads(n, 2)
    seqs = [s for pair in zip(pos, neg) for s in pair]
    kind = np.array([rb.POS_WITH, rb.NEG] * n, np.uint8)
    gate = np.array([g for k in range(n) for g in (-1, 2 * k)], np.int32)
    ws = [np.zeros(L)] * (2 * n)
    sc, off, wc = rb.pack_batch(seqs, ws)
    whole = ctx.estep_run(ctx.batch(sc, off, wc, kind, gate))
    h = n  # sequences of the first half (n pairs = 2n sequences; split between pairs)
    sa, oa, wa = rb.pack_batch(seqs[:h], ws[:h])
    sb, ob, wb = rb.pack_batch(seqs[h:], ws[h:])
    a = ctx.estep_run(ctx.batch(sa, oa, wa, kind[:h], gate[:h]))
    b = ctx.estep_run(ctx.batch(sb, ob, wb, kind[h:], gate[h:] - np.where(gate[h:] >= 0, h, 0)))
    assert whole.n_skipped == 0
    assert math.isclose(whole.fn, a.fn + b.fn, rel_tol=1e-11)
    assert math.isclose(whole.sum_eff, a.sum_eff + b.sum_eff, rel_tol=1e-12)
    scale = float(np.max(np.abs(whole.EN_diff))) + 1.0
    np.testing.assert_allclose(whole.EN_diff, a.EN_diff + b.EN_diff, rtol=0, atol=1e-9 * scale)
    np.testing.assert_allclose(whole.EH_diff, a.EH_diff + b.EH_diff, rtol=1e-9, atol=1e-9)
    # detail call (two channels) against the difference channel
    d = ctx.estep_run(ctx.batch(sa, oa, wa, kind[:h], gate[:h]), detail=True)
    np.testing.assert_allclose((d.ENo - d.ENx).sum(axis=0), a.EN_diff, rtol=0, atol=1e-9 * scale)
    fn = sum(d.Z[k, 0] - (d.Z[k, 1] if kind[k] == rb.POS_WITH else d.Z[k, 2]) for k in range(h))
    assert math.isclose(fn, a.fn, rel_tol=1e-11)


def test_scan_invariants_over_thousands_of_reads(gpu_lib):
    ctx = _ctx(gpu_lib)
    n = 3000
    seqs = _reads(n, 3)
    sc, off, wc = rb.pack_batch(seqs, [np.zeros(L)] * n)
    r = ctx.scan_run(ctx.batch(sc, off, wc))
    M = ctx.M
    ps = r.PysL.reshape(n, L)
    pe = r.PyeL.reshape(n, L + 1)
    np.testing.assert_allclose(np.exp(ps).sum(axis=1), r.exist_prob, rtol=1e-12)
    assert np.all(r.exist_prob > 0) and np.all(r.exist_prob <= 1 + 1e-12)
    for k in range(n):
        # max_index (util.hpp:231-241) = LAST maximum
        assert r.Ys[k] == L - 1 - int(np.argmax(ps[k][::-1]))
        assert r.Ye[k] == L - int(np.argmax(pe[k][::-1]))
        psi = r.psihat[k * L:(k + 1) * L]
        rss = r.rss[k * L:(k + 1) * L]
        inside = np.nonzero((psi != 0) & (psi != M - 1))[0]
        if r.Ye[k] > r.Ys[k]:
            # the alignment is constrained to the posterior arg-max region: motif nodes exactly on [Ys, Ye) -- position Ye
            # is the first base emitted by the end node (the goldens of the reference show the same convention)
            assert inside.size > 0 and inside[0] == r.Ys[k] and inside[-1] == r.Ye[k] - 1, (k, r.Ys[k], r.Ye[k], inside[:3])
            assert np.all(np.diff(psi[r.Ys[k]:r.Ye[k]]) >= 0)           # the node chain is monotone
            assert np.all(psi[:r.Ys[k]] == 0) and np.all(psi[r.Ye[k]:] == M - 1)
        assert set(rss) <= set("OLRHIBM")
        assert rss.count("L") == rss.count("R")                          # every opening base has its partner


def test_viterbi_matches_the_log_space_kernel_on_random_reads(gpu_lib, monkeypatch):
    ctx = _ctx(gpu_lib)
    n = 96
    seqs = _reads(n, 4)
    sc, off, wc = rb.pack_batch(seqs, [np.zeros(L)] * n)
    a = ctx.scan_run(ctx.batch(sc, off, wc))
    assert any(t[0].startswith("relem_viterbi_kernel") for t in ctx.timing())
    monkeypatch.setenv("RELEM_PATH", "log")
    b = ctx.scan_run(ctx.batch(sc, off, wc))
    assert [t[0] for t in ctx.timing()][0] == "relem_scan_kernel"
    np.testing.assert_array_equal(a.Ys, b.Ys)
    np.testing.assert_array_equal(a.Ye, b.Ye)
    np.testing.assert_array_equal(a.psihat, b.psihat)
    assert a.rss == b.rss
    np.testing.assert_allclose(a.exist_prob, b.exist_prob, rtol=1e-9)
