"""Pins the plain-C oracle (oracle/relem_oracle.c) to the reference:
  (a) the reference's own known-answer tests: PATH_COUNT_CASES / EMISSION_COUNT_CASES (RNAelem-test/test.cpp:93-203,
      run under its debug switches NO_THETA|NO_ENE|FIX_RSS|NO_TURN) and BPP_RNAFOLD (test-exact.cpp:90-137);
  (b) outputs of the unmodified reference compiled here (tests/golden, made by tests/golden/make_golden.py)."""
import json
import math
import os

import numpy as np
import pytest

import caselib
import rnaelem_b200 as rb
from oraclelib import Oracle

LARGE = 2 ** 31 - 1

# (pattern, sequence, fixed structure, number of motif alignments) -- RNAelem-test/test.cpp:105-178
PATH_COUNTS = [
    (".", "A", ".", 2), (".", "AA", "..", 4), (".", "CAAAG", "(...)", 7), (".", "ACAAAGA", ".(...).", 9),
    (".", "ACACAAAGGA", ".(.(...)).", 10), (".", "ACACAGACAGAAGA", ".(.(.).(.)..).", 10), (".", "CACAGAG", "(.(.).)", 4),
    ("(.)", "CAAAG", "(...)", 2), ("(.)", "CCAAAGG", "((...))", 3),
    ("(.*)", "CAAAG", "(...)", 4), ("(.*)", "CCAAAGG", "((...))", 7),
    (".*.", "AA", "..", 2), (".*.", "CAAAG", "(...)", 6),
    ("(.).(.)", "CAGACAG", "(.).(.)", 2), ("(.).(.)", "CCAGACAGG", "((.).(.))", 2),
    ("(.)*(.)", "CAGCAG", "(.)(.)", 2), ("(.)*(.)", "CCAGCAGG", "((.)(.))", 2),
]
# RNAelem-test/test.cpp:181-203 (counts accumulate nothing across calls: ENo is cleared per eval)
EMISSION_COUNTS = [
    ("A", ".", [[1, 0, 0, 0], [1, 0, 0, 0]]),
    ("CAG", "(.)", [[1, 2, 2, 0], [1, 0, 0, 0]]),
    ("CACGG", "(...)", [[4, 10, 11, 0], [3, 4, 3, 0]]),
    ("CAGAU", "(.)..", [[7, 5, 5, 3], [3, 0, 0, 2]]),
]


def _debug_oracle(pattern, rss):
    o = Oracle(pattern, "~T2004~", LARGE, LARGE, 0.0, 0, 0, 1)   # set_energy_params("~T2004~",large,large,0.,true)
    n = o.n_theta
    o.set_params(np.zeros(n), [1.0, 1.0], 1.0)                    # set_hyper_param(0,0,0,tau=1,-1)
    o.set_debug(1, 1, rss)
    return o


@pytest.mark.parametrize("pattern,seq,rss,count", PATH_COUNTS)
def test_reference_path_counts(pattern, seq, rss, count):
    o = _debug_oracle(pattern, rss)
    ws = rb.quality_to_ws([1] * (len(seq) + 1))
    pf, pfo, _, _ = o.debug_eval(rb.seq_codes(seq), ws)
    assert math.isclose(math.exp(pf), count, rel_tol=1e-12)
    assert math.isclose(math.exp(pfo), count, rel_tol=1e-12)


@pytest.mark.parametrize("seq,rss,expect", EMISSION_COUNTS)
def test_reference_emission_counts(seq, rss, expect):
    o = _debug_oracle(".", rss)
    ws = rb.quality_to_ws([1] * (len(seq) + 1))
    _, _, EN, _ = o.debug_eval(rb.seq_codes(seq), ws)
    np.testing.assert_allclose(EN, np.array(expect, dtype=float).ravel(), rtol=1e-12, atol=1e-12)


TABLES = ["hairpin", "mismatch_h", "mismatch_i", "mismatch_m", "mismatch_1ni", "mismatch_23i", "mismatch_ext", "stack",
          "bulge", "term_au", "int11", "int21", "int22", "internal", "dangle5", "dangle3", "ninio", "mlintern",
          "mlclosing", "ml_base", "lxc37"]


def _valid_mask(name, n):
    """entries the reference defines.  Its int22 fill clears 8 000 of 40 000 entries (energy_param.hpp:597-598): the
    rest, where the file has no value ('N' bases), reads as +0.0 in its binaries and is reproduced here entry by entry
    (reads with N bases depend on it, golden case nbases).  Row 7 of the 8-row reads of mismatch_multi/exterior spills
    into the next member."""
    if name == "int22":   # pair types 1..6 on both sides: every entry a loop can index; the others hold stack garbage
        idx = np.arange(n)
        t2, t1 = (idx // 625) % 8, idx // 5000
        return (t1 >= 1) & (t1 <= 6) & (t2 >= 1) & (t2 <= 6)
    if name in ("mismatch_1ni",):
        return np.arange(n) >= 25     # row 0 receives the overflow of mismatch_m's 8th row in the reference
    return np.ones(n, dtype=bool)


@pytest.mark.parametrize("which", ["T2004", "A2007"])
def test_energy_tables_match_reference_parse(which):
    g = np.load(os.path.join(caselib.GOLDEN, "tables_%s.npz" % which))
    o = Oracle(".", "~%s~" % which)
    for name in TABLES:
        a, b = o.energy_get(name), g[name]
        n = min(len(a), len(b))
        mask = _valid_mask(name, n)
        np.testing.assert_array_equal(a[:n][mask], b[:n][mask], err_msg=name)
    for name, key in (("triloop", "triloops"), ("tetraloop", "tetraloops"), ("hexaloop", "hexaloops")):
        k = len(str(g[key]).split())
        np.testing.assert_array_equal(o.energy_get(name)[:k], g[name][:k])


def _hmm_check(get, d):
    S = d["S"]
    assert get(0) == [x for s in d["states"] for x in s[1:]]
    assert get(1) == d["loop_states"]
    for kind, key in ((2, "right"), (3, "left"), (4, "pair")):
        v = get(kind)
        off, idx = v[:S + 1], v[S + 1:]
        for s in range(S):
            assert idx[off[s]:off[s + 1]] == d[key][str(s)], (key, s)
    assert get(5) == [x for q in d["quads"] for x in q]
    assert get(6) == d["nodes"]
    assert get(7) == d["theta_id"]
    assert get(9) == d["reachable"]


def test_automaton_matches_reference():
    g = json.load(open(os.path.join(caselib.GOLDEN, "hmm.json")))
    for pat, d in g.items():
        o = Oracle(pat)
        assert (o.M, o.S) == (d["M"], d["S"]), pat
        _hmm_check(o.hmm_get, d)
        assert o.hmm_get(8) == d["theta_rows"]


def test_bpp_matches_reference_and_rnafold():
    g = json.load(open(os.path.join(caselib.GOLDEN, "bpp_1fq.json")))
    m = g["model"]
    o = Oracle(".", m["ene-param"], m["max-span"], m["max-internal-loop"], m["min-bpp"])
    seq = rb.seq_codes(g["records"][0]["seq"])
    bp, lf, ln, eff, lnz = o.bpp(seq)
    L, W = g["L"], g["W"]
    assert caselib.close(lnz, g["lnZ"], 1e-13) and caselib.close(eff, g["bpp_eff"], 1e-15)
    assert {(i, d) for i in range(L + 1) for d in range(W + 1) if bp[i * (W + 1) + d]} == {tuple(x) for x in g["bp_ok"]}
    assert {(i, d) for i in range(L + 1) for d in range(W + 1) if lf[i * (W + 1) + d]} == {tuple(x) for x in g["left_ok"]}
    for i, d, v in g["lnbpp"]:
        assert caselib.close(ln[i * (W + 1) + d], v, 1e-12)
    n = 0
    for i1, j1, sp in g["rnafold_ubox"]:
        i, j = i1 - 1, j1
        if j - i <= W:
            assert abs(ln[i * (W + 1) + (j - i)] - 2 * math.log(sp)) < 1e-5
            n += 1
    assert n > 500


ORACLE_CASES = ["m0", "m1", "m3", "ragged", "trna", "nofilter", "nbases"]


@pytest.mark.parametrize("name", ORACLE_CASES)
def test_estep_matches_reference(name):
    case = caselib.load_case(name)
    o = Oracle.from_model(case["model"])
    seqs, wss, kind, gate, expect = caselib.estep_inputs(case)
    for seq, ws, kd, e in zip(seqs, wss, kind, expect):
        r = o.estep_seq(seq.astype(np.int32), list(ws) + [0.0], 1 if kd == rb.POS_WITH else 0, kd == rb.NEG)
        tag = "%s %s %s" % (name, e["tag"], e["id"])
        for a, b in zip(r["Z"], (e["Ztt"], e["Ztf"], e["Zft"])):
            assert caselib.close(a, b, 1e-13), tag
        assert r["skipped"] == e["skipped"], tag
        if e["skipped"]:
            continue
        sc = max(1.0, float(np.max(np.abs(e["ENo"]))))
        caselib.assert_close_vec(r["ENo"], e["ENo"], tag + " ENo", 1e-12, sc)
        caselib.assert_close_vec(r["ENx"], e["ENx"], tag + " ENx", 1e-12, sc)
        lam = case["model"]["lambda"]
        if lam[0] != lam[1]:
            caselib.assert_close_vec(r["EHo"], e["EHo"], tag + " EHo", 1e-12, max(1.0, float(np.max(np.abs(e["EHo"])))))
            caselib.assert_close_vec(r["EHx"], e["EHx"], tag + " EHx", 1e-12, max(1.0, float(np.max(np.abs(e["EHo"])))))


@pytest.mark.parametrize("name", ["m0", "m1", "m3", "ragged", "trna", "nbases"])
def test_scan_matches_reference(name):
    case = caselib.load_case(name)
    o = Oracle.from_model(case["model"])
    seqs, wss = caselib.scan_inputs(case)
    EN = np.zeros(o.n_theta)
    for seq, ws, g in zip(seqs, wss, case["scan"]["records"]):
        r = o.scan_seq(seq.astype(np.int32), list(ws) + [0.0])
        tag = "%s %s" % (name, g["id"])
        caselib.assert_close_vec(r["PysL"], g["start"], tag + " start", 1e-12)
        caselib.assert_close_vec(r["PyiL"], g["inner"], tag + " inner", 1e-12)
        caselib.assert_close_vec(r["PyeL"], g["end"], tag + " end", 1e-12)   # NaNs of the degenerate case included
        assert (r["Ys"], r["Ye"]) == (g["Ys"], g["Ye"]), tag
        assert list(map(int, r["psihat"])) == g["psihat"], tag
        assert r["rss"] == g["rss"], tag
        assert caselib.close(r["exist"], g["exist"], 1e-12), tag
        EN += r["EN"]
    caselib.assert_close_vec(EN, case["scan"]["EN"], name + " E[N]", 1e-12, max(1.0, float(np.max(np.abs(EN)))))
