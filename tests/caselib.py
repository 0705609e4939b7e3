"""Shared helpers of the parity tests: load a golden case (tests/golden/case_*.json, produced by the unmodified
reference through oracle/ref_harness.cpp), run it through a librelem context, compare.

Tolerances (BASELINE.json north_star): log-likelihood, posteriors, exist prob and gradient within 1e-9 relative in
fp64; Viterbi psihat / rss / motif region bit-exact.
"""
import json
import math
import os

import numpy as np

import rnaelem_b200 as rb

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
RTOL = 1e-9


def load_case(name):
    return json.load(open(os.path.join(GOLDEN, "case_%s.json" % name)))


def case_names():
    return sorted(f[5:-5] for f in os.listdir(GOLDEN) if f.startswith("case_") and f.endswith(".json"))


def close(a, b, rtol=RTOL):
    a, b = float(a), float(b)
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= rtol * max(1.0, abs(a), abs(b))


def assert_close_vec(a, b, what, rtol=RTOL, scale=None):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    s = scale if scale is not None else 1.0
    for k, (x, y) in enumerate(zip(a.ravel(), b.ravel())):
        if math.isnan(x) or math.isnan(y):
            assert math.isnan(x) and math.isnan(y), "%s[%d]: %r vs %r" % (what, k, x, y)
        elif math.isinf(x) or math.isinf(y):
            assert x == y, "%s[%d]: %r vs %r" % (what, k, x, y)
        else:
            assert abs(x - y) <= rtol * max(s, abs(x), abs(y)), "%s[%d]: %r vs %r" % (what, k, x, y)


def estep_inputs(case):
    """-> (seqs, ws, kind, gate, expect) in the order the reference evaluates: each positive followed by the
    negative the reference shuffled from it (the fixture records that negative's sequence)."""
    recs = {r["id"]: r for r in case["records"]}
    seqs, wss, kind, gate, expect = [], [], [], [], []
    last_pos = -1
    pos_iter = iter(case["records"])
    for e in case["estep"]["per_seq"]:
        if e["tag"] == "pos":
            r = next(pos_iter)
            assert r["id"] == e["id"]
            ws = rb.quality_to_ws(r["qual"])
            seqs.append(rb.seq_codes(r["seq"]))
            wss.append(ws[:-1])
            kind.append(rb.POS_WITH if ws[-1] == -math.inf else rb.POS_WITHOUT)
            gate.append(-1)
            last_pos = len(seqs) - 1
        else:
            L = len(e["seq"])
            ws = rb.quality_to_ws([0] * (L + 1))
            seqs.append(rb.seq_codes(e["seq"]))
            wss.append(ws[:-1])
            kind.append(rb.NEG)
            gate.append(last_pos)
        expect.append(e)
    return seqs, wss, kind, gate, expect


def make_ctx(case, lib=None, device=0):
    ctx = rb.Context(device, lib=lib)
    ctx.set_model(case["model"])
    return ctx


def gradient_from(case, r):
    """RNAelemTrainDP's update block (motif_trainer.hpp:248-271): softmax chain rule + lambda slots."""
    model = case["model"]
    theta = rb.model_theta_flat(model)
    en = np.asarray(r.EN_diff)
    gr = []
    if model.get("theta-softmax"):
        k = 0
        rows = model["s"]
        for row in rows:
            n = len(row)
            d = en[k:k + n]
            tot = float(np.sum(d))
            for j in range(n):
                p = math.exp(theta[k + j])
                gr.append((1 - p) * d[j] - p * (tot - d[j]))
            k += n
    else:
        gr = list(en)
    lam = model["lambda"]
    eh = list(r.EH_diff)
    if lam[0] == lam[1]:
        eh = [eh[0] + eh[1], 0.0]
    return np.array(gr + eh)


def check_estep(case, ctx, via_host_call=False):
    seqs, wss, kind, gate, expect = estep_inputs(case)
    sc, off, wc = rb.pack_batch(seqs, wss)
    if via_host_call:
        r = ctx.estep(sc, off, wc, kind, gate, detail=True)
    else:
        b = ctx.batch(sc, off, wc, kind, gate)
        r = ctx.estep_run(b, detail=True)
        b.close()
    for n, e in enumerate(expect):
        tag = "%s[%d %s %s]" % (case["name"], n, e["tag"], e["id"])
        assert close(r.Z[n, 0], e["Ztt"]), "%s Ztt %r vs %r" % (tag, r.Z[n, 0], e["Ztt"])
        assert close(r.Z[n, 1], e["Ztf"]), "%s Ztf %r vs %r" % (tag, r.Z[n, 1], e["Ztf"])
        assert close(r.Z[n, 2], e["Zft"]), "%s Zft %r vs %r" % (tag, r.Z[n, 2], e["Zft"])
        assert (r.skipped[n] == 1) == bool(e["skipped"]), "%s skipped" % tag
        if e["skipped"]:
            continue
        if not math.isnan(e["bpp_eff"]) and e["tag"] == "pos":
            assert close(r.bpp_eff[n], e["bpp_eff"], 1e-12), "%s bpp_eff" % tag
        scale = max(1.0, float(np.max(np.abs(e["ENo"]))) if e["ENo"] else 1.0)
        assert_close_vec(r.ENo[n], e["ENo"], tag + " ENo", scale=scale)
        assert_close_vec(r.ENx[n], e["ENx"], tag + " ENx", scale=scale)
        # lambda slots are chosen by value in the reference; compare slot sums when the lambdas coincide
        lam = case["model"]["lambda"]
        eho, ehx = r.EH[n, 0:2], r.EH[n, 2:4]
        if lam[0] == lam[1]:
            eho, ehx = [eho[0] + eho[1], 0.0], [ehx[0] + ehx[1], 0.0]
        hs = max(1.0, float(np.max(np.abs(e["EHo"]))))
        assert_close_vec(eho, e["EHo"], tag + " EHo", scale=hs)
        assert_close_vec(ehx, e["EHx"], tag + " EHx", scale=hs)
    est = case["estep"]
    assert close(r.fn, est["fn"]), "%s fn %r vs %r" % (case["name"], r.fn, est["fn"])
    if not math.isnan(est["sum_eff"]):
        assert close(r.sum_eff, est["sum_eff"]), "%s sum_eff" % case["name"]
    gr = gradient_from(case, r)
    gs = max(1.0, float(np.max(np.abs(est["gr"]))))
    # the gradient is a difference of O(ENo) sums: tolerance relative to the size of the summands
    summ = max(gs, max(float(np.max(np.abs(e["ENo"]))) for e in expect if not e["skipped"]))
    assert_close_vec(gr, est["gr"], case["name"] + " gr", scale=summ)
    return r


def scan_inputs(case):
    seqs, wss = [], []
    for r in case["records"]:
        ws = rb.quality_to_ws(r["qual"])
        seqs.append(rb.seq_codes(r["seq"]))
        wss.append(ws[:-1])
    return seqs, wss


def check_scan(case, ctx, host_buffers=False):
    seqs, wss = scan_inputs(case)
    sc, off, wc = rb.pack_batch(seqs, wss)
    if host_buffers:   # relem_scan: host buffers in, the context's staging batch underneath
        r = ctx.scan(sc, off, wc)
    else:
        b = ctx.batch(sc, off, wc)
        r = ctx.scan_run(b)
        b.close()
    gold = case["scan"]
    assert len(gold["records"]) == len(seqs)
    M = ctx.M
    nodes = ctx.hmm_get(6)
    for n, g in enumerate(gold["records"]):
        o, L = int(off[n]), int(off[n + 1] - off[n])
        tag = "%s scan[%d %s]" % (case["name"], n, g["id"])
        degenerate = any(math.isnan(x) for x in g["end"])
        assert_close_vec(r.PysL[o:o + L], g["start"], tag + " start")
        assert_close_vec(r.PyiL[o:o + L], g["inner"], tag + " inner")
        assert close(r.exist_prob[n], g["exist"]), "%s exist %r vs %r" % (tag, r.exist_prob[n], g["exist"])
        assert int(r.Ys[n]) == g["Ys"], "%s Ys %d vs %d" % (tag, r.Ys[n], g["Ys"])
        if degenerate:
            # motif cannot start anywhere (Z of the start-constrained pass is -inf): the reference's end
            # posteriors are -inf - (-inf) = NaN; here they are -inf.  Everything else must still agree.
            assert all(x == -math.inf for x in r.PyeL[o + n:o + n + L + 1]), tag + " end (degenerate)"
        else:
            assert_close_vec(r.PyeL[o + n:o + n + L + 1], g["end"], tag + " end")
        assert int(r.Ye[n]) == g["Ye"], "%s Ye %d vs %d" % (tag, r.Ye[n], g["Ye"])
        assert list(map(int, r.psihat[o:o + L])) == g["psihat"], tag + " psihat"
        assert r.rss[o:o + L] == g["rss"], "%s rss\n%s\n%s" % (tag, r.rss[o:o + L], g["rss"])
        mot = "".join(" " if (h == 0 or h == M - 1) else chr(nodes[h]) for h in r.psihat[o:o + L])
        assert mot == g["mot"], tag + " mot"
    assert_close_vec(r.EN, gold["EN"], case["name"] + " E[N]", scale=max(1.0, float(np.max(np.abs(gold["EN"])))))
    return r
