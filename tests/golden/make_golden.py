#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference and oracle/_ref/ref_harness, built by
`make -C oracle ref`).  Everything a test needs -- inputs included -- is written into the fixture, because
/root/reference does not exist on the GPU box.

    python tests/golden/make_golden.py

Fixtures:
  tables_<set>.npz      every EnergyParam table after the reference's parse   (ref_harness tables)
  hmm.json              ProfileHMM automata for a list of patterns             (ref_harness hmm)
  case_<name>.json      model + FASTQ records + the reference's E-step (per sequence and fn/gr) and scan output
  bpp_1fq.json          energy-only base-pair probabilities of RNAelem-test/1.fq: the reference's own values and
                        the RNAfold 2.3.1 dot-plot values its BPP_RNAFOLD test compares with (test-exact.cpp:90-137)
"""
import json
import math
import os
import random
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
sys.path.insert(0, ROOT)
from rnaelem_b200 import hostio  # noqa: E402


def run(*args):
    p = subprocess.run([HARNESS] + list(args), capture_output=True, text=True)
    if p.returncode != 0:
        raise SystemExit("ref_harness %s failed:\n%s" % (" ".join(args), p.stderr[-2000:]))
    return p.stdout, p.stderr


def fl(x):
    return float(x)


def parse_vec(line):
    w = line.split()
    n = int(w[1])
    v = [fl(x) for x in w[2:2 + n]]
    assert len(v) == n, line[:80]
    return v


# ------------------------------------------------------------------------------------------------- tables
def make_tables():
    for name in ("T2004", "A2007"):
        out, _ = run("tables", name)
        arrs = {}
        for line in out.split("\n"):
            if not line or line.startswith("motif"):
                continue
            w = line.split(" ", 2)
            if w[0] in ("triloops", "tetraloops", "hexaloops"):
                arrs[w[0]] = np.array(line.split('"')[1])
            else:
                arrs[w[0]] = np.array(parse_vec(line))
        np.savez_compressed(os.path.join(HERE, "tables_%s.npz" % name), **arrs)
        print("tables", name, len(arrs))


# ---------------------------------------------------------------------------------------------------- hmm
PATTERNS = ["((.*.))", "(.....)", "....", "(.*)", "..*..", ".....*.....", "(.).(.)", "(.)*(.)", "((..*(...)..))",
            ".", "*.*", "(.)", "((.).(.))", "(((...)))", ".(.*.)."]


def make_hmm():
    res = {}
    for pat in PATTERNS:
        out, _ = run("hmm", pat)
        d = {"right": {}, "left": {}, "pair": {}}
        for line in out.split("\n"):
            w = line.split()
            if not w:
                continue
            if w[0] in ("M", "S"):
                d[w[0]] = int(w[1])
            elif w[0] == "nodes":
                d["nodes"] = [ord(c) for c in w[1:]]
            elif w[0] in ("theta_id", "theta_rows", "loop_states", "reachable"):
                d[w[0]] = [int(x) for x in w[1:]]
            elif w[0] == "states":
                d["states"] = [[int(y) for y in x.split(":")] for x in w[1:]]
            elif w[0] in ("right", "left", "pair"):
                d[w[0]][w[1]] = [int(x) for x in w[2:]]
            elif w[0] == "quads":
                d["quads"] = [[int(y) for y in x.split(",")] for x in w[2:]]
        res[pat] = d
    json.dump(res, open(os.path.join(HERE, "hmm.json"), "w"))
    print("hmm", len(res))


# -------------------------------------------------------------------------------------------------- cases
def parse_estep(out):
    per, cur = [], None
    res = {}
    lines = out.split("\n")
    negseq = None
    for line in lines:
        w = line.split()
        if not w:
            continue
        if w[0] in ("pos", "neg") and len(w) > 3 and w[2] == "L":
            cur = {"tag": w[0], "id": w[1], "L": int(w[3]), "bpp_eff": fl(w[5]), "Ztt": fl(w[7]), "Ztf": fl(w[9]),
                   "Zft": fl(w[11])}
            if w[0] == "neg":
                cur["seq"] = negseq
            per.append(cur)
        elif w[0] == "skipped":
            cur["skipped"] = int(w[1])
        elif w[0] == "Zo":
            cur["Zo"], cur["Zx"] = fl(w[1]), fl(w[3])
        elif w[0] in ("ENo", "ENx", "EHo", "EHx"):
            cur[w[0]] = parse_vec(line)
        elif w[0] == "negseq":
            negseq = w[1]
        elif w[0] == "fn":
            res["fn"] = fl(w[1])
        elif w[0] == "gr":
            res["gr"] = parse_vec(line)
        elif w[0] == "sum_eff":
            res["sum_eff"] = fl(w[1])
    res["per_seq"] = per
    return res


def parse_scan(txt):
    recs, cur, EN = [], None, None
    for line in txt.split("\n"):
        if line.startswith("id: "):
            cur = {"id": line[4:]}
            recs.append(cur)
        elif line.startswith("E[N]: "):
            EN = [fl(x) for x in re.findall(r"-?inf|nan|[-+0-9.e]+", line[6:])]
        elif cur is not None and ":" in line:
            k, _, v = line.partition(":")
            v = v[1:] if v.startswith(" ") else v
            if k in ("start", "end", "inner"):
                cur[k] = [fl(x) for x in v.strip()[1:-1].split(",")]
            elif k == "psihat":
                cur[k] = [int(x) for x in v.strip()[1:-1].split(",")]
            elif k == "motif region":
                a, b = v.split(" - ")
                cur["Ys"], cur["Ye"] = int(a), int(b)
            elif k == "exist prob":
                cur["exist"] = fl(v)
            elif k in ("rss", "mot", "seq"):
                cur[k] = v
    return {"records": recs, "EN": EN}


def make_case(name, model_path, fq_path, shuffle, iteration=0, do_scan=True):
    model = hostio.read_model(model_path)
    model_text = open(model_path).read()
    recs = hostio.read_fastq(fq_path)
    case = {"name": name, "model": model, "model_text": model_text,
            "records": [{"id": r[0], "seq": r[1], "qual": r[2]} for r in recs], "shuffle": shuffle,
            "iteration": iteration}
    if not model.get("no-rss"):
        out, _ = run("estep", model_path, fq_path, str(shuffle), str(iteration))
        case["estep"] = parse_estep(out)
    else:
        out, _ = run("estep", model_path, fq_path, str(shuffle), str(iteration))
        e = parse_estep(out)
        case["estep"] = {"fn": e["fn"], "gr": e["gr"], "sum_eff": e["sum_eff"], "per_seq": e["per_seq"]}
    if do_scan:
        out, err = run("scan", model_path, fq_path)
        case["scan"] = parse_scan(out + "\n" + err)
    json.dump(case, open(os.path.join(HERE, "case_%s.json" % name), "w"))
    print("case", name, "fn", case["estep"]["fn"])


SYNTH_MODEL = """pattern: ((.*.))
theta: [[-1.38629436111989,-1.38629436111989,-1.38629436111989,-1.38629436111989],[-1.2,-1.5,-1.3,-1.6],[-1.1,-1.7,-1.4,-1.45],[-1.9,-1.7,-1.6,-1.8,-1.75,-2],[-1.5,-1.9,-1.6,-1.85,-1.7,-2.2]]
ene-param: %s
max-span: %d
max-internal-loop: 30
rho-theta: 0.1
rho-lambda: 0.1
tau: 0.1
lambda: [0.3,0.6]
min-bpp: %s
theta-softmax: 0
"""

TRNA_MODEL = """pattern: (.....)
theta: [[-1.3,-1.45,-1.35,-1.4],[-1.2,-1.5,-1.3,-1.6],[-1.1,-1.7,-1.4,-1.45],[-1.25,-1.55,-1.3,-1.5],[-1.0,-1.8,-1.5,-1.4],[-1.6,-1.2,-1.4,-1.35],[-1.9,-1.7,-1.6,-1.8,-1.75,-2]]
ene-param: ~T2004~
max-span: 50
max-internal-loop: 30
rho-theta: 0.1
rho-lambda: 0.1
tau: 0.1
lambda: [0.5,1.2]
min-bpp: 0.0001
theta-softmax: 0
"""


def write_fq(path, n, L, seed, plant=False):
    rnd = random.Random(seed)
    with open(path, "w") as f:
        for k in range(n):
            s = "".join(rnd.choice("ACGU") for _ in range(L))
            q = ["+"] * L
            # a few non-flat pseudo-qualities, like kmer-psp.py emits around enriched k-mers
            for _ in range(rnd.randrange(0, 6)):
                p = rnd.randrange(0, L)
                q[p] = rnd.choice("*,-")
            flag = "!" if (k % 3 != 2) else "+"
            f.write("@syn%d\n%s\n+\n%s%s\n" % (k, s, "".join(q), flag))


def make_cases():
    T = os.path.join(REF, "RNAelem-test")
    tmp = os.path.join(HERE, "_tmp")
    os.makedirs(tmp, exist_ok=True)
    make_case("m0", os.path.join(T, "0.model"), os.path.join(T, "0.fq"), 0)
    make_case("m1", os.path.join(T, "1.model"), os.path.join(T, "0.fq"), 0)
    make_case("m2", os.path.join(T, "2.model"), os.path.join(T, "0.fq"), 0)
    make_case("m3", os.path.join(T, "3.model"), os.path.join(T, "0.fq"), 0)
    # synthetic 200-nt set of BASELINE config 2 (pattern ((.*.)), W=50, T2004) with in-binary shuffled negatives
    mp = os.path.join(tmp, "synth.model")
    open(mp, "w").write(SYNTH_MODEL % ("~T2004~", 50, "0.0001"))
    fq = os.path.join(tmp, "synth.fq")
    write_fq(fq, 6, 200, 1)
    make_case("synth200", mp, fq, 1)
    # iteration counter changes the negatives
    make_case("synth200_it3", mp, fq, 1, iteration=3, do_scan=False)
    # ragged lengths incl. shorter than the span and shorter than any pair
    fq2 = os.path.join(tmp, "ragged.fq")
    with open(fq2, "w") as f:
        rnd = random.Random(7)
        for k, L in enumerate([3, 7, 12, 30, 51, 77, 130]):
            s = "".join(rnd.choice("ACGU") for _ in range(L))
            f.write("@rag%d\n%s\n+\n%s%s\n" % (k, s, "+" * L, "!" if k % 2 else "+"))
    make_case("ragged", mp, fq2, 1)
    # no BPP filter
    mp0 = os.path.join(tmp, "synth_nofilter.model")
    open(mp0, "w").write(SYNTH_MODEL % ("~T2004~", 50, "0"))
    fq3 = os.path.join(tmp, "synth2.fq")
    write_fq(fq3, 2, 120, 5)
    make_case("nofilter", mp0, fq3, 1)
    # long-span Andronescu case of BASELINE config 5 (scaled down)
    mpa = os.path.join(tmp, "a2007.model")
    open(mpa, "w").write(SYNTH_MODEL % ("~A2007~", 150, "0.0001"))
    fq4 = os.path.join(tmp, "long.fq")
    write_fq(fq4, 2, 320, 4)
    make_case("a2007_w150", mpa, fq4, 1)
    # tRNA-like toy (BASELINE config 1 shape: pattern (.....), S=29)
    mpt = os.path.join(tmp, "trna.model")
    open(mpt, "w").write(TRNA_MODEL)
    fq5 = os.path.join(tmp, "trna.fq")
    with open(fq5, "w") as f:
        seqs = ["GCGGAUUUAGCUCAGUUGGGAGAGCGCCAGACUGAAGAUCUGGAGGUCCUGUGUUCGAUCCACAGAAUUCGCACCA",
                "GGGGCUAUAGCUCAGCUGGGAGAGCGCUUGCAUGGCAUGCAAGAGGUCAGCGGUUCGAUCCCGCUUAGCUCCACCA",
                "GCCCGGAUAGCUCAGUCGGUAGAGCAGGGGAUUGAAAAUCCCCGUGUCCUUGGUUCGAUUCCGAGUCCGGGCACCA"]
        for k, s in enumerate(seqs):
            f.write("@trna%d\n%s\n+\n%s!\n" % (k, s, "+" * len(s)))
    make_case("trna", mpt, fq5, 1)


# ---------------------------------------------------------------------------------------------------- bpp
def make_bpp():
    T = os.path.join(REF, "RNAelem-test")
    tmp = os.path.join(HERE, "_tmp")
    mp = os.path.join(tmp, "bpp.model")
    open(mp, "w").write(SYNTH_MODEL % ("~T2004~", 50, "0.0001"))
    fq = os.path.join(T, "1.fq")
    out, _ = run("bpp", mp, fq)
    recs = hostio.read_fastq(fq)
    d = {"records": [{"id": r[0], "seq": r[1], "qual": r[2]} for r in recs], "model": hostio.read_model(mp)}
    for line in out.split("\n"):
        w = line.split()
        if not w:
            continue
        if w[0] == "seq":
            d["L"], d["W"], d["C"], d["bpp_eff"] = int(w[3]), int(w[5]), int(w[7]), fl(w[9])
        elif w[0] in ("bp_ok", "left_ok"):
            d[w[0]] = [[int(y) for y in x.split(",")] for x in w[1:]]
        elif w[0] == "lnbpp":
            d["lnbpp"] = [[int(x.split(",")[0]), int(x.split(",")[1]), fl(x.split(",")[2])] for x in w[1:]]
        elif w[0] == "lnZ":
            d["lnZ"] = fl(w[1])
    # RNAfold dot plot: "i j sqrt(p) ubox" lines (the reference's test squares the third column)
    rnafold = []
    for line in open(os.path.join(T, "1.0.ps")):
        w = line.split()
        if len(w) == 4 and w[3] == "ubox":
            try:
                rnafold.append([int(w[0]), int(w[1]), float(w[2])])
            except ValueError:
                pass
    d["rnafold_ubox"] = rnafold
    json.dump(d, open(os.path.join(HERE, "bpp_1fq.json"), "w"))
    print("bpp", d["L"], len(d["lnbpp"]), len(rnafold))


def pattern_model(pattern, param, span, seed, lam=(0.3, 0.6), min_bpp="0.0001"):
    """model text for any pattern: background row, one 4-vector per '.', one 6-vector per ')' (node order), seeded
    log-probabilities"""
    rnd = random.Random(seed)

    def row(n):
        v = [rnd.uniform(0.5, 1.5) for _ in range(n)]
        t = sum(v)
        return "[" + ",".join("%.6g" % math.log(x / t) for x in v) + "]"
    rows = [row(4)] + [row(4) if c == "." else row(6) for c in pattern if c in ".)"]
    return ("pattern: %s\ntheta: [%s]\nene-param: %s\nmax-span: %d\nmax-internal-loop: 30\nrho-theta: 0.1\n"
            "rho-lambda: 0.1\ntau: 0.1\nlambda: [%g,%g]\nmin-bpp: %s\ntheta-softmax: 0\n"
            % (pattern, ",".join(rows), param, span, lam[0], lam[1], min_bpp))


def make_cases_round2():
    """shapes the round-1 verdict found unpinned: configs[4] at its real size (1000 nt, max-span 150, Andronescu2007),
    the largest automaton of the shipped pattern_list (S=91) and a two-stem pattern_list entry."""
    tmp = os.path.join(HERE, "_tmp")
    os.makedirs(tmp, exist_ok=True)
    mpa = os.path.join(tmp, "a2007.model")
    open(mpa, "w").write(SYNTH_MODEL % ("~A2007~", 150, "0.0001"))
    fq = os.path.join(tmp, "long1000.fq")
    write_fq(fq, 2, 1000, 11)
    make_case("long1000", mpa, fq, 1)
    mp = os.path.join(tmp, "s91.model")
    open(mp, "w").write(pattern_model(".....*.....", "~T2004~", 50, 21))
    fq = os.path.join(tmp, "s91.fq")
    write_fq(fq, 2, 120, 12)
    make_case("s91", mp, fq, 1)
    mp = os.path.join(tmp, "plstem.model")
    open(mp, "w").write(pattern_model("(.(..*..).)", "~A2007~", 80, 22, lam=(0.4, 0.9)))
    fq = os.path.join(tmp, "plstem.fq")
    write_fq(fq, 3, 150, 13)
    make_case("plstem", mp, fq, 1)


def make_case_nbases():
    """reads with unknown bases (code 0: cannot pair, emitted like any base), lower case and T for U
    (bio_sequence.hpp:28-41)"""
    tmp = os.path.join(HERE, "_tmp")
    os.makedirs(tmp, exist_ok=True)
    mp = os.path.join(tmp, "synth.model")
    if not os.path.exists(mp):
        open(mp, "w").write(SYNTH_MODEL % ("~T2004~", 50, "0.0001"))
    fq = os.path.join(tmp, "nbases.fq")
    rnd = random.Random(31)
    with open(fq, "w") as f:
        for k in range(4):
            L = (90, 64, 120, 33)[k]
            s = [rnd.choice("ACGU") for _ in range(L)]
            for _ in range((6, 2, 15, 0)[k]):
                s[rnd.randrange(0, L)] = rnd.choice("NnXR-")
            for z in range(L):
                if rnd.random() < 0.15:
                    s[z] = {"A": "a", "C": "c", "G": "g", "U": rnd.choice("tTu")}.get(s[z], s[z])
            if k == 3:
                s[10:14] = list("NNNN")   # a run of unknown bases inside a short read
            f.write("@nb%d\n%s\n+\n%s%s\n" % (k, "".join(s), "+" * L, "!" if k != 1 else "+"))
    make_case("nbases", mp, fq, 1)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "round2":
        make_cases_round2()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "nbases":
        make_case_nbases()
        sys.exit(0)
    make_tables()
    make_hmm()
    make_cases()
    make_bpp()
    make_cases_round2()
    make_case_nbases()
