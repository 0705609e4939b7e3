"""Host-side input glue: FASTQ and train.model readers, position weights, batch packing.

Formats follow the reference: FASTQ records are strictly four lines (RNAelem/fastq_io.hpp:85-104), qualities are
sanger (base 33) with one extra trailing value that flags "contains the motif" (motif_model.hpp:62-70);
train.model is `key: value` per line (RNAelem/motif_io.hpp:29-57,118-262).
"""
import math
import numpy as np

_CODE = {"A": 1, "a": 1, "C": 2, "c": 2, "G": 3, "g": 3, "T": 4, "t": 4, "U": 4, "u": 4}


def seq_codes(s):
    """bio_sequence.hpp:30-41: ACGU(T) any case -> 1..4, everything else 0."""
    return np.fromiter((_CODE.get(ch, 0) for ch in s), dtype=np.uint8, count=len(s))


def read_fastq(path):
    """-> list of (id line incl. '@', sequence string, quality ints (value - 33, length L+1))."""
    recs = []
    with open(path) as f:
        lines = f.read().split("\n")
    k = 0
    # a record counts only when its fourth line is newline-terminated (the reference drops a record whose
    # quality line hits EOF, fastq_io.hpp:93-95)
    while k + 4 <= len(lines) - 1:
        rid, seq, _, qual = lines[k], lines[k + 1], lines[k + 2], lines[k + 3]
        k += 4
        recs.append((rid, seq, [ord(c) - 33 for c in qual]))
    return recs


def quality_to_ws(qual):
    """RNAelem::set_ws: ws[i] = ln((0.01+q_i)/(0.01+mode)), last entry -> -inf if 0 else 0. Returns L+1 values."""
    cnt = [0] * (127 - 33)
    for q in qual:
        cnt[q] += 1
    mode, best = 0, -1
    for i, c in enumerate(cnt):
        if best <= c:
            mode, best = i, c
    ws = [math.log((0.01 + float(q)) / (0.01 + mode)) for q in qual[:-1]]
    ws.append(-math.inf if qual[-1] == 0 else 0.0)
    return ws


def _parse_vec(s):
    s = s.strip()
    if s.startswith("[["):
        rows = s[2:-2].split("],[")
        return [[float(x) for x in r.split(",")] for r in rows]
    if s.startswith("["):
        return [float(x) for x in s[1:-1].split(",") if x != ""]
    return s


def read_model(path):
    """train.model -> dict with the reference's keys (values parsed)."""
    m = {}
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if ": " not in line:
                continue
            k, v = line.split(": ", 1)
            m[k] = _parse_vec(v)
    out = {
        "pattern": m["pattern"],
        "ene-param": m.get("ene-param", "~T2004~"),
        "max-span": int(float(m["max-span"])),
        "max-internal-loop": int(float(m["max-internal-loop"])),
        "theta-softmax": int(float(m.get("theta-softmax", 0))),
        "tau": float(m["tau"]),
        "lambda": [float(x) for x in m["lambda"]],
        "min-bpp": float(m["min-bpp"]),
        "no-rss": int(float(m.get("no-rss", 0))),
        "no-profile": int(float(m.get("no-profile", 0))),
        "no-energy": int(float(m.get("no-energy", 0))),
    }
    if "s" in m:
        out["s"] = m["s"]
    if "theta" in m:
        out["theta"] = m["theta"]
    if "_" in out["pattern"]:
        out["pattern"] = out["pattern"].replace("_", ".")
        out["no-rss"] = 1
    return out


def _logsumexp2(x, y):
    if y == -math.inf:
        return x
    if x == -math.inf:
        return y
    return y + math.log1p(math.exp(x - y)) if x < y else x + math.log1p(math.exp(y - x))


def model_theta_flat(model):
    """theta rows as the DP uses them: given directly, or softmax of `s` (ProfileHMM::calc_theta, profile_hmm.hpp:103-111)."""
    if model.get("theta-softmax") and "s" in model:
        rows = []
        for r in model["s"]:
            tot = -math.inf
            for e in r:
                tot = _logsumexp2(tot, e)
            rows.append([e - tot for e in r])
    else:
        rows = model["theta"]
    return np.array([x for r in rows for x in r], dtype=np.float64)


def band_cells(L, W):
    W = min(L, W)
    return (L + 1) * (W + 1) - W * (W + 1) // 2


def pack_batch(seqs, ws_list):
    """seqs: list of uint8 code arrays; ws_list: list of L-long float arrays -> (seq_cat, off, ws_cat)."""
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    for n, s in enumerate(seqs):
        off[n + 1] = off[n] + len(s)
    seq_cat = np.concatenate(seqs).astype(np.uint8) if seqs else np.zeros(0, np.uint8)
    ws_cat = np.concatenate([np.asarray(w, dtype=np.float64)[:len(s)] for w, s in zip(ws_list, seqs)]) \
        if seqs else np.zeros(0)
    return np.ascontiguousarray(seq_cat), off, np.ascontiguousarray(ws_cat)
