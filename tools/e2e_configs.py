#!/usr/bin/env python3
"""BASELINE.json configs[3] and configs[4] end to end through the shipped command line (rnaelem_b200/RNAelem) on
N GPUs, with the unmodified reference binary timed on a stated subsample of the same commands.

  configs[3]  100 000 x 200-nt reads, max-span 50, Turner2004, pattern ((.*.)): `RNAelem train` with the `elem` defaults
              (minibatches of 64, 300 iterations) on one fold, then `RNAelem scan` of the held-out fold with the trained
              model.  --gpus N shards every minibatch / the scan over N GPUs of one process.
  configs[4]  5 000 x 1000-nt reads, max-span 150, Andronescu2007, the first P patterns of the shipped pattern list
              (S = 15 .. 91 over the whole list): per pattern `RNAelem train` (64 x ITERS) + `RNAelem scan` of a held-out
              sample.  Patterns are independent models: they are spread over the N GPUs as N concurrent processes (the
              second, embarrassingly parallel axis of SURVEY.md 8e), one GPU each.

  python tools/e2e_configs.py --config 3 --gpus 1 [--reads 100000] [--iters 300] [--ref]
prints one JSON line; logs go to --out (default gpurun_out/e2e)."""
import argparse
import json
import os
import random
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "rnaelem_b200", "RNAelem")
REF = os.path.join(ROOT, "oracle", "_ref", "RNAelem")
# the first entries of the reference's pattern_list with a stem, plus its largest automaton
PATTERNS4 = ["(...).....", "((...))...", "(.(...))..", "((...*))...", "(....*)....", ".....*....."]


def write_fq(path, n, L, seed, plant="GGACUUCGGUCC", frac=0.5):
    rnd = random.Random(seed)
    with open(path, "w") as f:
        for k in range(n):
            s = [rnd.choice("ACGU") for _ in range(L)]
            if rnd.random() < frac:
                p = rnd.randrange(0, L - len(plant))
                s[p:p + len(plant)] = list(plant)
            f.write("@r%d\n%s\n+\n%s!\n" % (k, "".join(s), "+" * L))


def run(cmd, log, env=None):
    t0 = time.time()
    p = subprocess.run(cmd, capture_output=True, text=True, env=env)
    dt = time.time() - t0
    open(log, "w").write(" ".join(cmd) + "\n" + p.stderr[-20000:])
    if p.returncode != 0:
        raise SystemExit("command failed (%d): %s\n%s" % (p.returncode, " ".join(cmd), p.stderr[-2000:]))
    ev = re.search(r"wall clock time per eval: ([0-9.eE+-]+)", p.stderr)
    sc = re.search(r"scan end: ([0-9.eE+-]+)", p.stderr)
    return dt, (float(ev.group(1)) if ev else None), (float(sc.group(1)) if sc else None)


def config3(a):
    d = a.out
    tr, ho = os.path.join(d, "c3_train.fq"), os.path.join(d, "c3_heldout.fq")
    half = a.reads // 2
    write_fq(tr, half, 200, 3)
    write_fq(ho, half, 200, 33)
    model = os.path.join(d, "c3_g%d.model" % a.gpus)
    g = ["--gpus", str(a.gpus)]
    # the sub-command-less form (what script/elem calls): train, write the model to --out1, scan the training reads
    # (`RNAelem train` alone sends its model to the null channel, main.cpp:118-119)
    t_train, per_eval, scan_self = run([BIN, "-f", tr, "-m", "((.*.))", "--batch-size", "64", "--max-iter", str(a.iters),
                                        "--out1", model, "--out2", "/dev/null", "--out3", os.path.join(d, "c3_g%d.interim" % a.gpus)] + g,
                                       os.path.join(d, "c3_train_g%d.log" % a.gpus))
    t_scan, _, scan_end = run([BIN, "scan", "-f", ho, "-q", model, "--out1", os.path.join(d, "c3_g%d.raw" % a.gpus)] + g,
                              os.path.join(d, "c3_scan_g%d.log" % a.gpus))
    raw_mb = os.path.getsize(os.path.join(d, "c3_g%d.raw" % a.gpus)) / 1e6
    os.remove(os.path.join(d, "c3_g%d.raw" % a.gpus))
    line = {"config": "configs[3]", "gpus": a.gpus, "reads_train": half, "reads_scan": half, "iters": a.iters, "batch": 64,
            "train_write_scan_wall_s": t_train, "s_per_objective_evaluation": per_eval,
            "scan_of_training_fold_reads_per_s": half / scan_self if scan_self else None,
            "seq_evals_per_s_train": 128 / per_eval if per_eval else None,
            "scan_wall_s": t_scan, "scan_end_s": scan_end, "scan_reads_per_s": half / scan_end if scan_end else None,
            "scan_raw_MB": raw_mb}
    if a.ref and os.path.exists(REF):
        cores = os.cpu_count()
        sub = os.path.join(d, "c3_ref.fq")
        write_fq(sub, 2048, 200, 3)
        _, ref_eval, _ = run([REF, "train", "-f", sub, "-m", "((.*.))", "--batch-size", "64", "--max-iter", "3", "-t", str(cores),
                              "--out1", "/dev/null", "--out2", "/dev/null", "--out3", "/dev/null"], os.path.join(d, "c3_ref_train.log"))
        sub2 = os.path.join(d, "c3_ref_scan.fq")
        write_fq(sub2, 32 * cores, 200, 33)
        _, _, ref_scan = run([REF, "scan", "-f", sub2, "-q", model, "-t", str(cores), "--out1", "/dev/null"],
                             os.path.join(d, "c3_ref_scan.log"))
        line["reference"] = {"cores": cores, "s_per_objective_evaluation": ref_eval, "sample_train": "3 iterations x 64 reads",
                             "scan_reads_per_s": 32 * cores / ref_scan, "sample_scan": "%d reads" % (32 * cores)}
    return line


def config4(a):
    d = a.out
    tr, ho = os.path.join(d, "c4_train.fq"), os.path.join(d, "c4_heldout.fq")
    write_fq(tr, a.reads, 1000, 4)
    write_fq(ho, a.scan_reads, 1000, 44)
    pats = PATTERNS4[:a.patterns]
    t0 = time.time()
    procs, res = [], []
    # one process per pattern, GPUs handed out round robin, at most a.gpus processes at a time
    queue = list(enumerate(pats))
    running = {}
    while queue or running:
        while queue and len(running) < a.gpus:
            k, pat = queue.pop(0)
            free = [g for g in range(a.gpus) if g not in [r[0] for r in running.values()]][0]
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(free))
            model = os.path.join(d, "c4_p%d.model" % k)
            sh = ("%s -f %s -m '%s' -w 150 --energy-param '~A2007~' --batch-size 64 --max-iter %d --out1 %s --out2 /dev/null "
                  "--out3 /dev/null 2> %s && %s scan -f %s -q %s --out1 /dev/null 2> %s" %
                  (BIN, tr, pat, a.iters, model, os.path.join(d, "c4_p%d_train.log" % k), BIN, ho, model,
                   os.path.join(d, "c4_p%d_scan.log" % k)))
            running[k] = (free, subprocess.Popen(sh, shell=True, env=env), time.time(), pat)
        for k in list(running):
            g, p, ts, pat = running[k]
            if p.poll() is not None:
                if p.returncode != 0:
                    raise SystemExit("pattern %s failed: see %s" % (pat, os.path.join(d, "c4_p%d_*.log" % k)))
                tl = open(os.path.join(d, "c4_p%d_train.log" % k)).read()
                sl = open(os.path.join(d, "c4_p%d_scan.log" % k)).read()
                ev = re.search(r"wall clock time per eval: ([0-9.eE+-]+)", tl)
                sc = re.search(r"scan end: ([0-9.eE+-]+)", sl)
                res.append({"pattern": pat, "gpu": g, "wall_s": time.time() - ts,
                            "s_per_objective_evaluation": float(ev.group(1)) if ev else None,
                            "scan_reads_per_s": a.scan_reads / float(sc.group(1)) if sc else None})
                del running[k]
        time.sleep(0.2)
    wall = time.time() - t0
    line = {"config": "configs[4]", "gpus": a.gpus, "reads": a.reads, "scan_reads": a.scan_reads, "iters": a.iters, "batch": 64,
            "patterns": pats, "wall_s": wall, "per_pattern": res}
    if a.ref and os.path.exists(REF):
        cores = os.cpu_count()
        sub = os.path.join(d, "c4_ref.fq")
        write_fq(sub, cores, 1000, 4)
        _, ref_eval, _ = run([REF, "train", "-f", sub, "-m", pats[0], "-w", "150", "--energy-param", "~A2007~", "--batch-size",
                              str(cores), "--max-iter", "1", "-t", str(cores), "--out1", "/dev/null", "--out2", "/dev/null",
                              "--out3", "/dev/null"], os.path.join(d, "c4_ref_train.log"))
        line["reference"] = {"cores": cores, "pattern": pats[0], "s_per_objective_evaluation_of_%d_reads" % cores: ref_eval,
                             "seq_evals_per_s": 2 * cores / ref_eval}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 4])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--reads", type=int, default=0)
    ap.add_argument("--scan-reads", type=int, default=512)
    ap.add_argument("--iters", type=int, default=0)
    ap.add_argument("--patterns", type=int, default=4)
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "e2e"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    if a.config == 3:
        a.reads = a.reads or 100000
        a.iters = a.iters or 300
        line = config3(a)
    else:
        a.reads = a.reads or 5000
        a.iters = a.iters or 20
        line = config4(a)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
