#!/usr/bin/env python3
"""Strong scaling of one objective evaluation (E-step of a minibatch + all-reduce + host update) through the shipped
command line: `RNAelem train --batch-size B --gpus N` on 200-nt reads, seconds per evaluation as the binary prints
them (`wall clock time per eval`).  python tools/strong_scaling.py --gpus N [--batches 64,128,1024] [--iters 20]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import e2e_configs as ec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--batches", default="64,128,1024")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--reads", type=int, default=4096)
    ap.add_argument("--out", default=os.path.join(ec.ROOT, "gpurun_out", "strong"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    fq = os.path.join(a.out, "reads.fq")
    ec.write_fq(fq, a.reads, 200, 5)
    res = []
    for b in [int(x) for x in a.batches.split(",")]:
        wall, per_eval, _ = ec.run([ec.BIN, "train", "-f", fq, "-m", "((.*.))", "--batch-size", str(b), "--max-iter", str(a.iters),
                                    "--out1", "/dev/null", "--out2", "/dev/null", "--out3", "/dev/null", "--gpus", str(a.gpus)],
                                   os.path.join(a.out, "train_b%d_g%d.log" % (b, a.gpus)))
        res.append({"batch_reads": b, "sequence_evaluations": 2 * b, "s_per_objective_evaluation": per_eval,
                    "sequence_evaluations_per_s": 2 * b / per_eval if per_eval else None, "wall_s": wall})
    print(json.dumps({"gpus": a.gpus, "iters": a.iters, "reads": a.reads, "runs": res}))


if __name__ == "__main__":
    main()
