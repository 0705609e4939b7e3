#!/usr/bin/env python3
"""Golden outputs of the UNMODIFIED reference command line (oracle/_ref/RNAelem, built by oracle/Makefile from
/root/reference) for the host-side rows of SURVEY.md 8(f): minibatch order, shuffled negatives, Adam trajectory,
train.model / train.interim / scan.raw formats.  Run in the build container (needs oracle/_ref); the outputs under
tests/golden/cli/ are committed and travel to the GPU box.

    python tests/golden/make_cli_golden.py
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "RNAelem")
OUT = os.path.join(HERE, "cli")

# name -> (sub-command or None, fastq (relative to tests/golden/_tmp), extra args, model case to read or None)
CASES = {
    # Adam over two epochs of 6 reads in minibatches of 4 (the trailing 2 reads of an epoch are skipped, the epoch
    # ends, train.interim gets a line, the reader reshuffles), then model write and scan of the training reads
    "synth_adam": (None, "synth.fq", ["-m", "((.*.))", "--max-iter", "8", "--batch-size", "4", "--lambda-init", "1.5"], None),
    # softmax parametrisation, Andronescu parameters, narrower span, different pattern
    "trna_softmax": (None, "trna.fq", ["-m", "(.....)", "--theta-softmax", "--lambda-init", "0.5", "--max-iter", "5",
                                       "--batch-size", "2", "--energy-param", "~A2007~", "-w", "40"], None),
    # full batch (--batch-size -1) on reads of unequal length, `train` sub-command (model goes to the null channel)
    "ragged_train": ("train", "ragged.fq", ["-m", "(.*)", "--max-iter", "3", "--batch-size", "-1", "--lambda-init", "0.2"], None),
    # scan of ragged reads with a model file written by the reference
    "ragged_scan": ("scan", "ragged.fq", [], "synth_adam"),
    # --no-rss through the '_' pattern spelling (application.hpp:402-407): linear profile HMM, model file says no-rss: 1
    "norss_adam": (None, "ragged.fq", ["-m", "__*_", "--max-iter", "6", "--batch-size", "3"], None),
    # likelihood-ratio objective; the second and fifth read of ragged.fq are flagged "without motif"
    "ragged_likratio": ("train", "ragged.fq", ["-m", "(.*)", "--lik-ratio", "--max-iter", "4", "--batch-size", "-1",
                                               "--lambda-init", "0.4"], None),
    # scan of 200-nt reads with a model FILE: same parameter bits on both sides -> Viterbi strings exact
    "synth_scan": ("scan", "synth.fq", [], "synth_adam"),
    # --no-shuffle: positives only, L-BFGS-B (optimizer.hpp:175-2791) instead of Adam; the sub-command-less form also
    # writes the model and scans.  Only the reference's own optimizer can run this (the patched reference binary
    # oracle/_ref/RNAelem_gpu of tests/test_reference_dropin.py); the shipped command line refuses --no-shuffle.
    "synth_lbfgsb": (None, "synth.fq", ["-m", "((.*.))", "--no-shuffle", "--max-iter", "6", "--batch-size", "-1",
                                        "--lambda-init", "0.7"], None),
    "ragged_lbfgsb": ("train", "ragged.fq", ["-m", "(.*)", "--no-shuffle", "--max-iter", "5", "--batch-size", "-1",
                                             "--lambda-init", "0.3"], None),
    # shuffled negatives only
    "genneg_k2": ("gen-neg", "ragged.fq", ["-i", "3"], None),
    "genneg_k3": ("gen-neg", "synth.fq", ["-i", "2", "--kmer-shuf", "3"], None),
    "genneg_k1": ("gen-neg", "trna.fq", ["-i", "2", "--kmer-shuf", "1"], None),
}


def main():
    assert os.path.exists(REF), "build oracle/_ref first (make -C oracle ref)"
    manifest = {}
    only = set(sys.argv[1:])
    mp = os.path.join(OUT, "manifest.json")
    if only and os.path.exists(mp):
        manifest = json.load(open(mp))
    for name, (sub, fq, extra, model_case) in CASES.items():
        if only and name not in only:
            continue
        d = os.path.join(OUT, name)
        os.makedirs(d, exist_ok=True)
        cmd = [REF] + ([sub] if sub else []) + ["-f", os.path.join(HERE, "_tmp", fq), "-t", "1"] + extra
        if model_case:
            cmd += ["-q", os.path.join(OUT, model_case, "out1.txt")]
        cmd += ["--out1", os.path.join(d, "out1.txt"), "--out2", os.path.join(d, "out2.txt"),
                "--out3", os.path.join(d, "out3.txt")]
        p = subprocess.run(cmd, capture_output=True, text=True, check=True)
        open(os.path.join(d, "stderr.txt"), "w").write(p.stderr)
        for f in ("out1.txt", "out2.txt", "out3.txt"):   # drop empty channels
            fp = os.path.join(d, f)
            if os.path.exists(fp) and os.path.getsize(fp) == 0:
                os.remove(fp)
        manifest[name] = {"sub": sub, "fastq": fq, "args": extra, "model_case": model_case}
        print(name, "ok")
    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
