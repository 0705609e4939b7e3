#!/usr/bin/env python3
"""Golden outputs of the reference's preprocessing script (script/kmer-psp.py: k-mer enrichment by Fisher's exact test ->
per-base pseudo-qualities in FASTQ) for the host tool rnaelem_b200/kmer-psp.  Run in the build
container (needs /root/reference and scipy); inputs and outputs are committed under tests/golden/kmer/.

    python tests/golden/make_kmer_golden.py
"""
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "kmer")
SCRIPT = "/root/reference/script/kmer-psp.py"


def write_fa(path, n, L, seed, plant=None, frac=0.0, alphabet="ACGU"):
    rnd = random.Random(seed)
    with open(path, "w") as f:
        for k in range(n):
            s = [rnd.choice(alphabet) for _ in range(L + rnd.randrange(-5, 6))]
            if plant and rnd.random() < frac:
                for m in plant:
                    p = rnd.randrange(0, len(s) - len(m))
                    s[p:p + len(m)] = list(m)
            f.write(">seq%d some annotation\n%s\n" % (k, "".join(s)))


def run(name, pos, neg):
    cmd = [sys.executable, SCRIPT, pos] + ([neg] if neg else [])
    env = dict(os.environ, PYTHONHASHSEED="0")
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, check=True)
    open(os.path.join(OUT, name + ".fq"), "w").write(p.stdout)
    open(os.path.join(OUT, name + ".err"), "w").write("".join(sorted(p.stderr.splitlines(True))))
    print(name, len(p.stdout), "bytes,", p.stderr.count("\n"), "stderr lines")


def main():
    os.makedirs(OUT, exist_ok=True)
    # enriched motifs of two lengths (one self-overlapping: re.finditer matches do not overlap), a depleted one
    write_fa(os.path.join(OUT, "a_pos.fa"), 120, 60, 1, plant=["GGACU", "AAAAAA"], frac=0.6)
    write_fa(os.path.join(OUT, "a_neg.fa"), 150, 60, 2, plant=["CUCUC"], frac=0.5)
    run("a", os.path.join(OUT, "a_pos.fa"), os.path.join(OUT, "a_neg.fa"))
    # DNA alphabet, lower enrichment, more sequences
    write_fa(os.path.join(OUT, "b_pos.fa"), 300, 40, 3, plant=["TGCATG"], frac=0.35, alphabet="ACGT")
    write_fa(os.path.join(OUT, "b_neg.fa"), 300, 40, 4, alphabet="ACGT")
    run("b", os.path.join(OUT, "b_pos.fa"), os.path.join(OUT, "b_neg.fa"))
    # no negative set: flat qualities
    run("c", os.path.join(OUT, "a_pos.fa"), None)


if __name__ == "__main__":
    main()
