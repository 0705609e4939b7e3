"""kmer-psp at scale: N positives + N negatives of 200 nt, 40 % of the positives carry a planted
motif; times rnaelem_b200/kmer-psp on all of them and script/kmer-psp.py (when /root/reference is present) on a
subsample, and checks that the two agree on the subsample."""
import os, random, subprocess, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
SUB = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
rnd = random.Random(5)
def fa(path, n, plant):
    with open(path, "w") as f:
        for k in range(n):
            s = [rnd.choice("ACGU") for _ in range(200)]
            if plant and rnd.random() < 0.4:
                p = rnd.randrange(0, 194); s[p:p + 6] = list("UGCAUG")
            f.write(">s%d\n%s\n" % (k, "".join(s)))
with tempfile.TemporaryDirectory() as d:
    P, Q = os.path.join(d, "p.fa"), os.path.join(d, "n.fa")
    fa(P, N, True); fa(Q, N, False)
    t0 = time.time()
    p = subprocess.run([os.path.join(ROOT, "rnaelem_b200", "kmer-psp"), P, Q], capture_output=True, text=True, check=True)
    t1 = time.time()
    print("kmer-psp (C++, %d threads): %d + %d sequences in %.1f s; %s; %d significant k-mer lines" %
          (os.cpu_count(), N, N, t1 - t0, [l for l in p.stderr.splitlines() if l.startswith("k:")][0], p.stderr.count("\n") - 1))
    script = "/root/reference/script/kmer-psp.py"
    if os.path.exists(script):
        Ps, Qs = os.path.join(d, "ps.fa"), os.path.join(d, "ns.fa")
        for src, dst in ((P, Ps), (Q, Qs)):
            with open(src) as f, open(dst, "w") as g:
                for k, line in enumerate(f):
                    if k >= 2 * SUB: break
                    g.write(line)
        t0 = time.time()
        r = subprocess.run([sys.executable, script, Ps, Qs], capture_output=True, text=True, check=True)
        t1 = time.time()
        c = subprocess.run([os.path.join(ROOT, "rnaelem_b200", "kmer-psp"), Ps, Qs], capture_output=True, text=True, check=True)
        t2 = time.time()
        print("subsample of %d + %d: script/kmer-psp.py %.1f s, kmer-psp %.2f s, FASTQ identical: %s" %
              (SUB, SUB, t1 - t0, t2 - t1, r.stdout == c.stdout))
