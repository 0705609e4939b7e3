#!/bin/bash
# Times the shipped RNAelem (B200) and the reference binary on the same FASTQ / command line and diffs the outputs.
# usage: [SKIP_REF=1] [GPUS=n] tools/cli_compare.sh NSEQ ITER BATCH OUTDIR
set -e
N=${1:-256}; IT=${2:-5}; B=${3:-100}; OUT=${4:-gpurun_out/cli}
mkdir -p $OUT
python - "$N" "$OUT/in.fq" <<'PY'
import random, sys
n, path = int(sys.argv[1]), sys.argv[2]
random.seed(1)
with open(path, "w") as f:
    for k in range(n):
        s = "".join(random.choice("ACGU") for _ in range(200))
        f.write("@s%d\n%s\n+\n%s!\n" % (k, s, "+" * 200))
PY
T=$(nproc)
t0=$(date +%s.%N)
rnaelem_b200/RNAelem -f $OUT/in.fq -m "((.*.))" --max-iter $IT --batch-size $B --lambda-init 1.5 --gpus ${GPUS:-1} \
   --out1 $OUT/ours.model --out2 $OUT/ours.raw --out3 $OUT/ours.interim 2> $OUT/ours.err
echo "ours: $(python -c "import sys,time; print(round(time.time()-float(sys.argv[1]),2))" $t0) s wall (whole command: train $IT evaluations of $B reads + model + scan of $N reads)"
tail -4 $OUT/ours.err | cut -c1-200
if [ -x oracle/_ref/RNAelem ] && [ -z "$SKIP_REF" ]; then
t0=$(date +%s.%N)
oracle/_ref/RNAelem -f $OUT/in.fq -m "((.*.))" -t $T --max-iter $IT --batch-size $B --lambda-init 1.5 \
   --out1 $OUT/ref.model --out2 $OUT/ref.raw --out3 $OUT/ref.interim 2> $OUT/ref.err
echo "reference -t $T: $(python -c "import sys,time; print(round(time.time()-float(sys.argv[1]),2))" $t0) s wall"
tail -4 $OUT/ref.err | cut -c1-200
python - $OUT <<'PY'
import sys, os
sys.path.insert(0, "tests")
import clilib
d = sys.argv[1]
for k in ("model", "interim", "err"):
    clilib.compare_text(open(os.path.join(d, "ref." + k)).read(), open(os.path.join(d, "ours." + k)).read(), k)
    print(k, "agrees")
# scan.raw: the reference prints records in thread-completion order -> compare as sets of 10-line records
def recs(p):
    L = open(p).read().split("\n")
    return {L[i]: "\n".join(L[i:i + 10]) for i in range(0, len(L) - 1, 10)}
a, b = recs(os.path.join(d, "ref.raw")), recs(os.path.join(d, "ours.raw"))
assert a.keys() == b.keys()
bad = 0
for k in a:
    try:
        clilib.compare_text(a[k], b[k], k)
    except AssertionError as e:
        bad += 1
        print(str(e)[:300])
print("scan.raw of the train+scan command (each binary scans with its own in-memory parameters, which agree to ~1e-12;")
print("  equal-score Viterbi alternatives may then resolve differently): records", len(a), "differing:", bad)
PY
# same model FILE for both: every line, Viterbi strings included, must agree
rnaelem_b200/RNAelem scan -f $OUT/in.fq -q $OUT/ref.model --out1 $OUT/ours_scan.raw 2> $OUT/ours_scan.err
oracle/_ref/RNAelem scan -f $OUT/in.fq -q $OUT/ref.model -t $T --out1 $OUT/ref_scan.raw 2> $OUT/ref_scan.err
grep "scan end" $OUT/ours_scan.err $OUT/ref_scan.err
python - $OUT <<'PY'
import sys, os
sys.path.insert(0, "tests")
import clilib
d = sys.argv[1]
def recs(p):
    L = open(p).read().split("\n")
    return {L[i]: "\n".join(L[i:i + 10]) for i in range(0, len(L) - 1, 10)}
a, b = recs(os.path.join(d, "ref_scan.raw")), recs(os.path.join(d, "ours_scan.raw"))
assert a.keys() == b.keys()
bad = 0
for k in a:
    try:
        clilib.compare_text(a[k], b[k], k)
    except AssertionError as e:
        bad += 1
        print(str(e)[:300])
print("scan with the same model file: records", len(a), "differing:", bad)
PY
fi
