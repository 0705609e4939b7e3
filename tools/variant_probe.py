"""Non-default model settings and small patterns against the unmodified reference (needs /root/reference, CPU only):
for each variant the reference harness produces a golden case in a scratch directory and the kernel-source emulation
is checked against it like a committed golden case.  python tools/variant_probe.py [variant ...]"""
import sys, os, json, subprocess, random, math, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('', 'tests', os.path.join('tests', 'golden')):
    sys.path.insert(0, os.path.join(ROOT, p))
import make_golden as mg, caselib, conftest
subprocess.check_call(["make", "-C", conftest.EMU_DIR, "-s", "librelem_emu.so"])
OUT = os.environ.get('RELEM_PROBE_DIR', '/tmp/relem_probe')
os.makedirs(OUT, exist_ok=True)
def model(pattern="((.*.))", span=50, iloop=30, lam=(0.3, 0.6), tau=0.1, min_bpp="0.0001", extra="", seed=5, param="~T2004~"):
    t = mg.pattern_model(pattern, param, span, seed, lam=lam, min_bpp=min_bpp)
    t = t.replace("max-internal-loop: 30", "max-internal-loop: %d" % iloop).replace("tau: 0.1", "tau: %g" % tau)
    return t + extra
VARIANTS = {
  "iloop8": model(iloop=8),
  "iloop0": model(iloop=0),
  "iloop1": model(iloop=1),
  "iloop2": model(iloop=2),
  "iloop3": model(iloop=3),
  "iloop20": model(iloop=20),
  "iloop2_nofilter": model(iloop=2, min_bpp="0"),
  "iloop2_span20": model(iloop=2, span=20),
  "iloop40": model(iloop=40),
  "span12": model(span=12),
  "span8": model(span=8),
  "span6": model(span=6),
  "noprf": model(extra="no-profile: 1\n"),
  "noene": model(extra="no-energy: 1\n"),
  "bpp05": model(min_bpp="0.05"),
  "lam0": model(lam=(0.0, 1.5), tau=1.0),
  "tau0": model(tau=0.0),
  "pat_min": model(pattern="(.)"),
  "pat_dots": model(pattern=".."),
  "pat_nest": model(pattern="((..))"),
  "pat_star": model(pattern="(.*)*(.)"),
  # entries of the shipped pattern_list (and a two-hairpin pattern), Andronescu2007 for some
  "pl_a": model(pattern=".(.......)"),
  "pl_b": model(pattern="(..(...)).", param="~A2007~"),
  "pl_c": model(pattern="(...)*.....", lam=(0.5, 0.2)),
  "pl_d": model(pattern="....(*...).", param="~A2007~", span=30),
  "pl_e": model(pattern="((...*.)..)"),
  "pl_f": model(pattern="..(.(*...))", span=40),
  "two_hairpins": model(pattern="(...)(...)"),
  "two_hairpins_gap": model(pattern="(..).*(..)", param="~A2007~"),
}
names = [n for n in (sys.argv[1:] or list(VARIANTS)) if n in VARIANTS]
fq = os.path.join(OUT, "probe.fq")
mg.write_fq(fq, 3, 70, 77)
for name in names:
    mp = os.path.join(OUT, name + ".model")
    open(mp, "w").write(VARIANTS[name])
    try:
        mg.HERE = OUT   # case file goes to /tmp/probe
        mg.make_case("probe_" + name, mp, fq, 1)
        case = json.load(open(os.path.join(OUT, "case_probe_%s.json" % name)))
        ctx = caselib.make_ctx(case, lib=conftest.EMU_LIB)
        caselib.check_estep(case, ctx)
        if "scan" in case: caselib.check_scan(case, ctx)
        print("OK   ", name, flush=True)
    except BaseException as e:
        print("FAIL ", name, type(e).__name__, str(e)[:300].replace("\n", " | "), flush=True)

# unusual reads with the default model (names prefixed "in_")
READS = {
    "in_homopolymer": ["A" * 60, "U" * 45],
    "in_gc_stems": ["GGGGGGGGGGAAAACCCCCCCCCC" * 2, "GCGCGCGCGCGCGCGCGCGCGCGCGCGCGCGC"],
    "in_len1_2": ["A", "GC", "ACG", "ACGU"],
    "in_lenW": ["ACGUGCAUGCAGUCGAUCGAUGCAUGCUAGCUAGCAUGCAUCGAUGCAUGC", "ACGUGCAUGCAGUCGAUCGAUGCAUGCUAGCUAGCAUGCAUCGAUGCAUGCA",
                "ACGUGCAUGCAGUCGAUCGAUGCAUGCUAGCUAGCAUGCAUCGAUGCAUG"],
    "in_allN": ["N" * 30, "NNNNNACGUNNNNN"],
}
mp = os.path.join(OUT, "default.model")
open(mp, "w").write(mg.SYNTH_MODEL % ("~T2004~", 50, "0.0001"))
for name, seqs in READS.items():
    if sys.argv[1:] and name not in sys.argv[1:]:
        continue
    fq = os.path.join(OUT, name + ".fq")
    with open(fq, "w") as f:
        for k, sq in enumerate(seqs):
            f.write("@r%d\n%s\n+\n%s%s\n" % (k, sq, "+" * len(sq), "!" if k % 2 == 0 else "+"))
    try:
        mg.HERE = OUT
        mg.make_case("probe_" + name, mp, fq, 1)
        case = json.load(open(os.path.join(OUT, "case_probe_%s.json" % name)))
        ctx = caselib.make_ctx(case, lib=conftest.EMU_LIB)
        caselib.check_estep(case, ctx)
        if "scan" in case: caselib.check_scan(case, ctx)
        print("OK   ", name, flush=True)
    except BaseException as e:
        print("FAIL ", name, type(e).__name__, str(e)[:300].replace("\n", " | "), flush=True)
