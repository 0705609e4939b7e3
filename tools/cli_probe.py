"""Option combinations of the command line that have no committed golden: the unmodified reference binary and the shipped
host code linked with the kernel-source emulation (tests/emu/RNAelem_emu) run the same command, every output channel
is compared like a golden CLI case (tests/clilib.py).  CPU only, needs oracle/_ref/RNAelem.  python tools/cli_probe.py [case ...]"""
import sys, os, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT,'tests'))
import clilib
REF=os.path.join(ROOT,'oracle','_ref','RNAelem'); EMU=os.path.join(ROOT,'tests','emu','RNAelem_emu')
T=os.path.join(ROOT,'tests','golden','_tmp')
CASES = {
 "tau_rho": ["-f", T+"/ragged.fq", "-m", "(.*)", "--max-iter", "4", "--batch-size", "-1", "--tau", "0.3", "--rho-theta", "0.5", "--rho-lambda", "0.01", "--lambda-init", "0.6"],
 "lambda_prior": ["-f", T+"/ragged.fq", "-m", "(.*)", "--max-iter", "4", "--batch-size", "3", "--lambda-prior", "0.8", "--lambda-init", "0.2"],
 "minbpp_span": ["-f", T+"/synth.fq", "-m", "((.*.))", "--max-iter", "3", "--batch-size", "-1", "-p", "0.01", "-w", "30"],
 "noprofile": ["-f", T+"/ragged.fq", "-m", "(.*)", "--max-iter", "3", "--batch-size", "-1", "--no-profile"],
 "kmer3_likratio": ["-f", T+"/synth.fq", "-m", "((.*.))", "--max-iter", "3", "--batch-size", "2", "--kmer-shuf", "3", "--lik-ratio"],
 "softmax_rhos": ["-f", T+"/trna.fq", "-m", "(.....)", "--theta-softmax", "--rho-s", "0.3", "--max-iter", "4", "--batch-size", "-1"],
 "epsilon": ["-f", T+"/ragged.fq", "-m", "(.*)", "--max-iter", "30", "--batch-size", "-1", "--epsilon", "0.05"],
 "nbases_cli": ["-f", T+"/nbases.fq", "-m", "((.*.))", "--max-iter", "3", "--batch-size", "-1"],
 "iloop4": ["-f", T+"/synth.fq", "-m", "((.*.))", "--max-iter", "3", "--batch-size", "-1", "-c", "4"],
 "batch_gt_n": ["-f", T+"/synth.fq", "-m", "((.*.))", "--max-iter", "4", "--batch-size", "10"],
 "batch_1": ["-f", T+"/ragged.fq", "-m", "(.*)", "--max-iter", "9", "--batch-size", "1"],
 "iter_0": ["-f", T+"/ragged.fq", "-m", "(.*)", "--max-iter", "0", "--batch-size", "-1"],
 "two_stems_train": ["-f", T+"/plstem.fq", "-m", "(.(..*..).)", "--max-iter", "2", "--batch-size", "-1", "--energy-param", "~A2007~", "-w", "80"],
}
for name, args in CASES.items():
    if sys.argv[1:] and name not in sys.argv[1:]: continue
    outs={}
    ok=True
    for tag, binary in (("ref", REF), ("emu", EMU)):
        d=tempfile.mkdtemp(prefix="cli_%s_%s_"%(name,tag))
        cmd=[binary]+args+["-t","1","--out1",d+"/o1","--out2",d+"/o2","--out3",d+"/o3"]
        p=subprocess.run(cmd,capture_output=True,text=True,timeout=1500)
        outs[tag]={"rc":p.returncode,"stderr":p.stderr,**{k:(open(d+"/"+k).read() if os.path.exists(d+"/"+k) else "") for k in ("o1","o2","o3")}}
    try:
        assert outs["ref"]["rc"]==outs["emu"]["rc"], (outs["ref"]["rc"], outs["emu"]["rc"], outs["emu"]["stderr"][-300:])
        # `log sum:` is the reference scanner's warning that its start posteriors do not add up to one
        # (motif_scanner.hpp:212-213): it fires where its inside and outside passes enumerate different loops
        # (-c below 30); the numbers of those runs are compared, the warning is not reproduced
        outs["ref"]["stderr"] = "\n".join(l for l in outs["ref"]["stderr"].split("\n") if not l.startswith("log sum:"))
        for k in ("stderr","o1","o2","o3"):
            clilib.compare_text(outs["ref"][k], outs["emu"][k], name+"/"+k, tie_tolerant=(k=="o2"))
        print("OK   ", name, flush=True)
    except BaseException as e:
        print("FAIL ", name, type(e).__name__, str(e)[:500].replace("\n"," | "), flush=True)
