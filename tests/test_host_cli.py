"""Host side of `RNAelem train / scan` (rnaelem_b200/host) against the reference binary's
golden outputs.  The CPU half runs the command line linked with the single-thread emulation of the kernel source
(tests/emu; a debug aid, never the product) so that minibatch order, negatives, Adam and the writers are checked in
the GPU-less container; the GPU half runs the shipped binary rnaelem_b200/RNAelem (librelem.so, sm_100a)."""
import os
import subprocess

import pytest

import clilib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT = os.path.join(ROOT, "rnaelem_b200", "RNAelem")
GENNEG = ["genneg_k1", "genneg_k2", "genneg_k3"]
DP_CASES = ["synth_adam", "trna_softmax", "ragged_train", "ragged_scan", "norss_adam", "ragged_likratio", "synth_scan"]


@pytest.fixture(scope="session")
def emu_cli(emu_lib):
    subprocess.check_call(["make", "-C", os.path.dirname(emu_lib), "-s", "RNAelem_emu"])
    return os.path.join(os.path.dirname(emu_lib), "RNAelem_emu")


@pytest.fixture(scope="session")
def product_cli():
    assert os.path.exists(PRODUCT), "rnaelem_b200/RNAelem missing: run __graft_entry__.build()"
    return PRODUCT


@pytest.mark.parametrize("name", GENNEG)
def test_shuffled_negatives_bit_exact(name, product_cli, tmp_path):
    """gen-neg needs no device: the k-let shuffle + glibc rand() stream must reproduce the reference's negatives"""
    got = clilib.run_case(product_cli, name, str(tmp_path))
    assert got["out1"] == open(os.path.join(clilib.CLI_GOLDEN, name, "out1.txt")).read()


def test_product_refuses_without_gpu(product_cli, tmp_path):
    """no CPU path: without a CUDA device the shipped binary must fail loudly, not compute"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    c = clilib.MANIFEST["ragged_scan"]
    p = subprocess.run([product_cli, "scan", "-f", os.path.join(clilib.HERE, "golden", "_tmp", c["fastq"]),
                        "-q", os.path.join(clilib.CLI_GOLDEN, "synth_adam", "out1.txt"), "--out1", str(tmp_path / "x")],
                       capture_output=True, text=True)
    assert p.returncode == 1 and "no CUDA device" in p.stderr


@pytest.mark.parametrize("name", DP_CASES)
def test_cli_emulated(name, emu_cli, tmp_path):
    clilib.check_case(emu_cli, name, str(tmp_path))


def test_scan_rounds_keep_input_order(emu_cli, tmp_path):
    """the scanner works in rounds of RELEM_SCAN_CHUNK reads per GPU, formatted one round behind on host threads:
    rounds of 3 reads must print the same records, in input order, and the same E[N]"""
    clilib.check_case(emu_cli, "ragged_scan", str(tmp_path), extra_env={"RELEM_SCAN_CHUNK": "3"})


def test_unknown_option_and_subcommand(product_cli):
    p = subprocess.run([product_cli, "--no-such-flag"], capture_output=True, text=True)
    assert p.returncode == 1
    p = subprocess.run([product_cli, "frobnicate", "-f", "x"], capture_output=True, text=True)
    assert p.returncode == 1 and "unknown sub-command" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", DP_CASES)
def test_cli_gpu(name, product_cli, tmp_path):
    clilib.check_case(product_cli, name, str(tmp_path))


@pytest.mark.gpu
def test_cli_two_gpus(product_cli, tmp_path):
    """minibatch sharded over two contexts + NCCL all-reduce of the partial sums: same trajectory"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    clilib.check_case(product_cli, "synth_adam", str(tmp_path), extra_args=["--gpus", "2"])


def test_model_written_here_is_read_by_the_reference(emu_cli, tmp_path):
    """drop-in in the other direction: a train.model written by this host layer must be accepted by the unmodified
    reference binary, and both must then print the same scan records (same parameter bits -> exact Viterbi lines)"""
    ref = os.path.join(ROOT, "oracle", "_ref", "RNAelem")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/RNAelem not built")
    fq = os.path.join(clilib.HERE, "golden", "_tmp", "trna.fq")
    model = str(tmp_path / "ours.model")
    p = subprocess.run([emu_cli, "-f", fq, "-m", "(.*.)", "--max-iter", "3", "--batch-size", "2", "--lambda-init", "0.7",
                        "--out1", model, "--out2", str(tmp_path / "x.raw"), "--out3", str(tmp_path / "x.interim")],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    outs = {}
    for name, binary in (("ref", ref), ("ours", emu_cli)):
        raw = str(tmp_path / (name + ".raw"))
        p = subprocess.run([binary, "scan", "-f", fq, "-q", model, "-t", "1", "--out1", raw], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        outs[name] = (open(raw).read(), p.stderr)
    clilib.compare_text(outs["ref"][0], outs["ours"][0], "scan.raw with a model written here")
    clilib.compare_text(outs["ref"][1], outs["ours"][1], "scan stderr")


def _eval_case(binary, name, tmp_path):
    """`RNAelem eval` through the whole host stack (FASTQ reader, weights, negatives, packing, chain rule, lambda slots)
    against the fn / gr the unmodified reference computed for the same model and reads (tests/golden/case_*.json)."""
    import re
    import caselib
    case = caselib.load_case(name)
    model = tmp_path / "m.model"
    fq = tmp_path / "in.fq"
    model.write_text(case["model_text"])
    with open(fq, "w") as f:
        for r in case["records"]:
            f.write("%s\n%s\n+\n%s\n" % (r["id"], r["seq"], "".join(chr(33 + q) for q in r["qual"])))
    cmd = [binary, "eval", "-f", str(fq), "-q", str(model), "--out1", str(tmp_path / "fn.txt"), "--out2", str(tmp_path / "gr.txt")]
    if not case["shuffle"]:
        cmd.append("--no-shuffle")
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    fn = float((tmp_path / "fn.txt").read_text().split(": ")[1])
    gr = [float(x) for x in re.findall(r"-?(?:inf|nan|\d+\.?\d*(?:[eE][-+]?\d+)?)", (tmp_path / "gr.txt").read_text().split(": ")[1])]
    est = case["estep"]
    assert caselib.close(fn, est["fn"]), (fn, est["fn"])
    import numpy as np
    summ = max(1.0, max(float(np.max(np.abs(e["ENo"]))) for e in est["per_seq"] if not e["skipped"]))
    caselib.assert_close_vec(gr, est["gr"], name + " eval gr", scale=summ)


@pytest.mark.parametrize("name", ["m0", "m1", "m2", "m3", "trna", "ragged"])
def test_eval_emulated(name, emu_cli, tmp_path):
    _eval_case(emu_cli, name, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["m0", "m1", "m2", "m3", "trna", "ragged", "synth200", "a2007_w150"])
def test_eval_gpu(name, product_cli, tmp_path):
    _eval_case(product_cli, name, tmp_path)


@pytest.mark.gpu
def test_cli_two_gpus_failing_rank_does_not_hang(product_cli, tmp_path):
    """a rank whose E-step fails must still enter the all-reduce (the failure travels as one more summed element):
    the process ends with the rank's message instead of blocking the other rank in NCCL forever"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    c = clilib.MANIFEST["synth_adam"]
    cmd = [product_cli, "-f", os.path.join(clilib.HERE, "golden", "_tmp", c["fastq"])] + c["args"] + \
          ["--gpus", "2", "--out1", str(tmp_path / "m"), "--out2", str(tmp_path / "r"), "--out3", str(tmp_path / "i")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env=dict(os.environ, RELEM_TEST_FAIL_RANK="1"))
    assert p.returncode == 1 and "relem_estep failed on GPU 1" in p.stderr, p.stderr[-500:]
