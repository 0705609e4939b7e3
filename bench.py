#!/usr/bin/env python3
"""bench.py -- throughput of the RNAelem E-step hot path on B200 (BASELINE.json configs[1]).

Workload (per GPU, weak scaling): a synthetic eCLIP-like batch of NSEQ x 200-nt i.i.d. ACGU positives (all flagged
"contains motif") plus one shuffled negative each, pattern ((.*.)), max-span 50, Turner2004, min-bpp 1e-4, model
parameters as `elem train` has them at iteration 0 (uniform theta, lambda = lambda-init = 0, tau 0.1).
One step = one objective evaluation of RNAelemTrainer::operator() over that batch: energy-only base-pair filter +
coupled inside + outside with expected counts for every positive and negative, batch reduction of (fn, gr), and at
N > 1 the NCCL all-reduce of those P+3 doubles.

metric  dp_cells_per_s = band cells x 3 coupled passes (inside + 2 outside, the reference's pass count) per second,
        whole job.  `value`: batch resident in HBM.  `e2e`: the host-buffer entry point relem_estep (H2D of the
        batch from pinned memory and D2H of the result inside the timed region).
roofline  the E-step is one wavefront of small kernels (relem_lin_phase_kernel<phase> per span and phase, 592 launches
        per chunk of sequences); `achieved` = algorithmic bytes (cells x 168 x S) / summed device time of
        those launches (CUDA events on the launch stream) against the measured HBM copy bandwidth; `traffic` = measured
        DRAM bytes of the same launches (ncu dram__bytes_read+write, profiles/r1_dram_traffic.json) scaled to the batch.
cpu_baseline / --impl reference  the unmodified reference binary (oracle/_ref/RNAelem train ... --max-iter 1) on
        all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PATTERN = "((.*.))"
SEQ_LEN = 200
MAX_SPAN = 50
S_STATES = 22          # interval states of ((.*.))
BYTES_PER_CELL = 168   # x S: write inside once, read it in both outside passes
PASSES = 3


def cells(L, W):
    W = min(L, W)
    return (L + 1) * (W + 1) - W * (W + 1) // 2


def make_dataset(nseq, seed):
    rng = np.random.RandomState(seed)
    pos = rng.randint(1, 5, size=(nseq, SEQ_LEN)).astype(np.uint8)
    # negatives: seeded shuffle of each positive (composition preserving; the inputs are i.i.d. so a
    # dinucleotide-preserving shuffle has the same distribution)
    neg = np.stack([p[rng.permutation(SEQ_LEN)] for p in pos])
    return pos, neg


def pack(pos, neg):
    import rnaelem_b200 as rb
    n = len(pos)
    seqs, kind, gate = [], [], []
    for k in range(n):
        seqs.append(pos[k]); kind.append(rb.POS_WITH); gate.append(-1)
        seqs.append(neg[k]); kind.append(rb.NEG); gate.append(2 * k)
    seq_cat = np.ascontiguousarray(np.concatenate(seqs))
    off = np.arange(0, (2 * n + 1) * SEQ_LEN, SEQ_LEN, dtype=np.int64)
    ws = np.zeros(2 * n * SEQ_LEN)  # flat quality '+' x L -> ws = ln(1) = 0
    return seq_cat, off, ws, np.array(kind, np.uint8), np.array(gate, np.int32)


def uniform_model():
    rows = [4, 4, 4, 6, 6]  # background, '.', '.', ')', ')'
    theta = np.concatenate([np.full(r, -math.log(r)) for r in rows])
    return theta, [0.0, 0.0], 0.1


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                w = [x.strip() for x in out.strip().split(",")]
                if len(w) >= 8:
                    self.rows.append(w)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = []
        for k, name in ((4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"), (7, "sw_power_cap")):
            if any(r[k] == "Active" for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------- reference arm
def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "RNAelem")
    return p if os.path.exists(p) else None


def run_reference_once(nseq, seed, threads):
    """objective evaluation of the unmodified reference over nseq positives (+ its own shuffled negatives);
    returns (seconds per evaluation as the binary prints it, sequence evaluations)."""
    pos, _ = make_dataset(nseq, seed)
    with tempfile.TemporaryDirectory() as d:
        fq = os.path.join(d, "x.fq")
        with open(fq, "w") as f:
            for k in range(nseq):
                f.write("@s%d\n%s\n+\n%s!\n" % (k, "".join("NACGU"[c] for c in pos[k]), "+" * SEQ_LEN))
        cmd = [reference_binary(), "train", "-f", fq, "-m", PATTERN, "-w", str(MAX_SPAN), "-t", str(threads),
               "--batch-size", "-1", "--max-iter", "1", "--out1", "/dev/null", "--out2", "/dev/null",
               "--out3", "/dev/null"]
        t0 = time.perf_counter()
        p = subprocess.run(cmd, capture_output=True, text=True)
        wall = time.perf_counter() - t0
    m = re.search(r"wall clock time per eval: ([0-9.eE+-]+)", p.stderr + p.stdout)
    if p.returncode != 0 or not m:
        raise RuntimeError("reference run failed: " + (p.stderr[-500:]))
    return float(m.group(1)), 2 * nseq, wall


def cpu_baseline(sample_pos):
    cores = os.cpu_count() or 1
    if reference_binary() is None:
        return {"value": None, "unit": "dp_cells/s", "cores": cores, "kind": "reference",
                "sample": "oracle/_ref/RNAelem missing"}
    sec, evals, _ = run_reference_once(sample_pos, 12345, cores)
    v = evals * cells(SEQ_LEN, MAX_SPAN) * PASSES / sec
    return {"value": v, "unit": "dp_cells/s", "cores": cores, "kind": "reference",
            "seq_evals_per_s": evals / sec,
            "sample": "%d positives + %d in-binary shuffled negatives x %d nt, RNAelem train -t %d --batch-size -1 "
                      "--max-iter 1 (wall clock time per eval)" % (sample_pos, sample_pos, SEQ_LEN, cores)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    sample = args.ref_sample or max(16, 4 * cores)
    if reference_binary() is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/RNAelem was not built"}))
        return 0
    for _ in range(args.warmup):
        run_reference_once(max(8, cores), 999, cores)
    tot_s, tot_e = 0.0, 0
    for k in range(args.steps):
        sec, evals, _ = run_reference_once(sample, 12345 + k, cores)
        tot_s += sec; tot_e += evals
    v = tot_e * cells(SEQ_LEN, MAX_SPAN) * PASSES / tot_s
    line = {"impl": "reference", "metric": "dp_cells_per_s", "value": v, "unit": "dp_cells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(sample, 1),
            "cpu_baseline": {"value": v, "unit": "dp_cells/s", "cores": cores, "kind": "reference",
                             "seq_evals_per_s": tot_e / tot_s,
                             "sample": "%d positives + negatives per step, all host cores" % sample},
            "e2e": {"value": v, "unit": "dp_cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(npos, ngpu):
    return {"workload": "RNAelem E-step (configs[1]): %d x %d-nt synthetic positives + %d shuffled negatives per GPU, "
                        "pattern %s, max-span %d, Turner2004, min-bpp 1e-4, iteration-0 parameters"
                        % (npos, SEQ_LEN, npos, PATTERN, MAX_SPAN),
            "positives_per_gpu": npos, "seq_len": SEQ_LEN, "max_span": MAX_SPAN, "pattern": PATTERN,
            "states": S_STATES, "parallelism": "dp%d (sequences sharded, one all-reduce of fn/gr per step)" % ngpu,
            "cache": "working set (DP tables of all resident sequences, > 1 GB) exceeds the 126 MB L2"}


# -------------------------------------------------------------------------------------------------- own arm
def main_own(args):
    import torch
    import rnaelem_b200 as rb
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (librelem has no CPU path)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = rb.Context(local)
    ctx.set_energy("~T2004~", MAX_SPAN, 30, 1e-4, 0)
    ctx.set_pattern(PATTERN)
    theta, lam, tau = uniform_model()
    ctx.set_params(theta, lam, tau)
    NT = ctx.n_theta
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            import ctypes
            buf = (ctypes.c_uint8 * 128)()
            rc = ctx.lib.relem_comm_unique_id(buf)
            assert rc == 0, "ncclGetUniqueId failed"
            uid = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().tolist()), rank, world)

    npos = args.nseq
    pos, neg = make_dataset(npos, 1000 + rank)
    seq_cat, off, ws, kind, gate = pack(pos, neg)
    batch = ctx.batch(seq_cat, off, ws, kind, gate)
    step_cells = batch.cells * PASSES
    step_bytes = batch.cells * BYTES_PER_CELL * S_STATES

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        r = ctx.estep_run(batch)
        if world > 1:
            v = np.concatenate([[r.fn, r.sum_eff, float(r.n_skipped)], r.EN_diff, r.EH_diff])
            v = ctx.allreduce_sum(v)
        return r

    kernel_ms, launches = [], 0
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_resident()
        tm = ctx.timing()
        kernel_ms.append(sum(t[1] for t in tm if t[0] in ("relem_estep_lin_kernel", "relem_estep_kernel")))
        kname = max((t for t in tm if t[2] > 0), key=lambda t: t[1])[0]
        launches += sum(t[2] for t in tm)
    barrier()
    t1 = time.perf_counter()
    sampler.stop_flag = True
    sampler.join()
    dt = t1 - t0

    # end to end: host (pinned) buffers in, host results out, through the reference-facing entry point
    pin = [torch.from_numpy(a).pin_memory() for a in (seq_cat, off, ws, kind, gate)]
    pn = [p.numpy() for p in pin]
    h2d = sum(a.nbytes for a in pn) + 4 * len(kind)   # + the processing-order array built by the library
    d2h = 8 * (7 + 2 * NT)
    ctx.estep(*pn)
    barrier()
    t2 = time.perf_counter()
    e2e_kernel_ms = []
    for _ in range(args.steps):
        r = ctx.estep(*pn)
        e2e_kernel_ms.append(sum(t[1] for t in ctx.timing() if t[2] > 0))
        if world > 1:
            ctx.allreduce_sum(np.concatenate([[r.fn, r.sum_eff, float(r.n_skipped)], r.EN_diff, r.EH_diff]))
    barrier()
    dt_e2e = time.perf_counter() - t2

    # secondary figure of BASELINE.json's metric: `elem scan` sequences/s (posteriors on the linear-space kernels,
    # bit-exact Viterbi on the log-space kernel), on a bounded sample of the same sequences
    nscan = min(2 * npos, 2048)
    sb = ctx.batch(seq_cat[:nscan * SEQ_LEN], off[:nscan + 1], ws[:nscan * SEQ_LEN])
    ctx.scan_run(sb)
    barrier()
    t3 = time.perf_counter()
    ctx.scan_run(sb)
    barrier()
    dt_scan = time.perf_counter() - t3
    sb.close()

    if dist is not None:
        t = torch.tensor([dt, dt_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e = float(t[0]), float(t[1])
    if rank == 0:
        value = world * step_cells * args.steps / dt
        e2e = world * step_cells * args.steps / dt_e2e
        kms = float(np.mean(kernel_ms))
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, which = json.load(open(peaks_path))["hbm_gbs"], "measured"
        else:
            peak, which = 6650.0, "fallback"
        achieved = step_bytes / (kms * 1e-3) / 1e9
        traffic, issue = None, None
        tpath = os.path.join(ROOT, "profiles", "r1_dram_traffic.json")
        clocks = sampler.summary()
        if os.path.exists(tpath):
            prof = json.load(open(tpath))
            traffic = int(prof["dram_bytes_per_sequence_evaluation"] * 2 * npos)
            # the recursion is sparse gather work: the resource that binds is warp-instruction issue, not HBM or
            # the fp64 pipe.  instructions per sequence-evaluation come from the ncu pass of the same workload
            # (smsp__inst_executed.sum), the rate from this run's device time; peak = SMs x 4 schedulers x clock.
            sm = torch.cuda.get_device_properties(local).multi_processor_count
            mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
            ach = prof["warp_instructions_per_sequence_evaluation"] * 2 * npos / (kms * 1e-3)
            pk = sm * 4 * mhz * 1e6
            issue = {"achieved": ach, "peak": pk, "unit": "warp-instructions/s", "frac": ach / pk,
                     "source": "profiles/r1_dram_traffic.json (ncu smsp__inst_executed.sum per sequence-evaluation)"}
        line = {"metric": "dp_cells_per_s", "value": value, "unit": "dp_cells/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(npos, world),
                "seq_evals_per_s": world * 2 * npos * args.steps / dt,
                "e2e": {"value": e2e, "unit": "dp_cells/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": 1e3 * dt_e2e / args.steps, "kernel_ms_per_step": float(np.mean(e2e_kernel_ms))},
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": which,
                             "kernel_ms": kms, "algorithmic_bytes_per_launch": int(step_bytes),
                             "note": "one 'launch' = the whole wavefront of phase kernels over the step's batch"},
                "issue": issue,
                "scan": {"value": world * nscan / dt_scan, "unit": "seqs/s", "sample": "%d x %d nt per GPU" % (nscan, SEQ_LEN)}}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            line["cpu_baseline"] = cpu_baseline(args.ref_sample or max(16, 4 * cores))
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--nseq", type=int, default=int(os.environ.get("RELEM_BENCH_NSEQ", "10000")),
                    help="positives per GPU and step (configs[1]: 10000)")
    ap.add_argument("--ref-sample", type=int, default=0, help="positives in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    return main_own(args)


if __name__ == "__main__":
    sys.exit(main())
