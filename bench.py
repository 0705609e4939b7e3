#!/usr/bin/env python3
"""bench.py -- throughput of the RNAelem hot path on B200 against the unmodified reference on the host cores.

Workloads (BASELINE.json configs; --workload):
  estep (default, configs[1])  NSEQ x 200-nt i.i.d. ACGU positives (flagged "contains motif") + one shuffled negative
        each per GPU, pattern ((.*.)), max-span 50, Turner2004, min-bpp 1e-4, iteration-0 parameters.  One step = one
        objective evaluation of RNAelemTrainer::operator() (motif_trainer.hpp:595-633): base-pair filter + coupled
        inside + outside with expected counts for every sequence, batch reduction, at N > 1 the NCCL all-reduce of the
        P+3 sums (checked against a torch.distributed gather of the per-rank vectors: `allreduce_checked`).
  scan  (configs[2])  NSEQ x 200-nt reads per GPU through RNAelemScanner::scan's seam (motif_scanner.hpp:938-949) with a
        trained-model-like parameter set: posteriors, Ys / Ye, exist prob, Viterbi psihat / rss.  No collective.
  long  (configs[4])  NSEQ x 1000-nt positives + negatives per GPU, max-span 150, Andronescu2007.

metric  estep / long: dp_cells_per_s = band cells x 3 coupled passes per second (REFERENCE-EQUIVALENT cells: the
        reference runs inside + two outside passes per sequence; the device fuses the two outside passes
        into one, `device_passes` = 2).  scan: scan_seqs_per_s.
        `value`: inputs resident in HBM.  `e2e`: the host-buffer entry point (relem_estep / relem_scan) with pinned
        host inputs, host<->device copies inside the timed region.
roofline  whole wavefront of one step: algorithmic bytes (112 x S B/cell for the E-step with fused
        outside passes, 280 x S for scan) / summed device time of the step's kernels (CUDA events on the launch
        stream) against MEASURED_PEAKS.json:hbm_gbs.  `top_kernel` = the largest phase kernel with its own device
        time, from one extra instrumented step (RELEM_PHASE_TIMING=1, events around every phase launch, one lane).
        `traffic`, `issue`, `fp64.achieved` use per-sequence counters of a committed ncu pass (`source` names the file);
        `fp64.peak` is measured in this run (relem_fp64_peak).
cpu_baseline / --impl reference  the unmodified reference binary (oracle/_ref/RNAelem train|scan) with -t <all host
        cores> on a bounded sample of the same workload (>= 32 sequence-evaluations per thread for the 200-nt
        workloads; the sample is stated in the line).
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

PASSES = 3            # reference-equivalent coupled passes per sequence-evaluation (inside + 2 outside)
DEVICE_PASSES = 2     # what the device runs: inside + one outside pass carrying the difference of the two conditions
BYTES_ESTEP = 112     # x S per band cell: write inside once, read it in the fused outside pass
BYTES_SCAN = 280      # x S per band cell: inside-w, outside-r, insideEnd-w, outsideEnd-r, cyk-w

WORKLOADS = {
    "estep": dict(config="configs[1]", L=200, W=50, param="~T2004~", pattern="((.*.))", S=22, nseq=10000,
                  metric="dp_cells_per_s", unit="dp_cells/s"),
    "scan": dict(config="configs[2]", L=200, W=50, param="~T2004~", pattern="((.*.))", S=22, nseq=100000,
                 metric="scan_seqs_per_s", unit="seqs/s"),
    "long": dict(config="configs[4]", L=1000, W=150, param="~A2007~", pattern="((.*.))", S=22, nseq=128,
                 metric="dp_cells_per_s", unit="dp_cells/s"),
}

# parameters of a trained-model-like ((.*.)) motif (the model of tests/golden case synth200) for the scan workload
SCAN_MODEL = """pattern: ((.*.))
theta: [[-1.38629436111989,-1.38629436111989,-1.38629436111989,-1.38629436111989],[-1.2,-1.5,-1.3,-1.6],[-1.1,-1.7,-1.4,-1.45],[-1.9,-1.7,-1.6,-1.8,-1.75,-2],[-1.5,-1.9,-1.6,-1.85,-1.7,-2.2]]
ene-param: ~T2004~
max-span: 50
max-internal-loop: 30
rho-theta: 0.1
rho-lambda: 0.1
tau: 0.1
lambda: [0.3,0.6]
min-bpp: 0.0001
theta-softmax: 0
"""


def cells(L, W):
    W = min(L, W)
    return (L + 1) * (W + 1) - W * (W + 1) // 2


def make_dataset(nseq, seed, L):
    rng = np.random.RandomState(seed)
    pos = rng.randint(1, 5, size=(nseq, L)).astype(np.uint8)
    # negatives: seeded shuffle of each positive (composition preserving; the inputs are i.i.d. so a
    # dinucleotide-preserving shuffle has the same distribution)
    neg = np.stack([p[rng.permutation(L)] for p in pos])
    return pos, neg


def pack(pos, neg):
    import rnaelem_b200 as rb
    n, L = pos.shape
    seq = np.empty((2 * n, L), np.uint8)
    seq[0::2] = pos
    seq[1::2] = neg
    kind = np.empty(2 * n, np.uint8)
    kind[0::2] = rb.POS_WITH
    kind[1::2] = rb.NEG
    gate = np.full(2 * n, -1, np.int32)
    gate[1::2] = np.arange(0, 2 * n, 2)
    off = np.arange(0, (2 * n + 1) * L, L, dtype=np.int64)
    ws = np.zeros(2 * n * L)  # flat quality '+' x L -> ws = ln(1) = 0
    return np.ascontiguousarray(seq.reshape(-1)), off, ws, kind, gate


def uniform_theta(rows):
    return np.concatenate([np.full(r, -math.log(r)) for r in rows])


def uniform_model():
    """iteration-0 parameters of ((.*.)) (kept for tools/*.py)"""
    return uniform_theta([4, 4, 4, 6, 6]), [0.0, 0.0], 0.1


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


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    return 6650.0, "fallback of B200_PROFILING.md"


def static_profile():
    """per-sequence-evaluation counters of the committed ncu pass of this build (newest first)"""
    for name in ("r2_estep_counters.json", "r1_dram_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return json.load(open(p)), "profiles/" + name
    return None, None


# ------------------------------------------------------------------------------------------- reference arm
def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "RNAelem")
    return p if os.path.exists(p) else None


def write_fastq(path, seqs):
    L = seqs.shape[1]
    with open(path, "w") as f:
        for k in range(len(seqs)):
            f.write("@s%d\n%s\n+\n%s!\n" % (k, "".join("NACGU"[c] for c in seqs[k]), "+" * L))


def run_reference_train(wl, nseq, seed, threads):
    """one objective evaluation of the unmodified reference over nseq positives (+ its own shuffled negatives);
    returns (seconds per evaluation as the binary prints it, sequence evaluations)"""
    pos, _ = make_dataset(nseq, seed, wl["L"])
    with tempfile.TemporaryDirectory() as d:
        fq = os.path.join(d, "x.fq")
        write_fastq(fq, pos)
        cmd = [reference_binary(), "train", "-f", fq, "-m", wl["pattern"], "-w", str(wl["W"]), "--energy-param",
               wl["param"], "-t", str(threads), "--batch-size", "-1", "--max-iter", "1", "--out1", "/dev/null",
               "--out2", "/dev/null", "--out3", "/dev/null"]
        p = subprocess.run(cmd, capture_output=True, text=True)
    m = re.search(r"wall clock time per eval: ([0-9.eE+-]+)", p.stderr + p.stdout)
    if p.returncode != 0 or not m:
        raise RuntimeError("reference run failed: " + (p.stderr[-500:]))
    return float(m.group(1)), 2 * nseq


def run_reference_scan(wl, nseq, seed, threads):
    """RNAelem scan of nseq reads with the scan model; returns (seconds of `scan end:`, reads)"""
    pos, _ = make_dataset(nseq, seed, wl["L"])
    with tempfile.TemporaryDirectory() as d:
        fq = os.path.join(d, "x.fq")
        mp = os.path.join(d, "scan.model")
        write_fastq(fq, pos)
        open(mp, "w").write(SCAN_MODEL)
        cmd = [reference_binary(), "scan", "-f", fq, "-q", mp, "-t", str(threads), "--out1", "/dev/null"]
        p = subprocess.run(cmd, capture_output=True, text=True)
    m = re.search(r"scan end: ([0-9.eE+-]+)", p.stderr + p.stdout)
    if p.returncode != 0 or not m:
        raise RuntimeError("reference scan failed: " + (p.stderr[-500:]))
    return float(m.group(1)), nseq


def ref_sample_size(wl_name, cores, override):
    if override:
        return override
    if wl_name == "long":
        return cores          # 2 sequence-evaluations per thread at ~17 s each: more does not fit a bench run
    if wl_name == "scan":
        return 32 * cores     # 32 reads per thread
    return 16 * cores         # 16 positives + 16 negatives = 32 sequence-evaluations per thread


def reference_measure(wl_name, wl, sample, seed, cores):
    """-> (value in the workload's unit, sequence units per second, description)"""
    if wl_name == "scan":
        sec, n = run_reference_scan(wl, sample, seed, cores)
        return n / sec, n / sec, ("%d reads x %d nt, RNAelem scan -t %d (`scan end:` seconds), %d reads per thread"
                                  % (sample, wl["L"], cores, sample // cores))
    sec, evals = run_reference_train(wl, sample, seed, cores)
    v = evals * cells(wl["L"], wl["W"]) * PASSES / sec
    return v, evals / sec, ("%d positives + %d in-binary shuffled negatives x %d nt, RNAelem train -t %d --batch-size -1 "
                            "--max-iter 1 (`wall clock time per eval`), %d sequence-evaluations per thread"
                            % (sample, sample, wl["L"], cores, 2 * sample // cores))


def cpu_baseline(wl_name, wl, override):
    cores = os.cpu_count() or 1
    if reference_binary() is None:
        return {"value": None, "unit": wl["unit"], "cores": cores, "kind": "reference",
                "sample": "oracle/_ref/RNAelem missing"}
    sample = ref_sample_size(wl_name, cores, override)
    v, ups, desc = reference_measure(wl_name, wl, sample, 12345, cores)
    return {"value": v, "unit": wl["unit"], "cores": cores, "kind": "reference", "seq_units_per_s": ups, "sample": desc}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = dict(WORKLOADS[args.workload])
    cores = os.cpu_count() or 1
    if reference_binary() is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/RNAelem was not built"}))
        return 0
    sample = ref_sample_size(args.workload, cores, args.ref_sample)
    for _ in range(min(args.warmup, 1)):   # one short warm-up run pages the binary and the tables in
        reference_measure(args.workload, wl, max(2, cores // 2) if args.workload == "long" else max(8, cores), 999, cores)
    vals, t0 = [], time.perf_counter()
    desc = ""
    for k in range(args.steps):
        v, ups, desc = reference_measure(args.workload, wl, sample, 12345 + k, cores)
        vals.append((v, ups))
    wall = time.perf_counter() - t0
    # equal-sized steps: the mean rate over the steps is total work / total time
    v = len(vals) / sum(1. / x[0] for x in vals)
    ups = len(vals) / sum(1. / x[1] for x in vals)
    line = {"impl": "reference", "metric": wl["metric"], "value": v, "unit": wl["unit"], "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, wl, sample, 1),
            "same_config": False,
            "same_config_note": "same workload, parameters and code path as the GPU arm, on a bounded subsample of its "
                                "sequences (sequences are independent, throughput does not depend on the batch size "
                                "once every thread has tens of them): " + desc,
            "cpu_baseline": {"value": v, "unit": wl["unit"], "cores": cores, "kind": "reference", "seq_units_per_s": ups,
                             "sample": desc},
            "e2e": {"value": v, "unit": wl["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(name, wl, nseq, ngpu):
    if name == "scan":
        what = "%d x %d-nt synthetic reads per GPU" % (nseq, wl["L"])
        par = "dp%d (reads sharded contiguously, no collective)" % ngpu
    else:
        what = "%d x %d-nt synthetic positives + %d shuffled negatives per GPU" % (nseq, wl["L"], nseq)
        par = "dp%d (sequences sharded, one all-reduce of fn/gr per step)" % ngpu
    return {"workload": "RNAelem %s (%s): %s, pattern %s, max-span %d, %s, min-bpp 1e-4"
                        % (name, wl["config"], what, wl["pattern"], wl["W"], wl["param"]),
            "seqs_per_gpu": nseq, "seq_len": wl["L"], "max_span": wl["W"], "pattern": wl["pattern"],
            "states": wl["S"], "parallelism": par,
            "cache": "working set (DP tables of all resident sequences, > 1 GB) exceeds the 126 MB L2"}


# -------------------------------------------------------------------------------------------------- own arm
class Rig(object):
    """context + torch.distributed plumbing shared by the workloads"""

    def __init__(self):
        import torch
        import rnaelem_b200 as rb
        self.torch, self.rb = torch, rb
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (librelem has no CPU path)")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.ctx = rb.Context(self.local)

    def comm(self):
        if self.world == 1:
            return
        import ctypes
        torch, dist = self.torch, self.dist
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            buf = (ctypes.c_uint8 * 128)()
            assert self.ctx.lib.relem_comm_unique_id(buf) == 0, "ncclGetUniqueId failed"
            uid = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        self.ctx.comm_init(bytes(uid.cpu().tolist()), self.rank, self.world)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        if self.dist is None:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def pin(self, arrays):
        return [self.torch.from_numpy(a).pin_memory().numpy() for a in arrays]

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def result_vector(r):
    return np.concatenate([[r.fn, r.sum_eff, float(r.n_skipped)], r.EN_diff, r.EH_diff])


def check_allreduce(rig, mine, reduced):
    """every rank: the NCCL all-reduced vector must equal the sum of the per-rank vectors gathered through
    torch.distributed (summed in rank order in fp64)"""
    torch, dist = rig.torch, rig.dist
    t = torch.from_numpy(np.ascontiguousarray(mine)).cuda()
    parts = [torch.empty_like(t) for _ in range(rig.world)]
    dist.all_gather(parts, t)
    want = np.sum(np.stack([p.cpu().numpy() for p in parts]), axis=0)
    scale = np.maximum(1.0, np.maximum(np.abs(want), np.max(np.abs(np.stack([p.cpu().numpy() for p in parts])), axis=0)))
    bad = np.abs(reduced - want) > 1e-12 * scale
    ok = torch.tensor([0 if bad.any() else 1], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok) != 1:
        raise SystemExit("rank %d: all-reduced sums differ from the gathered per-rank sums: %r vs %r"
                         % (rig.rank, reduced[bad], want[bad]))
    return True


def measure_estep(rig, wl_name, wl, nseq, steps, warmup, want_profile=True):
    rb, ctx, torch = rig.rb, rig.ctx, rig.torch
    ctx.set_energy(wl["param"], wl["W"], 30, 1e-4, 0)
    ctx.set_pattern(wl["pattern"])
    ctx.set_params(uniform_theta(ctx.row_sizes), [0.0, 0.0], 0.1)
    NT = ctx.n_theta
    pos, neg = make_dataset(nseq, 1000 + rig.rank, wl["L"])
    seq_cat, off, ws, kind, gate = pack(pos, neg)
    batch = ctx.batch(seq_cat, off, ws, kind, gate)
    step_cells = batch.cells * PASSES
    step_bytes = batch.cells * BYTES_ESTEP * ctx.S
    checked = None

    def step_resident(check=False):
        r = ctx.estep_run(batch)
        if rig.world > 1:
            mine = result_vector(r)
            red = ctx.allreduce_sum(mine.copy())
            if check:
                return check_allreduce(rig, mine, red)
        return None

    for k in range(warmup):
        c = step_resident(check=(k == 0))
        checked = c if c is not None else checked
    rig.barrier()
    sampler = ClockSampler(rig.local)
    sampler.start()
    kernel_ms, launches = [], 0
    t0 = time.perf_counter()
    for _ in range(steps):
        step_resident()
        tm = ctx.timing()
        kernel_ms.append(sum(t[1] for t in tm if t[2] > 0))
        launches += sum(t[2] for t in tm if t[2] > 0)
    rig.barrier()
    dt = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join()

    # end to end: host (pinned) buffers in, host results out, through the reference-facing entry point
    pn = rig.pin((seq_cat, off, ws, kind, gate))
    h2d = sum(a.nbytes for a in pn) + 4 * len(kind)   # + the processing-order array built by the library
    d2h = 8 * (7 + 2 * NT)
    ctx.estep(*pn)
    rig.barrier()
    t2 = time.perf_counter()
    e2e_kernel_ms = []
    for _ in range(steps):
        r = ctx.estep(*pn)
        e2e_kernel_ms.append(sum(t[1] for t in ctx.timing() if t[2] > 0))
        if rig.world > 1:
            ctx.allreduce_sum(result_vector(r))
    rig.barrier()
    dt_e2e = time.perf_counter() - t2

    # one instrumented step: device time per phase kernel class (events around every launch, one lane)
    top = None
    if want_profile and rig.rank == 0:
        os.environ["RELEM_PHASE_TIMING"] = "1"
        try:
            ctx.estep_run(batch)
            ph = [(t[0], t[1], -t[2]) for t in ctx.timing() if t[2] < 0]
            tot = sum(t[1] for t in ctx.timing() if t[2] > 0 and t[0] == "relem_estep_lin_kernel")
        finally:
            del os.environ["RELEM_PHASE_TIMING"]
        if ph:
            name, ms, nl = max(ph, key=lambda t: t[1])
            top = {"name": name, "ms_per_step": ms, "launches_per_step": nl, "avg_launch_us": 1e3 * ms / max(1, nl),
                   "share_of_step": ms / tot if tot else None,
                   "phases": {n: round(m, 3) for n, m, _ in sorted(ph, key=lambda t: -t[1])},
                   "how": "RELEM_PHASE_TIMING=1: CUDA events around every phase launch on its stream, one extra step "
                          "outside the timed region (single lane, so the sum runs ~3 % above the two-lane step)"}
    rig.barrier()
    dt, dt_e2e = rig.max_over_ranks([dt, dt_e2e])
    batch.close()
    return dict(dt=dt, dt_e2e=dt_e2e, step_cells=step_cells, step_bytes=step_bytes, kernel_ms=float(np.mean(kernel_ms)),
                e2e_kernel_ms=float(np.mean(e2e_kernel_ms)), launches=int(launches), h2d=int(h2d), d2h=int(d2h),
                clocks=sampler.summary(), top=top, checked=checked, evals=2 * nseq, S=ctx.S)


def measure_scan(rig, wl, nseq, steps, warmup):
    rb, ctx = rig.rb, rig.ctx
    from rnaelem_b200 import hostio
    with tempfile.NamedTemporaryFile("w", suffix=".model", delete=False) as f:
        f.write(SCAN_MODEL)
        mp = f.name
    try:
        ctx.set_model(hostio.read_model(mp))
    finally:
        os.unlink(mp)
    L = wl["L"]
    pos, _ = make_dataset(nseq, 2000 + rig.rank, L)
    seq_cat = np.ascontiguousarray(pos.reshape(-1))
    off = np.arange(0, (nseq + 1) * L, L, dtype=np.int64)
    ws = np.zeros(nseq * L)
    batch = ctx.batch(seq_cat, off, ws)
    step_bytes = batch.cells * BYTES_SCAN * ctx.S
    for _ in range(warmup):
        ctx.scan_run(batch)
    rig.barrier()
    sampler = ClockSampler(rig.local)
    sampler.start()
    kernel_ms, launches = [], 0
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.scan_run(batch)
        tm = ctx.timing()
        kernel_ms.append(sum(t[1] for t in tm if t[2] > 0))
        launches += sum(t[2] for t in tm if t[2] > 0)
    rig.barrier()
    dt = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join()
    pn = rig.pin((seq_cat, off, ws))
    h2d = sum(a.nbytes for a in pn)
    tl = nseq * L
    d2h = 8 * tl * 3 + 8 * nseq + 4 * tl + tl + 4 * 2 * nseq + 8 * nseq * 2 + 8 * ctx.n_theta * nseq
    ctx.scan(*pn, decode_rss=False)
    rig.barrier()
    t2 = time.perf_counter()
    for _ in range(steps):
        ctx.scan(*pn, decode_rss=False)
    rig.barrier()
    dt_e2e = time.perf_counter() - t2
    dt, dt_e2e = rig.max_over_ranks([dt, dt_e2e])
    batch.close()
    return dict(dt=dt, dt_e2e=dt_e2e, step_bytes=step_bytes, kernel_ms=float(np.mean(kernel_ms)), launches=int(launches),
                h2d=int(h2d), d2h=int(d2h), clocks=sampler.summary(), S=ctx.S)


def roofline_block(step_bytes, kernel_ms, what):
    peak, which = hbm_peak()
    achieved = step_bytes / (kernel_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": what, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "peak_source": which, "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": int(step_bytes),
            "note": "one 'launch' = the whole wavefront of phase kernels over the step's batch (kernel_ms = their summed "
                    "device time, CUDA events on the launch streams)"}


def main_own(args):
    rig = Rig()
    rig.comm()
    name = args.workload
    wl = dict(WORKLOADS[name])
    nseq = args.nseq or wl["nseq"]
    world, rank = rig.world, rig.rank
    line = None
    if name in ("estep", "long"):
        m = measure_estep(rig, name, wl, nseq, args.steps, args.warmup)
        scan = None
        if name == "estep" and not args.no_scan:
            # the other half of BASELINE.json's metric on a bounded batch of the same shape
            ns = min(nseq, 8192)
            sm = measure_scan(rig, WORKLOADS["scan"], ns, 1, 1)
            scan = {"value": world * ns / sm["dt"], "unit": "seqs/s",
                    "e2e": {"value": world * ns / sm["dt_e2e"], "unit": "seqs/s", "h2d_bytes_per_step": sm["h2d"],
                            "d2h_bytes_per_step": sm["d2h"]},
                    "roofline": roofline_block(sm["step_bytes"], sm["kernel_ms"], "scan wavefront: filter + 2 x (inside, outside) + Viterbi"),
                    "sample": "%d x %d nt per GPU, 1 step after 1 warm-up (python bench.py --workload scan runs configs[2]'s "
                              "per-GPU share)" % (ns, wl["L"])}
        if rank == 0:
            value = world * m["step_cells"] * args.steps / m["dt"]
            e2e = world * m["step_cells"] * args.steps / m["dt_e2e"]
            roof = roofline_block(m["step_bytes"], m["kernel_ms"],
                                  "E-step wavefront: relem_lin_phase_kernel<0..9,13> + exterior-row, prep, filter, fold kernels")
            roof["top_kernel"] = m["top"]
            prof, src = static_profile()
            issue = fp64 = None
            dfma, fexp = rig.ctx.fp64_peak()
            if prof and name == "estep":
                per = m["evals"] / (m["kernel_ms"] * 1e-3)
                roof["traffic"] = int(prof["dram_bytes_per_sequence_evaluation"] * m["evals"])
                roof["traffic_source"] = "static profile: " + src
                sm_n = rig.torch.cuda.get_device_properties(rig.local).multi_processor_count
                mhz = m["clocks"].get("sm_mhz") or m["clocks"].get("sm_max_mhz") or 1965.0
                ach = prof["warp_instructions_per_sequence_evaluation"] * per
                pk = sm_n * 4 * mhz * 1e6
                issue = {"achieved": ach, "peak": pk, "unit": "warp-instructions/s", "frac": ach / pk,
                         "source": "static profile: %s (smsp__inst_executed.sum per sequence-evaluation) x this run's rate" % src}
                if "dfma_per_sequence_evaluation" in prof:
                    a = prof["dfma_per_sequence_evaluation"] * per
                    fp64 = {"achieved": a, "peak": dfma, "unit": "DFMA/s", "frac": a / dfma,
                            "source": "static profile: %s (fp64 fma+mul+add thread instructions) x this run's rate" % src}
            if fp64 is None:
                fp64 = {"achieved": None, "peak": dfma, "unit": "DFMA/s"}
            fp64["exp_peak_per_s"] = fexp
            fp64["peak_source"] = "relem_fp64_peak micro-benchmark in this run (8 FMA chains / 4 exp chains per thread)"
            line = {"metric": wl["metric"], "value": value, "unit": wl["unit"], "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": 1e3 * m["dt"] / args.steps, "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": workload_config(name, wl, nseq, world),
                    "cells_counted": "reference-equivalent: band cells x 3 coupled passes (inside + 2 outside, SURVEY.md "
                                     "8d); the device fuses the two outside passes (device_passes = 2)",
                    "device_passes": DEVICE_PASSES,
                    "device_cells_per_s": value * DEVICE_PASSES / PASSES,
                    "seq_evals_per_s": world * m["evals"] * args.steps / m["dt"],
                    "state_cells_per_s": value * 7 * wl["S"],   # cells x 7 state types x S automaton states
                    "e2e": {"value": e2e, "unit": wl["unit"], "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                            "ms_per_step": 1e3 * m["dt_e2e"] / args.steps, "kernel_ms_per_step": m["e2e_kernel_ms"]},
                    "gpu_launches": m["launches"], "clocks": m["clocks"], "roofline": roof, "issue": issue, "fp64": fp64}
            if world > 1:
                line["allreduce_checked"] = bool(m["checked"])
            if scan is not None:
                line["scan"] = scan
    else:
        sm = measure_scan(rig, wl, nseq, args.steps, args.warmup)
        if rank == 0:
            line = {"metric": wl["metric"], "value": world * nseq * args.steps / sm["dt"], "unit": wl["unit"],
                    "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": 1e3 * sm["dt"] / args.steps, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": workload_config(name, wl, nseq, world),
                    "dp_cells_per_s": world * nseq * cells(wl["L"], wl["W"]) * 5 * args.steps / sm["dt"],
                    "cells_counted": "scan = 5 coupled passes per read in the reference (SURVEY.md 8d)",
                    "e2e": {"value": world * nseq * args.steps / sm["dt_e2e"], "unit": wl["unit"],
                            "h2d_bytes_per_step": sm["h2d"], "d2h_bytes_per_step": sm["d2h"],
                            "ms_per_step": 1e3 * sm["dt_e2e"] / args.steps},
                    "gpu_launches": sm["launches"], "clocks": sm["clocks"],
                    "roofline": roofline_block(sm["step_bytes"], sm["kernel_ms"],
                                               "scan wavefront: filter + 2 x (inside, outside) phase kernels + relem_viterbi_kernel")}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(name, wl, args.ref_sample)
            if name == "estep" and "scan" in line:
                line["scan"]["cpu_baseline"] = cpu_baseline("scan", WORKLOADS["scan"], 0)
        print(json.dumps(line))
    rig.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="estep", choices=sorted(WORKLOADS))
    ap.add_argument("--nseq", type=int, default=int(os.environ.get("RELEM_BENCH_NSEQ", "0")),
                    help="sequences (positives for estep / long, reads for scan) per GPU and step; 0 = the workload's size")
    ap.add_argument("--ref-sample", type=int, default=0, help="sequences in the bounded CPU sample (0 = 16-32 per thread)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scan", action="store_true", help="estep workload: skip the bounded scan figure")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    return main_own(args)


if __name__ == "__main__":
    sys.exit(main())
