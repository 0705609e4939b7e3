#!/usr/bin/env python3
"""E-step time per call for small minibatches (the `elem` default is 64 reads = 128 sequence-evaluations per objective
call).  python tools/minibatch_probe.py [pairs ...]; RELEM_FUSE_CELLS selects the fused-diagonal path."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import rnaelem_b200 as rb

ctx = rb.Context(0)
ctx.set_energy("~T2004~", 50, 30, 1e-4, 0)
ctx.set_pattern("((.*.))")
theta, lam, tau = bench.uniform_model()
ctx.set_params(theta, lam, tau)
sizes = [int(x) for x in sys.argv[1:]] or [32, 64, 128, 256, 512, 1024, 2048]
for npos in sizes:
    pos, neg = bench.make_dataset(npos, 1000, 200)
    seq_cat, off, ws, kind, gate = bench.pack(pos, neg)
    batch = ctx.batch(seq_cat, off, ws, kind, gate)
    ts = []
    for k in range(7):
        t0 = time.perf_counter()
        r = ctx.estep_run(batch)
        ts.append((time.perf_counter() - t0) * 1e3)
    if os.environ.get("RELEM_PHASE_TIMING"):
        for t in ctx.timing():
            print("    %-80s %8.3f ms %d" % (t[0], t[1], t[2]))
    ts2 = []
    for k in range(5):
        t0 = time.perf_counter()
        r = ctx.estep(seq_cat, off, ws, kind, gate)
        ts2.append((time.perf_counter() - t0) * 1e3)
    fn = (r.fn, float(np.abs(r.EN_diff).sum()))
    print("seq-evals %5d  resident %.2f ms  host-buffers %.2f ms  fuse=%s  fn,|EN_diff|=%r" %
          (2 * npos, float(np.median(ts[2:])), float(np.median(ts2[1:])), os.environ.get("RELEM_FUSE_CELLS", "default"), fn),
          flush=True)
