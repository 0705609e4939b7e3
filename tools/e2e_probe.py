import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, rnaelem_b200 as rb
ctx = rb.Context(0)
ctx.set_energy("~T2004~", 50, 30, 1e-4, 0); ctx.set_pattern("((.*.))")
theta, lam, tau = bench.uniform_model(); ctx.set_params(theta, lam, tau)
npos = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
pos, neg = bench.make_dataset(npos, 1000, 200)
seq_cat, off, ws, kind, gate = bench.pack(pos, neg)
batch = ctx.batch(seq_cat, off, ws, kind, gate)
for k in range(3):
    t0 = time.perf_counter(); r = ctx.estep_run(batch); t1 = time.perf_counter()
    print('resident %.1f ms' % ((t1 - t0) * 1e3), [(t[0], round(t[1], 1)) for t in ctx.timing() if t[2] > 0])
for k in range(3):
    t0 = time.perf_counter(); r = ctx.estep(seq_cat, off, ws, kind, gate); t1 = time.perf_counter()
    print('e2e %.1f ms' % ((t1 - t0) * 1e3), [(t[0], round(t[1], 1)) for t in ctx.timing() if t[2] > 0])
