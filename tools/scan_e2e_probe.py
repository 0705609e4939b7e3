"""host-buffer scan (relem_scan) against the resident form: where the time of a call goes (batch upload, kernels, release)"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, rnaelem_b200 as rb
from rnaelem_b200 import hostio
ctx = rb.Context(0)
with tempfile.NamedTemporaryFile("w", suffix=".model", delete=False) as f:
    f.write(bench.SCAN_MODEL)
ctx.set_model(hostio.read_model(f.name)); os.unlink(f.name)
n, L = 8192, 200
pos, _ = bench.make_dataset(n, 2000, L)
seq_cat = np.ascontiguousarray(pos.reshape(-1)); off = np.arange(0, (n + 1) * L, L, dtype=np.int64); ws = np.zeros(n * L)
batch = ctx.batch(seq_cat, off, ws)
for k in range(2):
    t0 = time.perf_counter(); ctx.scan_run(batch); t1 = time.perf_counter()
    print('resident %.0f ms' % ((t1 - t0) * 1e3), [(t[0][:24], round(t[1]), t[2]) for t in ctx.timing()], flush=True)
for k in range(3):
    t0 = time.perf_counter(); b2 = ctx.batch(seq_cat, off, ws); t1 = time.perf_counter(); ctx.scan_run(b2); t2 = time.perf_counter(); b2.close(); t3 = time.perf_counter()
    print('second batch: create %.0f run %.0f close %.0f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), [(t[0][:24], round(t[1]), t[2]) for t in ctx.timing()], flush=True)
for k in range(2):
    t0 = time.perf_counter(); ctx.scan(seq_cat, off, ws, decode_rss=False); t1 = time.perf_counter()
    print('host-buffer call %.0f ms' % ((t1 - t0) * 1e3), [(t[0][:24], round(t[1]), t[2]) for t in ctx.timing()], flush=True)
