"""scan throughput probe: relem_scan_run over N synthetic 200-nt reads (BASELINE configs[2] shape) with the bench's
scan model; prints the device time of the posterior passes and of the Viterbi kernel"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
import rnaelem_b200 as rb
from rnaelem_b200 import hostio

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ctx = rb.Context(0)
with tempfile.NamedTemporaryFile("w", suffix=".model", delete=False) as f:
    f.write(bench.SCAN_MODEL)
ctx.set_model(hostio.read_model(f.name))
os.unlink(f.name)
rng = np.random.RandomState(2)
seqs = [rng.randint(1, 5, size=L).astype(np.uint8) for _ in range(n)]
sc, off, wc = rb.pack_batch(seqs, [np.zeros(L)] * n)
b = ctx.batch(sc, off, wc)
for k in range(3):
    t0 = time.perf_counter(); r = ctx.scan_run(b); t1 = time.perf_counter()
    print('scan %d reads: %.1f ms -> %.1f reads/s' % (n, (t1 - t0) * 1e3, n / (t1 - t0)),
          [(t[0], round(t[1], 1), t[2]) for t in ctx.timing()])
