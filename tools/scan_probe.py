"""scan throughput probe: relem_scan_run over N synthetic 200-nt sequences (BASELINE configs[2] shape)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, rnaelem_b200 as rb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = rb.Context(0)
ctx.set_energy("~T2004~", 50, 30, 1e-4, 0); ctx.set_pattern("((.*.))")
theta, lam, tau = bench.uniform_model(); ctx.set_params(theta, [0.3, 0.6], tau)
rng = np.random.RandomState(2)
seqs = [rng.randint(1, 5, size=200).astype(np.uint8) for _ in range(n)]
sc, off, wc = rb.pack_batch(seqs, [np.zeros(200)] * n)
b = ctx.batch(sc, off, wc)
for k in range(2):
    t0 = time.perf_counter(); r = ctx.scan_run(b); t1 = time.perf_counter()
    print('scan %d seqs: %.1f ms -> %.1f seqs/s' % (n, (t1 - t0) * 1e3, n / (t1 - t0)), ctx.timing()[:2])
