"""config[4]-like probe: long sequences on the linear-space path: how many fall back to the log-space kernel, throughput"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rnaelem_b200 as rb
L = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 150
n = int(sys.argv[3]) if len(sys.argv) > 3 else 64
pattern = sys.argv[4] if len(sys.argv) > 4 else "((.*.))"
for lam in [(0.0, 0.0), (0.8, 1.1)]:
    ctx = rb.Context(0)
    ctx.set_energy("~A2007~", W, 30, 1e-4, 0); ctx.set_pattern(pattern)
    rows = ctx.row_sizes
    theta = np.concatenate([np.full(r, -np.log(r)) for r in rows])
    ctx.set_params(theta, list(lam), 0.1)
    rng = np.random.RandomState(4)
    seqs = [rng.randint(1, 5, size=L).astype(np.uint8) for _ in range(n)]
    kind = [rb.POS_WITH if k % 2 == 0 else rb.NEG for k in range(n)]
    gate = [-1 if k % 2 == 0 else k - 1 for k in range(n)]
    sc, off, wc = rb.pack_batch(seqs, [np.zeros(L)] * n)
    b = ctx.batch(sc, off, wc, np.array(kind, np.uint8), np.array(gate, np.int32))
    for k in range(2):
        t0 = time.perf_counter(); r = ctx.estep_run(b); t1 = time.perf_counter()
    print('L', L, 'W', W, 'lambda', lam, 'n', n, '%.1f ms' % ((t1 - t0) * 1e3), '%.1f seq-evals/s' % (n / (t1 - t0)), 'fn', r.fn,
          [(t[0], round(t[1], 1)) for t in ctx.timing() if t[2] > 0])
    sb = ctx.batch(sc[:(n // 4) * L], off[:n // 4 + 1], wc[:(n // 4) * L])
    for k in range(2):
        t0 = time.perf_counter(); ctx.scan_run(sb); t1 = time.perf_counter()
    print('   scan', n // 4, 'reads: %.1f ms' % ((t1 - t0) * 1e3), '%.1f reads/s' % (n // 4 / (t1 - t0)))
