"""Multi-rank E-step: sequences are independent, so every rank evaluates a contiguous block of
training examples (ArrayJobManager::assigned_range, arrayjob_manager.hpp:141-149, through relem_assigned_range) and
the P+3 doubles [fn, sum_eff, n_skipped, EN_diff, EH_diff] are summed once before the host optimizer step -- the
reference does the same sum through text files (motif_array_trainer.hpp:20-61).  The collective is injected: NCCL
over NVLink on the GPU box (Context.allreduce_sum), gloo in the CPU tests."""
import ctypes

import numpy as np


def assigned_range(lib, total, n, k):
    a, b = ctypes.c_int64(), ctypes.c_int64()
    lib.relem_assigned_range(int(total), int(n), int(k), ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def shard_examples(lib, seqs, wss, kind, gate, rank, world):
    """examples = maximal runs [positive, negatives gated on it...]; returns this rank's (seqs, wss, kind, gate)
    with gate indices renumbered to the shard."""
    starts = [i for i in range(len(seqs)) if gate[i] < 0]
    lo, hi = assigned_range(lib, len(starts), world, rank)
    if lo == hi:
        return [], [], [], []
    a = starts[lo]
    b = starts[hi] if hi < len(starts) else len(seqs)
    g = [(-1 if gate[i] < 0 else gate[i] - a) for i in range(a, b)]
    return seqs[a:b], wss[a:b], list(kind[a:b]), g


def pack_result(r):
    return np.concatenate([[r.fn, r.sum_eff, float(r.n_skipped)], np.asarray(r.EN_diff), np.asarray(r.EH_diff)])


def unpack_result(v, n_theta):
    class R(object):
        pass
    r = R()
    r.fn, r.sum_eff, r.n_skipped = float(v[0]), float(v[1]), int(round(v[2]))
    r.EN_diff = np.array(v[3:3 + n_theta])
    r.EH_diff = np.array(v[3 + n_theta:5 + n_theta])
    return r


def sharded_estep(ctx, seqs, wss, kind, gate, rank, world, allreduce):
    """E-step of the whole batch evaluated by `world` ranks; every rank returns the global result."""
    from .hostio import pack_batch
    s, w, k, g = shard_examples(ctx.lib, seqs, wss, kind, gate, rank, world)
    if s:
        sc, off, wc = pack_batch(s, w)
        r = ctx.estep(sc, off, wc, np.asarray(k, np.uint8), np.asarray(g, np.int32))
        v = pack_result(r)
    else:
        v = np.zeros(5 + ctx.n_theta)
    return unpack_result(allreduce(v), ctx.n_theta)
