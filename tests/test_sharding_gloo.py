"""N > 1 path on the CPU: two gloo ranks shard a batch with the reference's block formula, each evaluates its
shard (through the single-threaded emulation of the kernel source -- there is no GPU here), one all-reduce of the
P+3 doubles, and every rank must hold the single-rank result.  On the GPU box the same host code runs with NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import caselib
import rnaelem_b200 as rb
from rnaelem_b200 import sharding

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, emu, name, out):
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = caselib.load_case(name)
        ctx = caselib.make_ctx(case, lib=emu)
        seqs, wss, kind, gate, _ = caselib.estep_inputs(case)

        def allreduce(v):
            t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64))
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return t.numpy()

        r = sharding.sharded_estep(ctx, seqs, wss, kind, gate, rank, world, allreduce)
        np.save(os.path.join(out, "r%d.npy" % rank), sharding.pack_result(r))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["m0", "ragged"])
def test_two_rank_estep_equals_single_rank(name, emu_lib, tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, emu_lib, name, str(tmp_path)), nprocs=world, join=True)
    case = caselib.load_case(name)
    ctx = caselib.make_ctx(case, lib=emu_lib)
    seqs, wss, kind, gate, _ = caselib.estep_inputs(case)
    sc, off, wc = rb.pack_batch(seqs, wss)
    whole = sharding.pack_result(ctx.estep(sc, off, wc, np.asarray(kind, np.uint8), np.asarray(gate, np.int32)))
    est = case["estep"]
    assert caselib.close(whole[0], est["fn"])
    for k in range(world):
        got = np.load(os.path.join(str(tmp_path), "r%d.npy" % k))
        np.testing.assert_allclose(got, whole, rtol=1e-12, atol=1e-12)


def test_shards_partition_the_examples():
    lib = rb.load_library()
    seqs = list(range(11))
    gate = [-1, 0, -1, 2, -1, 4, -1, -1, 7, -1, 9]   # 6 examples, some without a negative
    kind = [0] * 11
    for world in (1, 2, 3, 4, 8):
        seen = []
        for rank in range(world):
            s, _, _, g = sharding.shard_examples(lib, seqs, seqs, kind, gate, rank, world)
            seen += s
            for k, gg in enumerate(g):
                assert gg == -1 or (0 <= gg < k)
        assert seen == seqs
