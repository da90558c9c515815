"""N>1 host logic on CPU: two gloo ranks each run the (oracle) hot path on their locus shard and gather
their per-locus records on rank 0; the result must equal one process over the whole catalog."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Res:
    """HotPathResult-shaped view of an oracle pass (just what record_parts reads)."""

    def __init__(self, ref):
        spans, glue, (offs, words, scores), (mc_off, mc, span_off, hspans, purity, status) = ref
        self.glue = glue

        class A:
            pass
        self.annotations = A()
        self.annotations.motif_counts, self.annotations.spans, self.annotations.purity = mc, hspans, purity


def _payload(n_loci, begin, count):
    from oracle import oracle as orc
    from harness import workload
    from harness.pipeline import oracle_pass
    from trgt_b200.shard import record_parts
    w = workload.generate(count, 8, locus_begin=begin, seed=31337)
    return np.concatenate(record_parts([_Res(oracle_pass(orc, w, 2))]))


def _worker(rank, world, port, n_loci, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from trgt_b200.shard import RecordGather, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = shard_bounds(n_loci, world)
    mine = _payload(n_loci, b[rank], b[rank + 1] - b[rank])
    g = RecordGather(torch.device("cpu"))
    for _ in range(2):  # the buffers are reused from pass to pass
        got = g([mine])
    if rank == 0:
        np.save(out_path, np.concatenate(got))
        assert [x.size for x in got][0] == mine.size
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    from trgt_b200.shard import shard_bounds
    assert shard_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert shard_bounds(1_000_000, 8)[-1] == 1_000_000 and shard_bounds(3, 8).count(3) == 6


def test_two_rank_gather_equals_single_process(tmp_path):
    n_loci, world = 37, 2
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, _free_port(), n_loci, out), nprocs=world, join=True)
    gathered = np.load(out)
    from trgt_b200.shard import shard_bounds
    b = shard_bounds(n_loci, world)
    expect = np.concatenate([_payload(n_loci, b[r], b[r + 1] - b[r]) for r in range(world)])
    assert np.array_equal(gathered, expect)
    # shards regenerate the catalog exactly: rank 1's first locus is locus b[1] of a single-process run
    from harness import workload
    whole = workload.generate(n_loci, 8, seed=31337)
    part = workload.generate(b[2] - b[1], 8, locus_begin=b[1], seed=31337)
    assert part.reads.get(0) == whole.reads.get(b[1] * 8) and part.left.get(3) == whole.left.get(b[1] + 3)
