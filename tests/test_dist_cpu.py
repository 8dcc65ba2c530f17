"""World-size-2 gloo test of the sharding plumbing (dist.py): contiguous batch split, logits
all-gather and accuracy-count all-reduce reproduce the single-process result
(replaces DataParallel scatter/gather, /root/reference/run/test.py:69-70)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from honk2_b200.dist import LogitsGather, all_gather_rows, all_reduce_counts, shard_bounds
from oracle import model_ref


def test_shard_bounds_cover_batch():
    for n in (0, 1, 7, 8, 8192, 8193):
        for ws in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            chunks = torch.arange(n).chunk(ws) if n else []
            for (lo, hi), c in zip(spans, chunks):          # == torch.chunk == DataParallel scatter
                assert (lo, hi) == (int(c[0]), int(c[-1]) + 1)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(n, 12, generator=g)
    target = torch.randint(0, 12, (n,), generator=g)
    lo, hi = shard_bounds(n, rank, world)
    local = logits[lo:hi].clone()
    full = all_gather_rows(local, n)
    c, t = model_ref.acc_counts(local, target[lo:hi]) if hi > lo else (0, 0)
    counts = all_reduce_counts(torch.tensor([c, t], dtype=torch.int64))
    # the one-collective form: logits written into the gather's send block, counts next to them
    ok = True
    g1 = LogitsGather(n, 12, torch.device("cpu"))
    for rep in range(2):     # (buffers are reused from step to step; ragged shards: the last rank's block is padded)
        g1.logits.copy_(local + rep)
        g1.counts.copy_(torch.tensor([c + rep, t], dtype=torch.int64))
        full1, counts1 = g1.exchange()
        ok = ok and torch.equal(full1, logits + rep) and counts1.tolist() == [counts[0].item() + world * rep, n]
    q.put((rank, torch.equal(full, logits) and ok, counts.tolist(), list(model_ref.acc_counts(logits, target))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [13, 64])
def test_gather_and_count_world2(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, counts, want in results:
        assert same, f"rank {rank}: gathered logits differ from the unsharded batch"
        assert counts == want
