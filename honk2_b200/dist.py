"""Batch sharding for one-process-per-GPU inference (replaces the reference's
``torch.nn.DataParallel`` wrap, /root/reference/run/test.py:69-70): utterances are independent
(eval-mode BatchNorm, resnet.py:55), so the batch is split contiguously on dim 0 with no
data-path collective; NCCL (or gloo in the CPU tests) is used only to all-gather the logits and
to all-reduce the accuracy counts (metric/acc.py:16-22)."""
import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized()


def world():
    return (dist.get_rank(), dist.get_world_size()) if is_dist() else (0, 1)


def shard_bounds(n, rank, world_size):
    """Contiguous split like DataParallel's scatter (torch.chunk): ceil(n / world) per rank,
    trailing ranks may get fewer (or zero) utterances."""
    per = -(-n // world_size) if n > 0 else 0
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def all_gather_rows(local, n_total):
    """Concatenate per-rank row blocks (shard_bounds order) into the full [n_total, ...] tensor
    on every rank.  One all_gather of equal-sized (padded) blocks."""
    if not is_dist():
        return local
    rank, ws = world()
    per = -(-n_total // ws) if n_total > 0 else 0
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((ws * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad) if hasattr(dist, "all_gather_into_tensor") and local.is_cuda else \
        _all_gather_list(out, pad, ws, per)
    return out[:n_total]


def _all_gather_list(out, pad, ws, per):
    parts = [out[r * per:(r + 1) * per] for r in range(ws)]
    tmp = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(tmp, pad)
    for p, t in zip(parts, tmp):
        p.copy_(t)


def all_reduce_counts(counts):
    """Sum an int64 count vector ([correct, total] or per-class counts) over ranks, in place."""
    if is_dist():
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts
