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


class LogitsGather(object):
    """The per-step exchange of a sharded evaluation as ONE collective with no per-step allocation: every rank
    contributes a fixed block ``[per * n_labels float32 logits | int64 correct | int64 total]`` and one
    ``all_gather_into_tensor`` of those blocks returns both the full logits (shard_bounds order) and the summed
    accuracy counts (metric/acc.py:16-22).  The model writes its logits straight into ``self.logits`` (``out=`` of
    forward / forward_wave) and ``metric.Acc(counts=self.counts)`` counts into ``self.counts``."""

    def __init__(self, n_total, n_labels, device):
        rank, ws = world()
        self.rank, self.ws, self.n_total, self.n_labels = rank, ws, int(n_total), int(n_labels)
        self.per = -(-self.n_total // ws) if n_total > 0 else 0
        self.lo, self.hi = shard_bounds(self.n_total, rank, ws)
        row_bytes = self.per * self.n_labels * 4
        self.block = row_bytes + 16                      # (row_bytes is a multiple of 4; pad so the counts are 8-aligned)
        self.block += (-self.block) % 16
        self._cnt_off = self.block - 16
        self.send = torch.zeros(self.block, dtype=torch.uint8, device=device)
        self.recv = torch.zeros(ws * self.block, dtype=torch.uint8, device=device) if ws > 1 else self.send
        self.logits = self.send[:row_bytes].view(torch.float32).view(self.per, self.n_labels)[: self.hi - self.lo]
        self.counts = self.send[self._cnt_off:self._cnt_off + 16].view(torch.int64)
        n_blk = ws if ws > 1 else 1
        # strided views of the gathered blocks: logits [ws, per, n_labels] float32, counts [ws, 2] int64
        self._all_logits = self.recv.view(torch.float32).as_strided(
            (n_blk, self.per, self.n_labels), (self.block // 4, self.n_labels, 1))
        self._all_counts = self.recv.view(torch.int64).as_strided((n_blk, 2), (self.block // 8, 1), self._cnt_off // 8)
        self.full = torch.zeros((n_blk * self.per, self.n_labels), dtype=torch.float32, device=device)
        self.total_counts = torch.zeros(2, dtype=torch.int64, device=device)

    def exchange(self):
        """-> (logits [n_total, n_labels] on every rank, int64[2] = [correct, total] summed over ranks); both are
        views of buffers owned by this object (valid until the next exchange)."""
        if self.ws > 1:
            if self.send.is_cuda and hasattr(dist, "all_gather_into_tensor"):
                dist.all_gather_into_tensor(self.recv, self.send)
            else:   # gloo (CPU tests)
                parts = [torch.empty_like(self.send) for _ in range(self.ws)]
                dist.all_gather(parts, self.send)
                self.recv.copy_(torch.cat(parts))
        self.full.view(-1, self.per, self.n_labels).copy_(self._all_logits)
        torch.sum(self._all_counts, dim=0, out=self.total_counts)
        return self.full[: self.n_total], self.total_counts


def all_reduce_counts(counts):
    """Sum an int64 count vector ([correct, total] or per-class counts) over ranks, in place."""
    if is_dist():
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts
