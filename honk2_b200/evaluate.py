"""Evaluation loops.

``evaluate`` keeps the signature and result keys of the reference's
(/root/reference/run/test.py:18-41) so it can replace it under run/train.py:162,203; the loss is
accumulated on the device and read once at the end instead of ``.item()`` per batch.

``evaluate_waves`` is the batched form the reference approximates with DataLoader workers +
DataParallel: raw waveforms in, logits and accuracy out, sharded across ranks (dist.py).
"""
import torch

from .dist import all_gather_rows, shard_bounds, world
from .metric import Acc


def evaluate(device, prefix, model, data_loader, loss_fn, metrics, label_mapping):
    total_loss = None
    n_batches = 0
    model.eval()
    for data, target in data_loader:
        data, target = data.to(device, non_blocking=True), target.to(device, non_blocking=True)
        with torch.no_grad():
            output = model(data)
            loss = loss_fn(output, target)
        total_loss = loss.detach() if total_loss is None else total_loss + loss.detach()
        n_batches += 1
        for metric in metrics.values():
            metric.accumulate(output, target)
    results = {"loss": (float(total_loss) / n_batches) if n_batches else float("nan")}
    for name, metric in metrics.items():
        value = metric.get_metric()
        if isinstance(value, dict):  # per-class metrics are re-keyed by label (metric_utils.py:44-50)
            value = {label_mapping[k]: v for k, v in value.items()}
        results[f"metric_{name}"] = value
    return results


def evaluate_waves(model, audio_processor, waves, targets=None, batch_size=8192, device=None):
    """waves: [N, n_samples] float32 (CPU pinned or CUDA); this rank evaluates its contiguous
    shard in batches and returns (logits [N, n_labels] on every rank, accuracy or None)."""
    rank, ws = world()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = waves.shape[0]
    lo, hi = shard_bounds(n, rank, ws)
    acc = Acc()
    outs = []
    model.eval()
    with torch.no_grad():
        for b0 in range(lo, hi, batch_size):
            b1 = min(hi, b0 + batch_size)
            w = waves[b0:b1].to(device, non_blocking=True)
            logits = model.forward_wave(w, audio_processor)
            outs.append(logits)
            if targets is not None:
                acc.accumulate(logits, targets[b0:b1].to(device, non_blocking=True))
    local = torch.cat(outs) if outs else torch.empty((0, model.n_labels), dtype=torch.float32, device=device)
    full = all_gather_rows(local, n)
    accuracy = None
    if targets is not None:
        if acc._counts is None:
            acc._counts = torch.zeros(2, dtype=torch.int64, device=device)
        accuracy = acc.all_reduce().get_metric()
    return full, accuracy


class HostPipeline(object):
    """waveforms in (pinned) HOST memory -> logits in (pinned) HOST memory, with the host->device copies of the next
    sub-batches overlapped with the kernels of the current one (``slots`` staging buffers, one copy stream).  This is
    the e2e form of the collate + forward loop: the reference copies each batch synchronously from pageable memory
    (`data.to(device)`, run/test.py:23).

    ``__call__`` returns a CUDA event recorded after the last device->host copy of the call: the caller owns
    ``host_logits`` again once ``event.synchronize()`` returns (``sync=True``, the default, does that before
    returning).  Calls may be issued back to back without synchronising: a staging slot is only refilled after the
    forward that read it has finished (``consumed`` events, also across calls)."""

    def __init__(self, model, audio_processor, n_samples, sub_batch=2048, device=None, slots=3, dtype=torch.float32):
        self.model, self.ap = model, audio_processor
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.sub = int(sub_batch)
        self.slots = max(2, int(slots))
        if dtype not in (torch.float32, torch.int16):
            raise ValueError("HostPipeline stages float32 waveforms or int16 PCM samples")
        self.dtype = dtype   # int16: 16-bit PCM as in the wav files -- half the host->device bytes, identical logits
        self.stage = [torch.empty((self.sub, n_samples), dtype=dtype, device=self.device)
                      for _ in range(self.slots)]
        self.dev_logits = [torch.empty((self.sub, model.n_labels), dtype=torch.float32, device=self.device)
                           for _ in range(self.slots)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.copied = [torch.cuda.Event() for _ in range(self.slots)]
        self.consumed = [torch.cuda.Event() for _ in range(self.slots)]
        self._next = 0   # staging slot of the next sub-batch (keeps rotating across calls)

    def __call__(self, host_waves, host_logits, sync=True):
        n = host_waves.shape[0]
        if host_waves.dtype != self.dtype:
            raise ValueError(f"this pipeline stages {self.dtype} waveforms, got {host_waves.dtype}")
        main = torch.cuda.current_stream(self.device)
        spans = [(b0, min(n, b0 + self.sub)) for b0 in range(0, n, self.sub)]
        done = torch.cuda.Event()
        with torch.no_grad():
            # copies run `slots - 1` sub-batches ahead of the kernels
            issued = 0

            def issue_copy():
                nonlocal issued
                b0, b1 = spans[issued]
                slot = (self._next + issued) % self.slots
                with torch.cuda.stream(self.copy_stream):
                    # the forward that last read this slot (in this call or an earlier one) has finished
                    # (waiting on an event that was never recorded is a no-op)
                    self.copy_stream.wait_event(self.consumed[slot])
                    self.stage[slot][: b1 - b0].copy_(host_waves[b0:b1], non_blocking=True)
                    self.copied[slot].record(self.copy_stream)
                issued += 1

            while issued < min(self.slots - 1, len(spans)):
                issue_copy()
            for k, (b0, b1) in enumerate(spans):
                if issued < len(spans):
                    issue_copy()
                slot = (self._next + k) % self.slots
                main.wait_event(self.copied[slot])
                logits = self.model.forward_wave(self.stage[slot][: b1 - b0], self.ap,
                                                 out=self.dev_logits[slot][: b1 - b0])
                self.consumed[slot].record(main)
                host_logits[b0:b1].copy_(logits, non_blocking=True)
            done.record(main)
        self._next = (self._next + len(spans)) % self.slots
        if sync:
            done.synchronize()
        return done
