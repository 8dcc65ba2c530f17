"""``metric.Acc`` with the reference contract (/root/reference/metric/acc.py:8-31) but counted
on the device by ``kws_acc_accumulate``: ``accumulate`` never synchronises; ``get_metric`` does
one device->host read.  (The reference does a ``.item()`` per batch, acc.py:18.)"""
import ctypes as C

import torch

from . import _native
from .class_registry import register_cls
from .dist import all_reduce_counts


class _MetricType(object):
    """Stand-in for the reference's ``MetricType`` enum (/root/reference/metric/metric_utils.py:5-7) that compares equal
    to it by ``.value``: ``criterion.get_type() == MetricType.MACRO`` (run/train.py:104) and ``collect_metrics``
    (metric_utils.py:42-45) put OUR object on the left of ``==``, so both work with the reference's own enum."""
    __slots__ = ("name", "value")

    def __init__(self, value):
        self.name = self.value = value

    def __eq__(self, other):
        return getattr(other, "value", other) == self.value

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.value)

    def __repr__(self):
        return f"<MetricType.{self.name}: {self.value!r}>"


class MetricType(object):
    MACRO = _MetricType("MACRO")
    MICRO = _MetricType("MICRO")


@register_cls('metric.Acc')
class Acc(object):
    def __init__(self, counts=None):
        """counts: optional device int64[2] to accumulate into (e.g. a view of a collective's send buffer)."""
        self._counts = counts  # device int64 [correct, total]

    def get_type(self):       # MicroMetric.get_type (metric_utils.py:23-28): a scalar metric, reported as is
        return MetricType.MACRO

    def accumulate(self, output, target, return_pred=False):
        if not output.is_cuda:
            raise _native.NativeError("metric.Acc counts on the GPU; got a CPU tensor")
        assert output.shape[0] == len(target)
        if self._counts is None or self._counts.device != output.device:
            self._counts = torch.zeros(2, dtype=torch.int64, device=output.device)
        output = output.float().contiguous()
        target = target.to(output.device, torch.int64).contiguous()
        pred = torch.empty(output.shape[0], dtype=torch.int64, device=output.device) if return_pred else None
        lib = _native.load()
        with torch.cuda.device(output.device):
            _native.check(lib.kws_acc_accumulate(
                C.c_void_p(output.data_ptr()), C.c_void_p(target.data_ptr()), output.shape[0], output.shape[1],
                C.c_void_p(self._counts.data_ptr()), C.c_void_p(pred.data_ptr()) if return_pred else None,
                C.c_void_p(torch.cuda.current_stream(output.device).cuda_stream)), "kws_acc_accumulate")
        return pred

    def all_reduce(self):
        """Sum the counts over ranks (one NCCL all-reduce of int64[2])."""
        if self._counts is not None:
            all_reduce_counts(self._counts)
        return self

    def counts(self):
        if self._counts is None:
            return 0, 0
        c = self._counts.tolist()
        return int(c[0]), int(c[1])

    def get_metric(self):
        correct, total = self.counts()
        return correct / total

    def reset_metric(self):
        if self._counts is not None:
            self._counts.zero_()


def _eval_accumulate(output, target, counts=None, class_counts=None, loss_sum=None, pred=None):
    """kws_eval_accumulate on the current stream of `output`'s device."""
    if not output.is_cuda:
        raise _native.NativeError("evaluation statistics are counted on the GPU; got a CPU tensor")
    lib = _native.load()
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None   # noqa: E731
    with torch.cuda.device(output.device):
        _native.check(lib.kws_eval_accumulate(
            ptr(output), ptr(target), output.shape[0], output.shape[1], ptr(counts), ptr(class_counts), ptr(loss_sum),
            ptr(pred), C.c_void_p(torch.cuda.current_stream(output.device).cuda_stream)), "kws_eval_accumulate")


@register_cls('metric.PerClassAcc')
class PerClassAcc(object):
    """``metric.PerClassAcc`` (/root/reference/metric/per_class_acc.py:8-55) with the counts kept on the device:
    ``accumulate`` never synchronises (the reference does two ``.tolist()`` per batch, :19-20) and returns None
    instead of the batch's own per-class dict (``evaluate`` ignores it, run/test.py:33); ``get_metric`` returns
    ``{class index: correct / total}`` for the classes seen, exactly like the reference (:47-51)."""

    def __init__(self):
        self._counts = None   # device int64 [n_labels][2] = (total, correct)

    def get_type(self):       # MacroMetric.get_type (metric_utils.py:31-36): re-keyed by label in collect_metrics
        return MetricType.MICRO

    def accumulate(self, output, target):
        assert output.shape[0] == len(target)
        n_labels = output.shape[1]
        if self._counts is None or self._counts.device != output.device or self._counts.shape[0] != n_labels:
            self._counts = torch.zeros((n_labels, 2), dtype=torch.int64, device=output.device)
        output = output.float().contiguous()
        target = target.to(output.device, torch.int64).contiguous()
        _eval_accumulate(output, target, class_counts=self._counts)

    def all_reduce(self):
        if self._counts is not None:
            all_reduce_counts(self._counts)
        return self

    def get_metric(self):
        if self._counts is None:
            return {}
        c = self._counts.tolist()
        return {k: cor / tot for k, (tot, cor) in enumerate(c) if tot > 0}

    def reset_metric(self):
        self._counts = None


@register_cls('loss_fn.ce_loss')
def ce_loss(output, target):
    """``loss_fn.ce_loss`` (/root/reference/loss_function.py:7-9): mean categorical cross entropy of raw logits,
    as a 0-dim tensor on the logits' device; computed by kws_eval_accumulate (no host sync).  When the logits carry
    an autograd graph (run/train.py:142-143 calls ``loss.backward()``) or live on the CPU, this is the reference's own
    ``nn.CrossEntropyLoss`` -- training is outside the native path."""
    if output.requires_grad or not output.is_cuda:
        return torch.nn.functional.cross_entropy(output, target)
    output = output.detach().float().contiguous()
    target = target.to(output.device, torch.int64).contiguous()
    s = torch.zeros(1, dtype=torch.float64, device=output.device)
    if output.shape[0] > 0:
        _eval_accumulate(output, target, loss_sum=s)
    return (s[0] / max(output.shape[0], 1)).float()
