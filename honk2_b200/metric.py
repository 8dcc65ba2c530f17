"""``metric.Acc`` with the reference contract (/root/reference/metric/acc.py:8-31) but counted
on the device by ``kws_acc_accumulate``: ``accumulate`` never synchronises; ``get_metric`` does
one device->host read.  (The reference does a ``.item()`` per batch, acc.py:18.)"""
import ctypes as C

import torch

from . import _native
from .class_registry import register_cls
from .dist import all_reduce_counts


@register_cls('metric.Acc')
class Acc(object):
    def __init__(self):
        self._counts = None  # device int64 [correct, total]

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
