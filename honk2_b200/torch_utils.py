"""Output-size calculators with the contract of /root/reference/utils/torch_utils.py:29-65
(needed to size CNN ``lin_0``, cnn.py:24-46): floor((in + 2p - (d(k-1)+1)) / s + 1)."""
import math
from collections.abc import Iterable


def _tuple(x, n):
    return tuple(x) if isinstance(x, Iterable) else (x,) * n


def calculate_conv_output_size(input_size, kernel_size, stride=1, padding=0, dilation=1):
    n = len(input_size)
    stride, padding, dilation = _tuple(stride, n), _tuple(padding, n), _tuple(dilation, n)
    return [math.floor((s + 2 * padding[i] - (dilation[i] * (kernel_size[i] - 1) + 1)) / stride[i] + 1)
            for i, s in enumerate(input_size)]


def calculate_pool_output_size(input_size, kernel_size, stride=None, padding=0, dilation=1, ceil_mode=False):
    n = len(input_size)
    stride = _tuple(kernel_size if stride is None else stride, n)
    padding, dilation = _tuple(padding, n), _tuple(dilation, n)
    rnd = math.ceil if ceil_mode else math.floor
    return [rnd((s + 2 * padding[i] - (dilation[i] * (kernel_size[i] - 1) + 1)) / stride[i] + 1)
            for i, s in enumerate(input_size)]
