"""Small host helpers shared by the layer classes (reference: MinkowskiCommon.py)."""
from collections.abc import Sequence

import numpy as np
import torch


def convert_to_int_list(arg, dimension):
    """int | sequence | ndarray | Tensor -> list[int] of length `dimension`
    (reference: MinkowskiCommon.py:48-66)."""
    if isinstance(arg, list):
        out = [int(a) for a in arg]
    elif isinstance(arg, (Sequence, np.ndarray, torch.Tensor)):
        out = [int(a) for a in arg]
    elif np.isscalar(arg):
        out = [int(arg)] * dimension
    else:
        raise ValueError(f"cannot interpret {arg!r} as {dimension} integers")
    if len(out) != dimension:
        raise ValueError(f"expected {dimension} values, got {out}")
    return out


def convert_to_int_tensor(arg, dimension):
    return torch.IntTensor(convert_to_int_list(arg, dimension))


def prep_args(tensor_stride, stride, kernel_size, dilation, region_type, D=-1):
    assert all(k > 0 for k in kernel_size), f"kernel_size must be positive: {kernel_size}"
    assert all(s > 0 for s in stride), f"stride must be positive: {stride}"
    assert all(d > 0 for d in dilation), f"dilation must be positive: {dilation}"
    return tensor_stride, stride, kernel_size, dilation, region_type


class MinkowskiModuleBase(torch.nn.Module):
    pass
