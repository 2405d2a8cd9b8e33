"""Kernel shapes: sizes/strides/dilations of a layer and the offset enumeration.

Reference: MinkowskiKernelGenerator.py:245-310 (KernelGenerator) and the offset order of
kernel_region::coordinate_at (src/kernel_region.hpp:198-247), pinned by the goldens in
tests/cpp/kernel_region_cpu_test.py:24-82:
  HYPER_CUBE : kernel index k = i_0 + K_0*(i_1 + K_1*(i_2 ...)) (axis 0 fastest); odd K_a
               -> (i_a - K_a//2) * dilation_a * tensor_stride_a, even K_a -> i_a * dil * ts
  HYPER_CROSS: k = 0 is the centre, then per axis the K_a-1 non-centre taps, positive
               offsets first (1..r) followed by the negative ones (-r..-1)
  CUSTOM     : user offsets (the reference leaves this unimplemented,
               kernel_region.hpp:243-245; here they are used as given, scaled by ts)
"""
from functools import reduce

import torch

from .common import convert_to_int_list
from .enums import RegionType


def get_kernel_volume(region_type, kernel_size, region_offset, axis_types, dimension):
    if region_type == RegionType.HYPER_CUBE:
        assert all(k > 0 for k in kernel_size), "kernel_size must be positive"
        assert region_offset is None or len(region_offset) == 0, \
            "region_offset must be empty unless region_type is CUSTOM"
        assert axis_types is None, "axis_types must be None when region_type is given"
        return int(reduce(lambda a, b: a * b, kernel_size, 1))
    if region_type == RegionType.HYPER_CROSS:
        assert all(k > 0 for k in kernel_size), "kernel_size must be positive"
        assert all(k % 2 == 1 for k in kernel_size), \
            "kernel_size must be odd for region_type HYPER_CROSS"
        return int(sum(k - 1 for k in kernel_size) + 1)
    if region_type == RegionType.CUSTOM:
        assert region_offset is not None and region_offset.numel() > 0, \
            "region_offset must be non empty when region_type is CUSTOM"
        assert region_offset.size(1) == dimension, \
            "region_offset must have the same dimension as the network"
        return int(region_offset.size(0))
    raise NotImplementedError(region_type)


def region_offsets(region_type, kernel_size, dilation, tensor_stride, custom_offsets=None):
    """List of per-offset coordinate deltas [K][D] (python ints), in kernel-index order."""
    D = len(kernel_size)
    if region_type == RegionType.HYPER_CUBE:
        K = reduce(lambda a, b: a * b, kernel_size, 1)
        out = []
        for k in range(K):
            rem, delta = k, []
            for a in range(D):
                ks = kernel_size[a]
                i = rem % ks
                rem //= ks
                step = dilation[a] * tensor_stride[a]
                delta.append(i * step if ks % 2 == 0 else (i - ks // 2) * step)
            out.append(delta)
        return out
    if region_type == RegionType.HYPER_CROSS:
        out = [[0] * D]
        for a in range(D):
            r = (kernel_size[a] - 1) // 2
            step = dilation[a] * tensor_stride[a]
            for ind in range(kernel_size[a] - 1):
                off = (ind + 1) if ind < r else (ind - 2 * r)
                delta = [0] * D
                delta[a] = off * step
                out.append(delta)
        return out
    if region_type == RegionType.CUSTOM:
        assert custom_offsets is not None and custom_offsets.numel() > 0
        return [[int(v) * tensor_stride[a] for a, v in enumerate(row)]
                for row in custom_offsets.tolist()]
    raise NotImplementedError(region_type)


class KernelGenerator:
    __slots__ = ("cache", "kernel_size", "kernel_stride", "kernel_dilation", "region_type",
                 "region_offsets", "axis_types", "dimension", "kernel_volume",
                 "requires_strided_coordinates", "expand_coordinates")

    def __init__(self, kernel_size=-1, stride=1, dilation=1, is_transpose: bool = False,
                 region_type: RegionType = RegionType.HYPER_CUBE,
                 region_offsets: torch.Tensor = None, expand_coordinates: bool = False,
                 axis_types=None, dimension=-1):
        assert dimension > 0
        assert isinstance(region_type, RegionType)
        self.cache = {}
        self.kernel_size = convert_to_int_list(kernel_size, dimension)
        self.kernel_stride = convert_to_int_list(stride, dimension)
        self.kernel_dilation = convert_to_int_list(dilation, dimension)
        self.region_type = region_type
        has_off = region_offsets is not None and region_offsets.numel() > 0
        self.region_offsets = region_offsets.int() if has_off else torch.IntTensor()
        self.axis_types = axis_types
        self.dimension = dimension
        self.kernel_volume = get_kernel_volume(
            region_type, self.kernel_size, self.region_offsets if has_off else None,
            axis_types, dimension)
        # (sic) the reference's name: True iff every stride is 1
        # (MinkowskiKernelGenerator.py:303-305)
        self.requires_strided_coordinates = all(s == 1 for s in self.kernel_stride)
        self.expand_coordinates = expand_coordinates

    def get_kernel(self, tensor_stride, is_transpose):
        assert len(tensor_stride) == self.dimension
        key = tuple(tensor_stride)
        if key not in self.cache:
            self.cache[key] = (self.region_type, self.region_offsets, self.kernel_volume)
        return self.cache[key]

    def __repr__(self):
        return (f"{self.__class__.__name__}(kernel_size={self.kernel_size}, "
                f"kernel_stride={self.kernel_stride}, kernel_dilation={self.kernel_dilation}, "
                f"region_type={self.region_type}, expand_coordinates={self.expand_coordinates}, "
                f"dimension={self.dimension})")


class KernelRegion:
    """Thin record of a kernel region (reference: MinkowskiKernelGenerator.py:356-395)."""
    __slots__ = ("kernel_size", "kernel_stride", "kernel_dilation", "region_type",
                 "region_offset", "D")

    def __init__(self, kernel_size=-1, kernel_stride=1, kernel_dilation=1,
                 region_type=RegionType.HYPER_CUBE, region_offset=None, dimension=-1):
        assert dimension > 0
        self.kernel_size = convert_to_int_list(kernel_size, dimension)
        self.kernel_stride = convert_to_int_list(kernel_stride, dimension)
        self.kernel_dilation = convert_to_int_list(kernel_dilation, dimension)
        self.region_type = region_type
        self.region_offset = region_offset
        self.D = dimension


def convert_region_type(region_type, tensor_stride, kernel_size, up_stride, dilation,
                        region_offset, axis_types, dimension, center=True):
    """Reference helper kept for API parity (MinkowskiKernelGenerator.py:80-240): returns
    (region_type, region_offset tensor, kernel_volume)."""
    if region_type == RegionType.HYPER_CUBE:
        off = torch.IntTensor()
    else:
        off = torch.IntTensor(region_offsets(
            region_type, list(kernel_size), list(dilation), [1] * dimension,
            region_offset))
    vol = get_kernel_volume(region_type, list(kernel_size),
                            region_offset if region_type == RegionType.CUSTOM else None,
                            None, dimension)
    return region_type, off, vol
