"""Sparse (transposed) convolution layers and their autograd functions
(reference: MinkowskiConvolution.py:42-634)."""
import math
from typing import Union

import torch
from torch.autograd import Function
from torch.nn import Parameter

from . import backend as _C
from .backend import CoordinateMapKey
from .common import MinkowskiModuleBase
from .coordinate_manager import CoordinateManager
from .enums import ConvolutionMode, RegionType
from .kernel_generator import KernelGenerator
from .sparse_tensor import SparseTensor, _get_coordinate_map_key


class MinkowskiConvolutionFunction(Function):
    @staticmethod
    def forward(ctx, input_features: torch.Tensor, kernel_weights: torch.Tensor,
                kernel_generator: KernelGenerator, convolution_mode: ConvolutionMode,
                in_coordinate_map_key: CoordinateMapKey,
                out_coordinate_map_key: CoordinateMapKey = None,
                coordinate_manager: CoordinateManager = None):
        if out_coordinate_map_key is None:
            out_coordinate_map_key = CoordinateMapKey(
                in_coordinate_map_key.get_coordinate_size())
        input_features = input_features.contiguous()
        ctx.input_features = input_features
        ctx.kernel_weights = kernel_weights
        ctx.misc = (kernel_generator, convolution_mode, in_coordinate_map_key,
                    out_coordinate_map_key, coordinate_manager)
        return _C.ConvolutionForwardGPU(
            input_features, kernel_weights, kernel_generator.kernel_size,
            kernel_generator.kernel_stride, kernel_generator.kernel_dilation,
            kernel_generator.region_type, kernel_generator.region_offsets,
            kernel_generator.expand_coordinates, convolution_mode, in_coordinate_map_key,
            out_coordinate_map_key, coordinate_manager._manager)

    @staticmethod
    def backward(ctx, grad_out_feat: torch.Tensor):
        grad_out_feat = grad_out_feat.contiguous()
        kgen, mode, in_key, out_key, manager = ctx.misc
        grad_in_feat, grad_kernel = _C.ConvolutionBackwardGPU(
            ctx.input_features, grad_out_feat, ctx.kernel_weights, kgen.kernel_size,
            kgen.kernel_stride, kgen.kernel_dilation, kgen.region_type, kgen.region_offsets,
            mode, in_key, out_key, manager._manager,
            need_in=ctx.needs_input_grad[0], need_w=ctx.needs_input_grad[1])
        return grad_in_feat, grad_kernel, None, None, None, None, None


class MinkowskiConvolutionTransposeFunction(Function):
    @staticmethod
    def forward(ctx, input_features: torch.Tensor, kernel_weights: torch.Tensor,
                kernel_generator: KernelGenerator, convolution_mode: ConvolutionMode,
                in_coordinate_map_key: CoordinateMapKey,
                out_coordinate_map_key: CoordinateMapKey = None,
                coordinate_manager: CoordinateManager = None):
        if out_coordinate_map_key is None:
            out_coordinate_map_key = CoordinateMapKey(
                in_coordinate_map_key.get_coordinate_size())
        input_features = input_features.contiguous()
        ctx.input_features = input_features
        ctx.kernel_weights = kernel_weights
        ctx.misc = (kernel_generator, convolution_mode, in_coordinate_map_key,
                    out_coordinate_map_key, coordinate_manager)
        return _C.ConvolutionTransposeForwardGPU(
            input_features, kernel_weights, kernel_generator.kernel_size,
            kernel_generator.kernel_stride, kernel_generator.kernel_dilation,
            kernel_generator.region_type, kernel_generator.region_offsets,
            kernel_generator.expand_coordinates, convolution_mode, in_coordinate_map_key,
            out_coordinate_map_key, coordinate_manager._manager)

    @staticmethod
    def backward(ctx, grad_out_feat: torch.Tensor):
        grad_out_feat = grad_out_feat.contiguous()
        kgen, mode, in_key, out_key, manager = ctx.misc
        grad_in_feat, grad_kernel = _C.ConvolutionTransposeBackwardGPU(
            ctx.input_features, grad_out_feat, ctx.kernel_weights, kgen.kernel_size,
            kgen.kernel_stride, kgen.kernel_dilation, kgen.region_type, kgen.region_offsets,
            mode, in_key, out_key, manager._manager,
            need_in=ctx.needs_input_grad[0], need_w=ctx.needs_input_grad[1])
        return grad_in_feat, grad_kernel, None, None, None, None, None


class MinkowskiConvolutionBase(MinkowskiModuleBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1,
                 bias=False, kernel_generator=None, is_transpose=False,
                 expand_coordinates=False, convolution_mode=ConvolutionMode.DEFAULT,
                 dimension=-1):
        super().__init__()
        assert dimension > 0, \
            f"Invalid dimension. Please provide a valid dimension argument. dimension={dimension}"
        if kernel_generator is None:
            kernel_generator = KernelGenerator(kernel_size=kernel_size, stride=stride,
                                               dilation=dilation,
                                               expand_coordinates=expand_coordinates,
                                               dimension=dimension)
        else:
            kernel_generator.expand_coordinates = expand_coordinates
        self.is_transpose = is_transpose
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_generator = kernel_generator
        self.dimension = dimension
        # kernel volume 1 and stride 1: plain matrix product (MinkowskiConvolution.py:265-270)
        self.use_mm = (kernel_generator.kernel_volume == 1
                       and kernel_generator.requires_strided_coordinates)
        if self.use_mm:
            kernel_shape = (in_channels, out_channels)
        else:
            kernel_shape = (kernel_generator.kernel_volume, in_channels, out_channels)
        self.kernel = Parameter(torch.empty(*kernel_shape, dtype=torch.float32))
        self.bias = Parameter(torch.empty(1, out_channels, dtype=torch.float32)) if bias else None
        self.convolution_mode = convolution_mode
        self.conv = (MinkowskiConvolutionTransposeFunction if is_transpose
                     else MinkowskiConvolutionFunction)

    def forward(self, input: SparseTensor,
                coordinates: Union[torch.Tensor, CoordinateMapKey, SparseTensor] = None):
        assert isinstance(input, SparseTensor)
        assert input.D == self.dimension
        if self.use_mm:
            out_coordinate_map_key = input.coordinate_map_key
            kernel = self.kernel if self.kernel.dtype == input.F.dtype else self.kernel.to(input.F.dtype)
            outfeat = input.F.mm(kernel)
        else:
            out_coordinate_map_key = _get_coordinate_map_key(
                input, coordinates, expand_coordinates=self.kernel_generator.expand_coordinates)
            outfeat = self.conv.apply(input.F, self.kernel, self.kernel_generator,
                                      self.convolution_mode, input.coordinate_map_key,
                                      out_coordinate_map_key, input._manager)
        if self.bias is not None:
            outfeat = outfeat + self.bias.to(outfeat.dtype)
        return SparseTensor(outfeat, coordinate_map_key=out_coordinate_map_key,
                            coordinate_manager=input._manager)

    def reset_parameters(self, is_transpose=False):
        with torch.no_grad():
            n = (self.out_channels if is_transpose else self.in_channels) \
                * self.kernel_generator.kernel_volume
            stdv = 1.0 / math.sqrt(n)
            self.kernel.data.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.data.uniform_(-stdv, stdv)

    def __repr__(self):
        s = f"(in={self.in_channels}, out={self.out_channels}, "
        if self.kernel_generator.region_type in [RegionType.CUSTOM]:
            s += (f"region_type={self.kernel_generator.region_type}, "
                  f"kernel_volume={self.kernel_generator.kernel_volume}, ")
        else:
            s += f"kernel_size={self.kernel_generator.kernel_size}, "
        s += (f"stride={self.kernel_generator.kernel_stride}, "
              f"dilation={self.kernel_generator.kernel_dilation})")
        return self.__class__.__name__ + s


class MinkowskiConvolution(MinkowskiConvolutionBase):
    """Sparse convolution layer (reference: MinkowskiConvolution.py:360-451)."""

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1,
                 bias=False, kernel_generator=None, expand_coordinates=False,
                 convolution_mode=ConvolutionMode.DEFAULT, dimension=None):
        MinkowskiConvolutionBase.__init__(
            self, in_channels, out_channels, kernel_size, stride, dilation, bias,
            kernel_generator, is_transpose=False, expand_coordinates=expand_coordinates,
            convolution_mode=convolution_mode, dimension=dimension)
        self.reset_parameters()


class MinkowskiConvolutionTranspose(MinkowskiConvolutionBase):
    """Sparse transposed convolution layer (reference: MinkowskiConvolution.py:454-543)."""

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1,
                 bias=False, kernel_generator=None, expand_coordinates=False,
                 convolution_mode=ConvolutionMode.DEFAULT, dimension=None):
        if kernel_generator is None:
            kernel_generator = KernelGenerator(kernel_size=kernel_size, stride=stride,
                                               dilation=dilation, dimension=dimension)
        MinkowskiConvolutionBase.__init__(
            self, in_channels, out_channels, kernel_size, stride, dilation, bias,
            kernel_generator, is_transpose=True, expand_coordinates=expand_coordinates,
            convolution_mode=convolution_mode, dimension=dimension)
        self.reset_parameters(True)


class MinkowskiGenerativeConvolutionTranspose(MinkowskiConvolutionBase):
    """Transposed convolution that always generates new output coordinates
    (reference: MinkowskiConvolution.py:546-634)."""

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1,
                 bias=False, kernel_generator=None,
                 convolution_mode=ConvolutionMode.DEFAULT, dimension=None):
        if kernel_generator is None:
            kernel_generator = KernelGenerator(kernel_size=kernel_size, stride=stride,
                                               dilation=dilation, expand_coordinates=True,
                                               dimension=dimension)
        else:
            kernel_generator.expand_coordinates = True
        MinkowskiConvolutionBase.__init__(
            self, in_channels, out_channels, kernel_size, stride, dilation, bias,
            kernel_generator, is_transpose=True, expand_coordinates=True,
            convolution_mode=convolution_mode, dimension=dimension)
        self.reset_parameters(True)
