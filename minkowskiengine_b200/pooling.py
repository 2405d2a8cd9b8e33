"""Local pooling layers (reference: MinkowskiPooling.py:42-440)."""
import torch
from torch.autograd import Function

from . import backend as _C
from .backend import CoordinateMapKey
from .common import MinkowskiModuleBase
from .coordinate_manager import CoordinateManager
from .enums import PoolingMode
from .kernel_generator import KernelGenerator
from .sparse_tensor import SparseTensor, _get_coordinate_map_key


class MinkowskiLocalPoolingFunction(Function):
    @staticmethod
    def forward(ctx, input_features: torch.Tensor, pooling_mode: PoolingMode,
                kernel_generator: KernelGenerator, in_coordinate_map_key: CoordinateMapKey,
                out_coordinate_map_key: CoordinateMapKey = None,
                coordinate_manager: CoordinateManager = None):
        if out_coordinate_map_key is None:
            out_coordinate_map_key = CoordinateMapKey(
                in_coordinate_map_key.get_coordinate_size())
        input_features = input_features.contiguous()
        ctx.input_features = input_features
        ctx.pooling_mode = pooling_mode
        ctx.kernel_generator = kernel_generator
        ctx.in_coordinate_map_key = in_coordinate_map_key
        ctx.out_coordinate_map_key = out_coordinate_map_key
        ctx.coordinate_manager = coordinate_manager
        out_feat, num_nonzero = _C.LocalPoolingForwardGPU(
            input_features, kernel_generator.kernel_size, kernel_generator.kernel_stride,
            kernel_generator.kernel_dilation, kernel_generator.region_type,
            kernel_generator.region_offsets, pooling_mode, in_coordinate_map_key,
            out_coordinate_map_key, coordinate_manager._manager)
        ctx.num_nonzero = num_nonzero
        return out_feat

    @staticmethod
    def backward(ctx, grad_out_feat):
        grad_out_feat = grad_out_feat.contiguous()
        kgen = ctx.kernel_generator
        grad_in_feat = _C.LocalPoolingBackwardGPU(
            ctx.input_features, grad_out_feat, ctx.num_nonzero, kgen.kernel_size,
            kgen.kernel_stride, kgen.kernel_dilation, kgen.region_type, kgen.region_offsets,
            ctx.pooling_mode, ctx.in_coordinate_map_key, ctx.out_coordinate_map_key,
            ctx.coordinate_manager._manager)
        return grad_in_feat, None, None, None, None, None


class MinkowskiPoolingBase(MinkowskiModuleBase):
    def __init__(self, kernel_size, stride=1, dilation=1, kernel_generator=None,
                 is_transpose=False, pooling_mode=PoolingMode.LOCAL_AVG_POOLING, dimension=-1):
        super().__init__()
        assert dimension > 0, \
            f"Invalid dimension. Please provide a valid dimension argument. dimension={dimension}"
        if stride == 1 and not is_transpose:
            pass  # the reference warns that stride 1 pooling keeps the coordinates
        if kernel_generator is None:
            kernel_generator = KernelGenerator(kernel_size=kernel_size, stride=stride,
                                               dilation=dilation, dimension=dimension)
        self.is_transpose = is_transpose
        self.kernel_generator = kernel_generator
        self.pooling_mode = pooling_mode
        self.dimension = dimension
        self.pooling = MinkowskiLocalPoolingFunction

    def forward(self, input: SparseTensor, coordinates=None):
        assert isinstance(input, SparseTensor)
        assert input.D == self.dimension
        out_key = _get_coordinate_map_key(input, coordinates)
        outfeat = self.pooling.apply(input.F, self.pooling_mode, self.kernel_generator,
                                     input.coordinate_map_key, out_key, input._manager)
        return SparseTensor(outfeat, coordinate_map_key=out_key,
                            coordinate_manager=input.coordinate_manager)

    def __repr__(self):
        return (f"{self.__class__.__name__}(kernel_size={self.kernel_generator.kernel_size}, "
                f"stride={self.kernel_generator.kernel_stride}, "
                f"dilation={self.kernel_generator.kernel_dilation})")


class MinkowskiAvgPooling(MinkowskiPoolingBase):
    """Average over the NON-ZERO members of each window (MinkowskiPooling.py:187-260)."""

    def __init__(self, kernel_size=-1, stride=1, dilation=1, kernel_generator=None,
                 dimension=None):
        MinkowskiPoolingBase.__init__(self, kernel_size, stride, dilation, kernel_generator,
                                      is_transpose=False,
                                      pooling_mode=PoolingMode.LOCAL_AVG_POOLING,
                                      dimension=dimension)


class MinkowskiSumPooling(MinkowskiPoolingBase):
    def __init__(self, kernel_size, stride=1, dilation=1, kernel_generator=None,
                 dimension=None):
        MinkowskiPoolingBase.__init__(self, kernel_size, stride, dilation, kernel_generator,
                                      is_transpose=False,
                                      pooling_mode=PoolingMode.LOCAL_SUM_POOLING,
                                      dimension=dimension)


class MinkowskiMaxPooling(MinkowskiPoolingBase):
    def __init__(self, kernel_size, stride=1, dilation=1, kernel_generator=None,
                 dimension=None):
        MinkowskiPoolingBase.__init__(self, kernel_size, stride, dilation, kernel_generator,
                                      is_transpose=False,
                                      pooling_mode=PoolingMode.LOCAL_MAX_POOLING,
                                      dimension=dimension)
