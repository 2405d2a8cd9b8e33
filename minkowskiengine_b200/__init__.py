"""minkowskiengine_b200 — a Blackwell-native (sm_100a) backend for the sparse-convolution
hot path of NVIDIA/MinkowskiEngine, behind the reference's own Python API.

    import minkowskiengine_b200 as ME
    x = ME.SparseTensor(features, coordinates, device="cuda")
    y = ME.MinkowskiConvolution(64, 128, kernel_size=3, stride=2, dimension=3).cuda()(x)

Scope: SparseTensor / CoordinateManager / kernel maps, MinkowskiConvolution,
MinkowskiConvolutionTranspose, local pooling, and the thin torch wrappers MinkUNet needs.
CUDA tensors only; all device work runs in csrc/libmeb200.so (include/meb200.h).
"""
__version__ = "0.1.0"

from . import _lib
from .backend import (CoordinateMapKey, cuda_version, cudart_version, get_gpu_memory_info,
                      is_cuda_available)
from .enums import (BroadcastMode, ConvolutionMode, CoordinateMapType, CUDAKernelMapMode,
                    GPUMemoryAllocatorType, MinkowskiAlgorithm, PoolingMode, RegionType)
from .kernel_generator import (KernelGenerator, KernelRegion, convert_region_type,
                               get_kernel_volume)
from .sparse_tensor import (SparseTensor, SparseTensorOperationMode,
                            SparseTensorQuantizationMode, clear_global_coordinate_manager,
                            global_coordinate_manager, set_global_coordinate_manager,
                            set_sparse_tensor_operation_mode, sparse_tensor_operation_mode)
from .common import MinkowskiModuleBase, convert_to_int_list, convert_to_int_tensor
from .coordinate_manager import (CoordinateManager, CoordsManager, set_gpu_allocator,
                                 set_memory_manager_backend)
from .convolution import (MinkowskiConvolution, MinkowskiConvolutionFunction,
                          MinkowskiConvolutionTranspose, MinkowskiConvolutionTransposeFunction,
                          MinkowskiGenerativeConvolutionTranspose)
from .pooling import (MinkowskiAvgPooling, MinkowskiLocalPoolingFunction, MinkowskiMaxPooling,
                      MinkowskiSumPooling)
from .normalization import MinkowskiBatchNorm, MinkowskiSyncBatchNorm, fused_bn_relu
from .nonlinearity import *  # noqa: F401,F403
from .ops import MinkowskiLinear, cat
from . import modules, utils  # noqa: F401
