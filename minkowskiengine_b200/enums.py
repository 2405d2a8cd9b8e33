"""Enumerations of the reference's pybind boundary (pybind/extern.hpp:669-741,
src/types.hpp:116-170), same member names and integer values."""
from enum import Enum


class GPUMemoryAllocatorType(Enum):
    PYTORCH = 0
    CUDA = 1


class CUDAKernelMapMode(Enum):
    MEMORY_EFFICIENT = 0
    SPEED_OPTIMIZED = 1


class MinkowskiAlgorithm(Enum):
    DEFAULT = 0
    MEMORY_EFFICIENT = 1
    SPEED_OPTIMIZED = 2


class CoordinateMapType(Enum):
    CPU = 0
    CUDA = 1


class RegionType(Enum):
    HYPER_CUBE = 0
    HYPER_CROSS = 1
    CUSTOM = 2


class PoolingMode(Enum):
    LOCAL_SUM_POOLING = 0
    LOCAL_AVG_POOLING = 1
    LOCAL_MAX_POOLING = 2
    GLOBAL_SUM_POOLING_DEFAULT = 3
    GLOBAL_AVG_POOLING_DEFAULT = 4
    GLOBAL_MAX_POOLING_DEFAULT = 5
    GLOBAL_SUM_POOLING_KERNEL = 6
    GLOBAL_AVG_POOLING_KERNEL = 7
    GLOBAL_MAX_POOLING_KERNEL = 8
    GLOBAL_SUM_POOLING_PYTORCH_INDEX = 9
    GLOBAL_AVG_POOLING_PYTORCH_INDEX = 10
    GLOBAL_MAX_POOLING_PYTORCH_INDEX = 11


class BroadcastMode(Enum):
    ELEMENTWISE_ADDITON = 0
    ELEMENTWISE_MULTIPLICATION = 1


class ConvolutionMode(Enum):
    DEFAULT = 0
    DIRECT_GEMM = 1
    COPY_GEMM = 2
