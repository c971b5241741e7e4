"""erl_gaussian_process_b200 — B200-native (sm_100a) train/predict hot path of
ExistentialRobotics/erl_gaussian_process behind the C ABI of include/erl_gp_b200.h.

Python here is only the thin host-side mirror used by tests and bench.py; the product is the
CUDA library (csrc/) and the C++ drop-in headers (cpp/).
"""
from . import _capi, sharding
from ._capi import ErlGpError, KERNELS, load
from .host import (
    BatchGp,
    MultiDeviceBatchGp,
    Context,
    LidarGaussianProcess2D,
    NoisyInputGaussianProcess,
    RangeSensorGaussianProcess3D,
    SparsePseudoInputGaussianProcess,
    SpGpOccupancyMap,
    VanillaGaussianProcess,
    compute_ktest,
    compute_ktrain,
)

__all__ = [
    "BatchGp",
    "MultiDeviceBatchGp",
    "Context",
    "ErlGpError",
    "KERNELS",
    "LidarGaussianProcess2D",
    "NoisyInputGaussianProcess",
    "RangeSensorGaussianProcess3D",
    "SparsePseudoInputGaussianProcess",
    "SpGpOccupancyMap",
    "VanillaGaussianProcess",
    "compute_ktest",
    "compute_ktrain",
    "load",
    "sharding",
]
