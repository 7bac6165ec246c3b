"""pangulu_b200 -- sm_100a (B200) platform module for PanguLU's numeric factorisation.

csrc/platform : CUDA kernels + the C ABI of include/pangulu_platform_0201001.h
csrc/host     : host driver mirroring the reference API (include/pangulu_sm100_driver.h)
solver.py     : ctypes face of the driver;  matrices.py : synthetic inputs of the benchmark configs
"""
from ._lib import load, lib_path, PLATFORM_FUNCTIONS, EXTENSION_FUNCTIONS, DRIVER_FUNCTIONS  # noqa: F401
from .solver import PanguLU, driver_config  # noqa: F401
