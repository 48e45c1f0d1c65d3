"""genomic_b200 -- B200 (sm_100a) implementation of the CBS + smoothing hot path of djhshih/genomic.

Only what that path needs lives here: csrc/ (CUDA kernels + the C ABI of include/cbs_gpu.h),
host/ (C++ mirror of the reference interface, synthetic inputs) and this ctypes binding.
"""
from .binding import (  # noqa: F401
    BatchResult, CbsGpuError, Context, Params, RNG_MT19937_64, RNG_PHILOX, EXPORTED_SYMBOLS, LIB_PATH,
    load_library,
)
