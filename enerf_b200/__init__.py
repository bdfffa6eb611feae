"""enerf_b200 -- B200-native render-time hot path of ENeRF (zju3dv/ENeRF lib/networks/enerf).

Python host (this package) -> C-ABI shared library (enerf_b200/csrc -> libenerf_b200.so, declared in
include/enerf_b200.h) -> hand-written sm_100a CUDA kernels.  There is no CPU fallback: every
compute entry point raises if the CUDA library cannot be loaded.
"""
__version__ = "0.1.0"
