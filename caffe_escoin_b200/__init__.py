"""caffe_escoin_b200 -- B200-native (sm_100a) rebuild of caffe-escoin's one hot path: the Escort direct
sparse convolution (CSR pack -> direct sparse conv forward with fused bias/ReLU -> masked backward).

The product is the C-ABI shared library `libescort_b200.so` (include/escort_b200.h) plus the C++ host mirror
of the reference's ConvolutionLayer surface (caffe_escoin_b200/host).  This Python package only binds the
C ABI with ctypes (torch supplies device memory, streams and torch.distributed); it contains no compute
fallback: if the CUDA library is missing, importing `capi` raises.
"""
from . import workloads  # noqa: F401  (numpy only)

__all__ = ["workloads", "capi"]


def __getattr__(name):
    if name == "capi":
        import importlib
        return importlib.import_module(".capi", __name__)
    raise AttributeError(name)
