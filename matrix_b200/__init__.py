"""matrix_b200 -- B200-native batched engine for ppx's data-parallel hot path.

  matrix_b200.batch   device tensors in / out (torch CUDA float64), one kernel launch per call
  matrix_b200.host    numpy host arrays in / out through the chunked, copy-overlapped host ABI
  matrix_b200.synth   counter-based synthetic workloads (numpy; mirrored on the device)
  matrix_b200._lib    ctypes binding of libppx_cuda.so (include/ppx_cuda.h)

Everything numerical happens inside libppx_cuda.so (hand-written sm_100a kernels, matrix_b200/csrc).
There is no Python/CPU implementation: importing .batch or .host without the built library raises.
"""
import importlib

__all__ = ["batch", "host", "synth", "_lib"]


def __getattr__(name):
    if name in __all__:
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
