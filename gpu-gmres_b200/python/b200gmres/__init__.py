"""ctypes binding of libb200gmres.so (the C ABI in include/b200gmres.h) plus synthetic workloads.

Python is only the test/bench harness here; the product is the CUDA library and the C++ shim
that mirrors the reference's solver API (gpu-gmres_b200/host).
"""
from . import synth  # noqa: F401
