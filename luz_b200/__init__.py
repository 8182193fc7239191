"""luz_b200 -- B200-native implementation of Luz's ray-traced deferred lighting path.

The product is native: luz_b200/csrc (CUDA kernels + the C ABI of include/luzrt.h) and
luz_b200/host (C++ mirror of Luz's GPUScene / DeferredRenderer / .luz loader).  The Python in this
package is only the ctypes harness the tests and bench.py drive those libraries with."""
__all__ = ["wire", "rt", "build"]
