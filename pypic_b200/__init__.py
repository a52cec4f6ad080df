"""pypic_b200 -- B200 (sm_100a) implementation of pyPIC's per-timestep PIC loop.

Host side in Python (mirrors the reference's module-level API), hot path in
hand-written CUDA behind the C ABI of include/pic_b200.h (libpic_b200.so, bound
with ctypes in pypic_b200._lib).  PyTorch is used only for device memory,
streams and torch.distributed plumbing.  There is no CPU fallback.
"""
from ._lib import PicError, load  # noqa: F401

__version__ = "0.1.0"
