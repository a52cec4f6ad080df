"""Thin helpers between torch device tensors and the raw pointers the C ABI takes."""
import ctypes as C
import threading

import numpy as np
import torch

from . import _lib


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise _lib.PicError(_lib.PIC_ERR_NODEVICE, "no CUDA device: pypic_b200 has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    _lib.load()
    return dev


def ptr(t):
    """Device (or host) address of a tensor / numpy array / None."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return t.data_ptr()
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    raise TypeError(type(t))


def stream():
    return torch.cuda.current_stream().cuda_stream


def f64(n, device, zero=False):
    return (torch.zeros if zero else torch.empty)(int(n), dtype=torch.float64, device=device)


def to_dev(a, device, dtype=torch.float64):
    """numpy -> device tensor (contiguous copy)."""
    return torch.as_tensor(np.ascontiguousarray(a)).to(device=device, dtype=dtype)


_pin = threading.local()
PIN_MAX = 1 << 20


def _pinned(nbytes):
    """Per-thread page-locked landing buffer for the small per-step reads (loop statistics, absorption
    log): a copy into pageable memory is staged by the driver and costs ~3x as long (the device idles
    meanwhile: every step ends in one of these reads)."""
    b = getattr(_pin, "buf", None)
    if b is None:
        b = _pin.buf = torch.empty(PIN_MAX, dtype=torch.uint8, pin_memory=True).numpy()
    return b[:nbytes]


def read_raw(t, n, dtype):
    """Synchronous device->host read of n elements through the C ABI."""
    nbytes = int(n) * np.dtype(dtype).itemsize
    if 0 < nbytes <= PIN_MAX:
        land = _pinned(nbytes)
        _lib.call("pic_dev_read", ptr(t), land.ctypes.data, nbytes, stream())
        return land.view(dtype).copy()
    out = np.empty(n, dtype=dtype)
    _lib.call("pic_dev_read", ptr(t), out.ctypes.data, out.nbytes, stream())
    return out


def read_f64(t, n=None):
    """Synchronous device->host read of a small fp64 tensor through the C ABI."""
    return read_raw(t, t.numel() if n is None else n, np.float64)


def check_range(err_t, what):
    """Raises if a kernel reported out-of-grid particle indices (reference UB)."""
    n = int(read_raw(err_t, 1, np.int32)[0])
    if n:
        err_t.zero_()
        raise _lib.PicError(_lib.PIC_ERR_RANGE, "%s: %d particle position(s) outside the grid "
                            "(undefined behaviour in the reference); indices were clamped" % (what, n))


def sort_counts_size(Ng):
    """int32 entries of the scratch pic_dev_dd_sort_by_cell / pic_dev_sort_perm_by_cell need:
    2*Ng keys + 2, plus the block sums of the three-pass scan used for large grids."""
    nk = 2 * int(Ng)
    return nk + 2 + (nk + 1023) // 1024 + 2


def sort_stable_scratch_size(N):
    """int32 entries of the scratch pic_dev_dd_sort_by_cell_stable needs for species blocks of at
    most N particles: the [digit][tile] histogram of one radix pass plus the block sums of its scan."""
    nh = 256 * ((int(N) + 4095) // 4096)
    return nh + (nh + 1023) // 1024 + 2
