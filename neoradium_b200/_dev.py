"""Buffer plumbing between caller-owned NumPy arrays and device memory (PyTorch is only the buffer carrier)."""
import ctypes
import os

import numpy as np
import torch

from . import _managed, _native


def device():
    if not torch.cuda.is_available():
        raise _native.NrldpcError("neoradium_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    device()
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def handle():
    return _native.handle(device().index)


def to_dev(arr, dtype=None):
    """Host array -> contiguous device tensor of `dtype` (a torch dtype); conversion happens on the device."""
    if isinstance(arr, torch.Tensor):
        t = arr.to(device())
    else:
        a = np.ascontiguousarray(arr)
        if a.dtype == np.bool_:
            a = a.view(np.uint8)
        t = torch.from_numpy(a).to(device(), non_blocking=False)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


# ---- unified-memory arrays (HARQ state resident on the device behind host-visible ndarrays, _managed.py) --------------
MANAGED_MIN_BYTES = 1 << 19


def managed_ok(nbytes=None):
    """True when the drop-in classes may hand out a ManagedArray result (env NRLDPC_NO_MANAGED=1 turns them off).
    Below MANAGED_MIN_BYTES (env NRLDPC_MANAGED_MIN) a plain copy is cheaper than the bookkeeping of a managed block
    (measured on the Harq.ipynb loop, 270 KB buffers: 0.67 ms vs 0.81 ms per decodeLLRs), so small results stay plain."""
    if os.environ.get("NRLDPC_NO_MANAGED") or not _managed.supported(device().index):
        return False
    if nbytes is None:
        return True
    return nbytes >= int(os.environ.get("NRLDPC_MANAGED_MIN", MANAGED_MIN_BYTES))


def managed_out(shape, np_dtype, zero=False):
    return _managed.alloc(shape, np_dtype, device().index, stream_ptr(), zero)


def dev_in(arr, torch_dtype, np_dtype):
    """Device view of an input array: the array itself when it is a whole ManagedArray of the right dtype (the kernels
    read its pages in place, nothing is copied), otherwise a fresh device tensor (H2D copy + conversion)."""
    m = _managed.root_of(arr, np_dtype)
    if m is not None and m._nr_dev == device().index:
        return m
    return to_dev(arr, torch_dtype)


def sync():
    """Results in managed memory are read by the host right after the call returns: wait for the kernels."""
    torch.cuda.current_stream().synchronize()


def to_host(t):
    return t.cpu().numpy()
