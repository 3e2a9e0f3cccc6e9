"""Buffer plumbing between caller-owned NumPy arrays and device memory (PyTorch is only the buffer carrier)."""
import ctypes

import numpy as np
import torch

from . import _native


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


def to_host(t):
    return t.cpu().numpy()
