"""NumPy arrays over CUDA managed (unified) memory -- HARQ state that stays on the device behind host-visible arrays.

The reference keeps ``HarqCW.encBuffer`` / ``HarqCW.decBuffer`` (neoradium/harq.py:120-121, 145-178) and the [C, N] LLR array
that ``recoverRate`` hands to ``decode`` (ldpc.py:1414-1418) as host NumPy arrays and passes them back into the codec on
the next call.  The drop-in classes return ``ManagedArray`` objects for these: ordinary ``numpy.ndarray`` instances (every
NumPy access path works, C-level ones included, because the memory IS host-addressable) whose pages live on the GPU until
the CPU touches them.  When such an array comes back into the codec, the kernels use its device pointer directly -- no
H2D / D2H copy per retransmission.  Coherence is the hardware's job (page migration on access), not
a Python-level mirror, so there is no stale-copy hazard.

Only ROOT arrays (as returned by ``alloc``) carry the device pointer; views and copies are plain host data and take the
normal upload path.
"""
import atexit
import ctypes
import weakref

import numpy as np
import torch

from . import _native

_TORCH_OF = {np.dtype(np.int8): torch.int8, np.dtype(np.uint8): torch.uint8, np.dtype(np.float32): torch.float32,
             np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32}

_pool = {}            # (device index, nbytes) -> [pointer, ...]: freed blocks kept for reuse (cudaFree synchronises)
_pool_bytes = [0]
_POOL_CAP = 1 << 29
_alive = [True]
_supported = {}


@atexit.register
def _shutdown():
    _alive[0] = False


class ManagedArray(np.ndarray):
    """ndarray over managed memory; ``data_ptr()`` / ``device`` make it usable wherever the binding takes a device tensor."""
    _nr_ptr = None
    _nr_dev = None

    def __array_finalize__(self, obj):
        self._nr_ptr = None     # views / results of operations are plain host data
        self._nr_dev = None

    def data_ptr(self):
        return self._nr_ptr

    @property
    def device(self):
        return torch.device("cuda", self._nr_dev)

    def __reduce__(self):       # pickles as a plain array
        return np.asarray(self).__reduce__()


def supported(device_index):
    v = _supported.get(device_index)
    if v is None:
        v = bool(_native.lib().nrldpc_managed_supported(_native.handle(device_index)))
        _supported[device_index] = v
    return v


def _release(dev, nbytes, p):
    if not _alive[0]:
        return                  # interpreter shutdown: the CUDA context may be gone already
    if _pool_bytes[0] + nbytes <= _POOL_CAP:
        _pool.setdefault((dev, nbytes), []).append(p)
        _pool_bytes[0] += nbytes
        return
    try:
        _native.lib().nrldpc_managed_free(_native.handle(dev), ctypes.c_void_p(p))
    except Exception:
        pass


def alloc(shape, dtype, device_index, stream_ptr, zero=False):
    """A ROOT ManagedArray of `shape` / `dtype`, resident on the device (cleared there when `zero`)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    nbytes = max(4096, (n + 4095) & ~4095)
    L = _native.lib()
    h = _native.handle(device_index)
    free = _pool.get((device_index, nbytes))
    if free:
        p = free.pop()
        _pool_bytes[0] -= nbytes
        _native.check(L.nrldpc_managed_prefetch(h, ctypes.c_void_p(p), nbytes, 1, stream_ptr))
        if zero:
            _native.check(L.nrldpc_managed_clear(h, ctypes.c_void_p(p), nbytes, stream_ptr))
    else:
        out = ctypes.c_void_p()
        _native.check(L.nrldpc_managed_alloc(h, nbytes, 1 if zero else 0, ctypes.byref(out), stream_ptr))
        p = out.value
    buf = (ctypes.c_byte * nbytes).from_address(p)
    weakref.finalize(buf, _release, device_index, nbytes, p)
    a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape).view(ManagedArray)
    a._nr_ptr = p
    a._nr_dev = device_index
    return a


def root_of(x, dtype=None, shape=None):
    """`x` itself if it is a whole, C-contiguous ROOT ManagedArray (optionally of the given dtype / shape), else None."""
    if not isinstance(x, ManagedArray) or x._nr_ptr is None:
        return None
    if not x.flags.c_contiguous or x.ctypes.data != x._nr_ptr:
        return None
    if dtype is not None and x.dtype != np.dtype(dtype):
        return None
    if shape is not None and tuple(x.shape) != tuple(shape):
        return None
    return x


def torch_dtype(x):
    return _TORCH_OF[x.dtype]
