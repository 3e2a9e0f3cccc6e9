"""Gold-sequence scrambling on the device: mirrors ``goldSequence`` (neoradium/utils.py:70-94) and
``PDSCH.scrambleBits`` / ``PDSCH.scrambleLLRs`` (neoradium/pdsch.py:603-616) -- the steps between the LDPC chain and
the modem in every PDSCH loop.  ``cInit = rnti * 2^15 + q * 2^14 + nID`` as in the reference (pdsch.py:605)."""
import numpy as np
import torch

from . import _dev, _native


def goldSequence(cInit, numBits):
    """List of `numBits` bits of c(n) (the reference returns a Python list, utils.py:94)."""
    if numBits <= 0:
        return []
    out = torch.empty((int(numBits),), dtype=torch.int8, device=_dev.device())
    _native.check(_native.lib().nrldpc_gold_sequence(_dev.handle(), int(cInit), int(numBits), _dev.ptr(out), _dev.stream_ptr()))
    return _dev.to_host(out).astype(np.int64).tolist()


def scrambleBits(cInit, bits):
    """bits ^ c, same dtype and shape as the 1-D input (pdsch.py:603-608)."""
    b = np.asarray(bits)
    d = _dev.to_dev(b.reshape(-1), torch.int8)
    if d.numel():
        _native.check(_native.lib().nrldpc_scramble_bits(_dev.handle(), int(cInit), _dev.ptr(d), d.numel(), _dev.ptr(d), _dev.stream_ptr()))
    return _dev.to_host(d).astype(b.dtype if b.dtype != np.bool_ else np.int8).reshape(b.shape)


def scrambleLLRs(cInit, llrs):
    """llrs * (1 - 2 c) as float64 (pdsch.py:611-616)."""
    x = np.asarray(llrs, dtype=np.float64)
    d = _dev.to_dev(x.reshape(-1), torch.float64)
    if d.numel():
        _native.check(_native.lib().nrldpc_scramble_llrs(_dev.handle(), int(cInit), _native.F64, _dev.ptr(d), d.numel(), _dev.ptr(d), _dev.stream_ptr()))
    return _dev.to_host(d).reshape(x.shape)


def scramble_(cInit, tensor):
    """In-place device form: int8 bits are XORed with c, float32/float64 LLRs change sign where c = 1."""
    assert tensor.is_cuda and tensor.is_contiguous()
    L, h, s = _native.lib(), _native.handle(tensor.device.index), _dev.stream_ptr()
    if tensor.dtype == torch.int8:
        _native.check(L.nrldpc_scramble_bits(h, int(cInit), _dev.ptr(tensor), tensor.numel(), _dev.ptr(tensor), s))
    else:
        dt = {torch.float32: _native.F32, torch.float64: _native.F64}[tensor.dtype]
        _native.check(L.nrldpc_scramble_llrs(h, int(cInit), dt, _dev.ptr(tensor), tensor.numel(), _dev.ptr(tensor), s))
    return tensor
